/* TEST INFRASTRUCTURE — AudioStream stand-in for compiling the reference's freq_conv.cpp (two inputs, two outputs, and a
 * working allocate(): AudioEffectFreqConv::update takes four scratch blocks, freq_conv.cpp:58-64).  Contract inferred from
 * freq_conv.cpp:37-113.  `fail_alloc` makes allocate() return NULL to exercise the silent-drop path (freq_conv.cpp:64,111). */
#ifndef MSDR_STUB_FC_AUDIOSTREAM_H
#define MSDR_STUB_FC_AUDIOSTREAM_H
#include <stdint.h>
#include <stddef.h>
#define AUDIO_BLOCK_SAMPLES 128
typedef struct audio_block_struct {
  uint8_t ref_count;
  uint8_t reserved1;
  uint16_t memory_pool_index;
  int16_t data[AUDIO_BLOCK_SAMPLES];
} audio_block_t;

class AudioStream {
public:
  AudioStream(unsigned char ninput, audio_block_t **iqueue) : num_inputs(ninput), inputQueue(iqueue)
  { for (int i = 0; i < 2; i++) { in_slot[i] = NULL; out_slot[i] = NULL; } }
  virtual ~AudioStream() {}
  virtual void update(void) = 0;
  audio_block_t *in_slot[2];
  audio_block_t *out_slot[2];
  static int fail_alloc;
protected:
  audio_block_t *receiveReadOnly(unsigned int index = 0) { audio_block_t *b = in_slot[index]; in_slot[index] = NULL; return b; }
  audio_block_t *receiveWritable(unsigned int index = 0) { return receiveReadOnly(index); }
  void transmit(audio_block_t *block, unsigned char index = 0) { out_slot[index] = block; }
  static void release(audio_block_t *) {}
  static audio_block_t *allocate(void)
  {
    static audio_block_t pool[8];
    static unsigned next = 0;
    if (fail_alloc) return NULL;
    return &pool[next++ & 7u];
  }
  unsigned char num_inputs;
  audio_block_t **inputQueue;
};
#endif
