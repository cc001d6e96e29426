/* TEST INFRASTRUCTURE — minimal AudioStream stand-in (the Teensy core header is not vendored in
 * the reference).  Contract inferred from use: filter_biquad.cpp:39-41,80-81 (receiveWritable,
 * transmit, release), filter_biquad.h:36 (ctor AudioStream(ninputs, queue)), mixer.cpp:134-159.
 * One input slot, one output slot; the test driver puts a block in `in_slot` and reads `out_slot`. */
#ifndef MSDR_STUB_AUDIOSTREAM_H
#define MSDR_STUB_AUDIOSTREAM_H
#include <stdint.h>
#include <stddef.h>
#define AUDIO_BLOCK_SAMPLES 128
#ifndef AUDIO_SAMPLE_RATE_EXACT
#define AUDIO_SAMPLE_RATE_EXACT 44117.64706
#endif
#define AUDIO_SAMPLE_RATE AUDIO_SAMPLE_RATE_EXACT
typedef struct audio_block_struct {
  uint8_t ref_count;
  uint8_t reserved1;
  uint16_t memory_pool_index;
  int16_t data[AUDIO_BLOCK_SAMPLES];
} audio_block_t;

class AudioStream {
public:
  AudioStream(unsigned char ninput, audio_block_t **iqueue) : num_inputs(ninput), inputQueue(iqueue)
  { for (int i = 0; i < 4; i++) { in_slot[i] = NULL; out_slot[i] = NULL; } }
  virtual ~AudioStream() {}
  virtual void update(void) = 0;
  audio_block_t *in_slot[4];
  audio_block_t *out_slot[4];
protected:
  audio_block_t *receiveReadOnly(unsigned int index = 0) { audio_block_t *b = in_slot[index]; in_slot[index] = NULL; return b; }
  audio_block_t *receiveWritable(unsigned int index = 0) { audio_block_t *b = in_slot[index]; in_slot[index] = NULL; return b; }
  void transmit(audio_block_t *block, unsigned char index = 0) { out_slot[index] = block; }
  static void release(audio_block_t *) {}
  static audio_block_t *allocate(void) { return NULL; }
  unsigned char num_inputs;
  audio_block_t **inputQueue;
};
#endif
