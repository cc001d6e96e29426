// sketch_port.cpp — the DSP skeleton of Minimal-SDR.ino rebuilt on the façade (include/msdr/Audio.h), for a batch of
// channels.  Same objects, same patch cords, same setup()/loop() shape as the sketch (Minimal-SDR.ino:66-81, 372-442);
// the Teensy ADC/DAC objects are replaced by a file-fed injector and a capture sink.
//
//   sketch_port <channels> <blocks> <mode> <tables.bin> <in.bin> <out.bin>
//     tables.bin : int32 numTaps, int16 cI[numTaps], int16 cQ[numTaps], int32 lowpass[5], int32 notch[5]
//     in.bin     : int16 [channels][blocks*128]   raw IF samples
//     out.bin    : int16 [channels][blocks*128]   audio as it leaves biquad2_dac
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "msdr/Audio.h"

using namespace msdr;

// stands in for adc1 -> amp_adc: emits one batch block per audio interrupt from a host array
class AudioInjector : public AudioStream {
public:
  AudioInjector() : AudioStream(0, nullptr) {}
  const int16_t *src = nullptr;
  size_t stride = 0;
  uint32_t block_index = 0, n_blocks = 0;
  virtual void update(void)
  {
    if (!src || block_index >= n_blocks) return;
    audio_block_t *b = allocate();
    if (!b) return;
    for (uint32_t c = 0; c < b->channels; ++c)
      memcpy(b->data + (size_t)c * AUDIO_BLOCK_SAMPLES, src + c * stride + (size_t)block_index * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
    ++block_index;
    transmit(b);
    release(b);
  }
};

// the sketch's graph (Minimal-SDR.ino:66-81)
AudioInjector adc1;                 // AudioInputAnalog adc1 + AudioAmplifier amp_adc
AudioRecordQueue queue_adc;
AudioPlayQueue queue_dac;
AudioFilterBiquad biquad1_dac;
AudioFilterBiquad biquad2_dac;
AudioCapture dac1;                  // AudioAmplifier amp_dac (unity) + AudioOutputAnalog dac1
AudioConnection patchCord3(adc1, queue_adc);
AudioConnection patchCord2(queue_dac, biquad1_dac);
AudioConnection patchCord4a(biquad1_dac, biquad2_dac);
AudioConnection patchCord5(biquad2_dac, dac1);

static bool read_all(const char *path, void *dst, size_t bytes)
{
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  const size_t n = fread(dst, 1, bytes, f);
  fclose(f);
  return n == bytes;
}

int main(int argc, char **argv)
{
  if (argc != 7) { fprintf(stderr, "usage: %s channels blocks mode tables.bin in.bin out.bin\n", argv[0]); return 2; }
  const uint32_t C = (uint32_t)atoi(argv[1]), NB = (uint32_t)atoi(argv[2]);
  const int mode = atoi(argv[3]);
  const size_t L = (size_t)NB * AUDIO_BLOCK_SAMPLES;

  FILE *tf = fopen(argv[4], "rb");
  if (!tf) { perror("tables"); return 2; }
  int32_t numTaps = 0, lowpass[5], notch[5];
  if (fread(&numTaps, 4, 1, tf) != 1) return 2;
  std::vector<int16_t> cI(numTaps), cQ(numTaps);
  if (fread(cI.data(), 2, numTaps, tf) != (size_t)numTaps || fread(cQ.data(), 2, numTaps, tf) != (size_t)numTaps) return 2;
  if (fread(lowpass, 4, 5, tf) != 5 || fread(notch, 4, 5, tf) != 5) return 2;
  fclose(tf);

  std::vector<int16_t> in((size_t)C * L), out((size_t)C * L, 0);
  if (!read_all(argv[5], in.data(), in.size() * 2)) { perror("in.bin"); return 2; }

  // ---- setup() (Minimal-SDR.ino:372-411)
  AudioMemory(C, 20);                                  // AudioMemory(AUDIOMEMORY)
  Receiver rx(C);
  if (!rx.ok()) { fprintf(stderr, "msdr_chain_create: %s\n", msdr_last_error(nullptr)); return 3; }
  rx.bind(biquad1_dac, biquad2_dac);
  biquad1_dac.setCoefficients(0, lowpass);             // biquad1_dac.setLowpass(0, IF*0.9*CORR_FACT, 0.54)  (.ino:391-393), integer form
  rx.set_mode(mode);                                   // mode = ...; tune():
  if (rx.init_FIR((uint16_t)numTaps, cI.data(), cQ.data()) != MSDR_OK) { fprintf(stderr, "init_FIR: %s\n", rx.last_error()); return 3; }
  biquad2_dac.setCoefficients(0, notch);               // biquad2_dac.setNotch(0, pdb_freq/8*CORR_FACT, 15.0) (.ino:356), integer form
  queue_adc.begin();                                   // .ino:410

  // ---- audio interrupts + loop() (Minimal-SDR.ino:436-442)
  adc1.src = in.data(); adc1.stride = L; adc1.n_blocks = NB;
  uint32_t captured = 0;
  unsigned long processed = 0;
  for (uint32_t tick = 0; tick < NB + 4 && captured < NB; ++tick) {
    AudioStream::update_all();                         // one audio interrupt: adc1 -> queue_adc ; queue_dac -> biquads -> dac1
    if (dac1.blocks > captured) {                      // a block left the graph
      for (uint32_t c = 0; c < C; ++c)
        memcpy(out.data() + c * L + (size_t)captured * AUDIO_BLOCK_SAMPLES, dac1.last.data() + (size_t)c * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
      captured = dac1.blocks;
    }
    processed += rx.demodulation(queue_adc, queue_dac); // loop(): time_needed = demodulation();
  }
  if (rx.last_status() != MSDR_OK) { fprintf(stderr, "demodulation: %s\n", rx.last_error()); return 3; }
  FILE *of = fopen(argv[6], "wb");
  if (!of) { perror("out.bin"); return 2; }
  fwrite(out.data(), 2, out.size(), of);
  fclose(of);
  printf("sketch_port: %u channels, %lu blocks demodulated, %u captured, pool max %u blocks\n", C, processed, captured, AudioMemoryUsageMax());
  return captured == NB ? 0 : 4;
}
