// fir_usage_objects.cpp — AudioFilterFIR in an AudioConnection graph, and AudioProcessorUsageMax()/…Reset() fed by a Receiver.
//   fir_usage_objects <channels> <blocks> <taps.bin> <in.bin> <out_fir.bin> <sent.bin>
// taps.bin: int32 n, int16[n].  Blocks 0-2: begin(taps, n); block 3: end() (input consumed, nothing transmitted); block 4:
// begin(FIR_PASSTHRU); block 5: begin(taps, n - 1) (odd count: init fails like arm_fir_init_q15, nothing transmitted); from block 6:
// begin(taps, n) again (delay line zeroed).  sent.bin: one byte per block, 1 = a block reached the sink.
// Then a Receiver runs the same input (AM table = taps on both branches) and the usage figures are printed and sanity-checked.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "msdr/Audio.h"
using namespace msdr;

class Injector : public AudioStream {
public:
  Injector() : AudioStream(0, nullptr) {}
  const int16_t *src = nullptr;
  size_t stride = 0;
  uint32_t block_index = 0;
  virtual void update(void)
  {
    audio_block_t *b = allocate();
    if (!b) return;
    for (uint32_t c = 0; c < b->channels; ++c)
      memcpy(b->data + (size_t)c * AUDIO_BLOCK_SAMPLES, src + c * stride + (size_t)block_index * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
    ++block_index;
    transmit(b);
    release(b);
  }
};

Injector src;
AudioFilterFIR fir;
AudioCapture cap;
AudioConnection c1(src, fir), c2(fir, cap);

int main(int argc, char **argv)
{
  if (argc != 7) return 2;
  const uint32_t C = (uint32_t)atoi(argv[1]), NB = (uint32_t)atoi(argv[2]);
  const size_t L = (size_t)NB * AUDIO_BLOCK_SAMPLES;
  FILE *f = fopen(argv[3], "rb");
  int32_t n = 0;
  if (!f || fread(&n, 4, 1, f) != 1 || n < 4 || n > 256) return 2;
  std::vector<int16_t> taps((size_t)n);
  if (fread(taps.data(), 2, taps.size(), f) != taps.size()) return 2;
  fclose(f);
  std::vector<int16_t> in((size_t)C * L), out((size_t)C * L, 0);
  std::vector<unsigned char> sent(NB, 0);
  f = fopen(argv[4], "rb");
  if (!f || fread(in.data(), 2, in.size(), f) != in.size()) return 2;
  fclose(f);
  AudioMemory(C, 12);
  src.src = in.data();
  src.stride = L;
  fir.begin(taps.data(), n);
  for (uint32_t b = 0; b < NB; ++b) {
    if (b == 3) fir.end();
    if (b == 4) fir.begin(FIR_PASSTHRU, 0);
    if (b == 5) fir.begin(taps.data(), n - 1);
    if (b == 6) fir.begin(taps.data(), n);
    const unsigned before = cap.blocks;
    AudioStream::update_all();
    sent[b] = cap.blocks != before;
    if (sent[b])
      for (uint32_t c = 0; c < C; ++c)
        memcpy(out.data() + c * L + (size_t)b * AUDIO_BLOCK_SAMPLES, cap.last.data() + (size_t)c * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
  }
  if (fir.last_status() != MSDR_OK) { fprintf(stderr, "fir status %d\n", fir.last_status()); return 3; }
  f = fopen(argv[5], "wb");
  if (!f) return 2;
  fwrite(out.data(), 2, out.size(), f);
  fclose(f);
  f = fopen(argv[6], "wb");
  if (!f) return 2;
  fwrite(sent.data(), 1, sent.size(), f);
  fclose(f);

  // AudioProcessorUsageMax(): nothing counted before the first call, then the largest update, then 0 again after the reset
  Receiver rx(C);
  if (!rx.ok() || rx.init_FIR((uint16_t)n, taps.data(), taps.data()) != MSDR_OK) { fprintf(stderr, "receiver: %s\n", rx.last_error()); return 5; }
  AudioProcessorUsageSampleRate() = 24000.0; // the sketch's SAMPLE_RATE (Minimal-SDR.ino:85)
  const float u0 = AudioProcessorUsageMax();
  std::vector<int16_t> audio((size_t)C * L);
  if (rx.update(in.data(), audio.data(), NB, L) != MSDR_OK) return 5;
  const float last = AudioProcessorUsage(), mx = AudioProcessorUsageMax();
  if (rx.update(in.data(), audio.data(), 1, L) != MSDR_OK) return 5; // one block: launch overhead per block is larger
  const float mx2 = AudioProcessorUsageMax();
  AudioProcessorUsageMaxReset();
  const float after = AudioProcessorUsageMax();
  printf("fir_usage_objects: usage before %.4f%%, last %.4f%%, max %.4f%%, max after a 1-block update %.4f%%, after reset %.4f%%\n", u0, last, mx, mx2, after);
  if (u0 != 0.0f || !(last > 0.0f) || mx < last || mx2 < mx || after != 0.0f) return 6;
  return 0;
}
