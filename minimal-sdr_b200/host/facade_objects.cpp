// facade_objects.cpp — stand-alone façade objects (not fused): AudioFilterBiquad with the reference's designer calls and
// AudioEffectFreqConv, driven block by block through AudioConnection/update_all like any Teensy Audio graph.
//   facade_objects <channels> <blocks> <in.bin> <out_biquad.bin> <out_fcI.bin> <out_fcQ.bin>
// in.bin: int16 [channels][blocks*128].  Biquad cascade: setLowpass(0, 3000, 0.7071); setNotch(1, 5514.7, 15) and, from block 3 on,
// setCoefficients(0, double[5]{0.2, 0.4, 0.2, -0.5, 0.3}).  Frequency converter: I = in, Q = in reversed in time per block,
// default fs/4 tables, direction(1) from block 2 on, passthrough(0) for the last block.
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
  bool reverse = false;
  virtual void update(void)
  {
    audio_block_t *b = allocate();
    if (!b) return;
    for (uint32_t c = 0; c < b->channels; ++c)
      for (int n = 0; n < AUDIO_BLOCK_SAMPLES; ++n)
        b->data[(size_t)c * AUDIO_BLOCK_SAMPLES + n] = src[c * stride + (size_t)block_index * AUDIO_BLOCK_SAMPLES + (reverse ? AUDIO_BLOCK_SAMPLES - 1 - n : n)];
    ++block_index;
    transmit(b);
    release(b);
  }
};

Injector srcA, srcI, srcQ;
AudioFilterBiquad biquad;
AudioEffectFreqConv conv;
AudioCapture capB, capI, capQ;
AudioConnection c1(srcA, biquad), c2(biquad, capB);
AudioConnection c3(srcI, 0, conv, 0), c4(srcQ, 0, conv, 1), c5(conv, 0, capI, 0), c6(conv, 1, capQ, 0);

int main(int argc, char **argv)
{
  if (argc != 7) return 2;
  const uint32_t C = (uint32_t)atoi(argv[1]), NB = (uint32_t)atoi(argv[2]);
  const size_t L = (size_t)NB * AUDIO_BLOCK_SAMPLES;
  std::vector<int16_t> in((size_t)C * L), ob((size_t)C * L), oi((size_t)C * L), oq((size_t)C * L);
  FILE *f = fopen(argv[3], "rb");
  if (!f || fread(in.data(), 2, in.size(), f) != in.size()) return 2;
  fclose(f);
  AudioMemory(C, 24);
  for (Injector *s : {&srcA, &srcI, &srcQ}) { s->src = in.data(); s->stride = L; }
  srcQ.reverse = true;
  biquad.setLowpass(0, 3000.0f, 0.7071f);
  biquad.setNotch(1, 5514.7f, 15.0f);
  biquad.setCoefficients(7, (const int *)nullptr); // stage >= 4: ignored before touching the pointer (filter_biquad.cpp:86)
  for (uint32_t b = 0; b < NB; ++b) {
    if (b == 3) { const double c[5] = {0.2, 0.4, 0.2, -0.5, 0.3}; biquad.setCoefficients(0, c); }
    if (b == 2) conv.direction(true);
    if (b + 1 == NB) conv.passthrough(false);
    AudioStream::update_all();
    if (capB.blocks != b + 1 || capI.blocks != b + 1 || capQ.blocks != b + 1) { fprintf(stderr, "graph stalled at block %u\n", b); return 4; }
    for (uint32_t c = 0; c < C; ++c) {
      memcpy(ob.data() + c * L + (size_t)b * AUDIO_BLOCK_SAMPLES, capB.last.data() + (size_t)c * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
      memcpy(oi.data() + c * L + (size_t)b * AUDIO_BLOCK_SAMPLES, capI.last.data() + (size_t)c * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
      memcpy(oq.data() + c * L + (size_t)b * AUDIO_BLOCK_SAMPLES, capQ.last.data() + (size_t)c * AUDIO_BLOCK_SAMPLES, AUDIO_BLOCK_SAMPLES * 2);
    }
  }
  if (biquad.last_status() != MSDR_OK) { fprintf(stderr, "biquad status %d\n", biquad.last_status()); return 3; }
  const char *paths[3] = {argv[4], argv[5], argv[6]};
  std::vector<int16_t> *bufs[3] = {&ob, &oi, &oq};
  for (int i = 0; i < 3; ++i) {
    FILE *o = fopen(paths[i], "wb");
    if (!o) return 2;
    fwrite(bufs[i]->data(), 2, bufs[i]->size(), o);
    fclose(o);
  }
  printf("facade_objects: %u channels x %u blocks, pool max %u\n", C, NB, AudioMemoryUsageMax());
  return 0;
}
