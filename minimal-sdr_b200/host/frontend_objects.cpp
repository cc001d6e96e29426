// frontend_objects.cpp — the front end of the sketch on the façade: raw ADC codes -> Frontend (DC block, amp_adc, AGC) block by
// block like the audio interrupt delivers them, and a stand-alone AudioAmplifier driven through AudioConnection/update_all.
//   frontend_objects <channels> <blocks> <codes.bin> <out_frontend.bin> <out_amp.bin> <out_agc.bin>
// codes.bin: uint16 [channels][blocks*128].  out_frontend: int16 conditioned samples.  out_amp: the same input reinterpreted as
// int16 through amp.gain(1.7) (blocks 0-2), gain(1.0) (block 3), gain(0) (block 4: nothing transmitted, zeros written here),
// gain(-0.33) afterwards.  out_agc: float AGC_val of every channel after the last block.
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
      for (int n = 0; n < AUDIO_BLOCK_SAMPLES; ++n) b->data[(size_t)c * AUDIO_BLOCK_SAMPLES + n] = src[c * stride + (size_t)block_index * AUDIO_BLOCK_SAMPLES + n];
    ++block_index;
    transmit(b);
    release(b);
  }
};

Injector src;
AudioAmplifier amp;
AudioCapture cap;
AudioConnection c1(src, amp), c2(amp, cap);

int main(int argc, char **argv)
{
  if (argc != 7) return 2;
  const uint32_t C = (uint32_t)atoi(argv[1]), NB = (uint32_t)atoi(argv[2]);
  const size_t L = (size_t)NB * AUDIO_BLOCK_SAMPLES;
  std::vector<uint16_t> codes((size_t)C * L);
  std::vector<int16_t> ofe((size_t)C * L), oamp((size_t)C * L, 0);
  FILE *f = fopen(argv[3], "rb");
  if (!f || fread(codes.data(), 2, codes.size(), f) != codes.size()) return 2;
  fclose(f);

  Frontend fe(C);
  if (!fe.ok()) { fprintf(stderr, "%s\n", msdr_frontend_last_error(nullptr)); return 3; }
  fe.begin(codes[0]);
  for (uint32_t b = 0; b < NB; ++b) // one audio block per call, state carried inside the object
    if (fe.update(codes.data() + (size_t)b * AUDIO_BLOCK_SAMPLES, ofe.data() + (size_t)b * AUDIO_BLOCK_SAMPLES, 1, L) != MSDR_OK) return 4;

  AudioMemory(C, 8);
  src.src = reinterpret_cast<const int16_t *>(codes.data());
  src.stride = L;
  for (uint32_t b = 0; b < NB; ++b) {
    amp.gain(b < 3 ? 1.7f : b == 3 ? 1.0f : b == 4 ? 0.0f : -0.33f);
    const unsigned before = cap.blocks;
    AudioStream::update_all();
    if (cap.blocks == before) continue; // zero gain: no block arrived
    for (uint32_t c = 0; c < C; ++c)
      for (int n = 0; n < AUDIO_BLOCK_SAMPLES; ++n) oamp[c * L + (size_t)b * AUDIO_BLOCK_SAMPLES + n] = cap.last[(size_t)c * AUDIO_BLOCK_SAMPLES + n];
  }
  std::vector<float> agc(C);
  for (uint32_t c = 0; c < C; ++c) agc[c] = fe.AGC_val(c);
  f = fopen(argv[4], "wb"); fwrite(ofe.data(), 2, ofe.size(), f); fclose(f);
  f = fopen(argv[5], "wb"); fwrite(oamp.data(), 2, oamp.size(), f); fclose(f);
  f = fopen(argv[6], "wb"); fwrite(agc.data(), 4, agc.size(), f); fclose(f);
  return 0;
}
