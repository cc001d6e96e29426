/* TEST INFRASTRUCTURE — the REFERENCE's front-end conditioning, compiled from /root/reference by oracle/Makefile:
 *   - the DC-blocking loop of AudioInputAnalog::update (src/Audio/input_adc.cpp:198-212 and the coefficient define :32),
 *     extracted by line range at build time into _ref/hpf_extract.inc / hpf_coef_extract.inc (the rest of that file is
 *     Kinetis ADC/PDB/DMA register code that cannot compile on a host);
 *   - AudioAmplifier (src/Audio/mixer.{h,cpp}), the whole translation unit, against the AudioStream stub;
 *   - AGC() (Minimal-SDR.ino:445-515), extracted by line range into _ref/agc_extract.inc.  One sed edit in the Makefile
 *     gives `agc_buffer` a guard element in front: the reference stores every 26th value at index -1 (.ino:481), which
 *     is undefined behaviour; with the guard that store lands in a defined place and is never read.
 * The ARMv7E-M intrinsics AGC() uses are emulated below from their architectural definition (APSR.GE flags). */
#include "mixer.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int32_t FRACMUL_SHL_ref(int32_t x, int32_t y, int z) { return FRACMUL_SHL(x, y, z); }

#include "hpf_coef_extract.inc" /* #define COEF_HPF_DCBLOCK ... */

/* ---- intrinsics for the AGC code ---------------------------------------------------------------------------------- */
static uint32_t g_ge;
static inline uint32_t __SSUB16(uint32_t a, uint32_t b)
{
  const int32_t lo = (int16_t)(a & 0xFFFF) - (int16_t)(b & 0xFFFF), hi = (int16_t)(a >> 16) - (int16_t)(b >> 16);
  g_ge = (lo >= 0 ? 3u : 0u) | (hi >= 0 ? 12u : 0u);
  return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu);
}
static inline uint32_t __SEL(uint32_t a, uint32_t b)
{
  const uint32_t m = ((g_ge & 1u) ? 0x000000FFu : 0u) | ((g_ge & 2u) ? 0x0000FF00u : 0u) | ((g_ge & 4u) ? 0x00FF0000u : 0u) | ((g_ge & 8u) ? 0xFF000000u : 0u);
  return (a & m) | (b & ~m);
}
#define __SIMD32(addr) (*(int32_t **)&(addr)) /* arm_math.h */

/* globals the extracted AGC() refers to (.ino:67,94-104) */
static AudioAmplifier amp_adc;
static float AGC_Max = 40.0f;
static int AGC_on = 1;
static float AGC_val = 0.25f;

#include "agc_extract.inc" /* void AGC(int16_t * block) */

namespace {
struct AmpView : public AudioStream { AmpView() : AudioStream(1, NULL) {} void update() {} int32_t multiplier; };
int32_t amp_multiplier(AudioAmplifier &a) { return reinterpret_cast<AmpView *>(static_cast<AudioStream *>(&a))->multiplier; }
}

extern "C" {
/* input_adc.cpp:198-212 on one block; data128 holds the raw ADC codes (bit patterns) and receives the result */
void ref_adc_hpf_block(int16_t *data128, int32_t *x1, int32_t *y1)
{
  struct { int16_t *data; } blk = {data128}, *out_left = &blk;
  int32_t tmp;
  int16_t s, *p, *end;
  int32_t hpf_x1 = *x1, hpf_y1 = *y1;
#include "hpf_extract.inc"
  *x1 = hpf_x1; *y1 = hpf_y1;
}
int32_t ref_amp_multiplier(float gain)
{
  AudioAmplifier a;
  a.gain(gain);
  return amp_multiplier(a);
}
/* AudioAmplifier::update on one block; returns 0 when the object transmitted nothing */
int ref_amp_block(float gain, int16_t *data128)
{
  AudioAmplifier a;
  audio_block_t blk;
  a.gain(gain);
  memcpy(blk.data, data128, sizeof(blk.data));
  a.in_slot[0] = &blk;
  a.out_slot[0] = NULL;
  a.update();
  if (!a.out_slot[0]) return 0;
  memcpy(data128, a.out_slot[0]->data, sizeof(blk.data));
  return 1;
}
/* the sketch's AGC() is a function with static state: one process-wide instance.  reset = re-create that state is not
 * possible (function statics), so a test drives ONE stream through it per process and compares the whole trajectory. */
void ref_agc_config(float start, float max, int on) { AGC_val = start; AGC_Max = max; AGC_on = on; amp_adc.gain(start); }
void ref_agc_block(int16_t *block128, float *agc_val, int32_t *mult)
{
  AGC(block128);
  *agc_val = AGC_val;
  *mult = amp_multiplier(amp_adc);
}
} /* extern "C" */
