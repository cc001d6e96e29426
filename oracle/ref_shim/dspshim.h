/* TEST INFRASTRUCTURE — portable C versions of the Cortex-M4 inline-asm helpers in the reference's
 * src/Audio/utility/dspinst.h, pre-included so that header (guard dspinst_h_, dspinst.h:27-28)
 * expands to nothing.  Semantics from the comments/asm at dspinst.h:32-51 (SSAT #bits, Rn, ASR #s),
 * :69-92 (SMULWB/SMULWT), :173-184 (PKHBT ... LSL #16), :233-249 (SMLAWB/SMLAWT: top 32 bits of the
 * 48-bit product, accumulate with 32-bit wrap), :358-368 (FRACMUL_SHL).  Build with -fwrapv. */
#ifndef dspinst_h_
#define dspinst_h_
#include <stdint.h>
static inline int32_t signed_saturate_rshift(int32_t val, int bits, int rshift)
{
  int32_t out = val >> rshift;
  int32_t max = (int32_t)((1u << (bits - 1)) - 1u);
  if (out > max) out = max;
  if (out < -max - 1) out = -max - 1;
  return out;
}
static inline int16_t saturate16(int32_t val)
{
  if (val > 32767) val = 32767; else if (val < -32768) val = -32768;
  return (int16_t)val;
}
static inline int32_t signed_multiply_32x16b(int32_t a, uint32_t b)
{
  return (int32_t)(((int64_t)a * (int16_t)(b & 0xFFFF)) >> 16);
}
static inline int32_t signed_multiply_32x16t(int32_t a, uint32_t b)
{
  return (int32_t)(((int64_t)a * (int16_t)(b >> 16)) >> 16);
}
static inline uint32_t pack_16t_16t(int32_t a, int32_t b) { return ((uint32_t)a & 0xFFFF0000u) | ((uint32_t)b >> 16); }
static inline uint32_t pack_16t_16b(int32_t a, int32_t b) { return ((uint32_t)a & 0xFFFF0000u) | ((uint32_t)b & 0x0000FFFFu); }
static inline uint32_t pack_16b_16b(int32_t a, int32_t b) { return ((uint32_t)a << 16) | ((uint32_t)b & 0x0000FFFFu); }
static inline int32_t signed_multiply_accumulate_32x16b(int32_t sum, int32_t a, uint32_t b)
{
  return (int32_t)((uint32_t)sum + (uint32_t)(int32_t)(((int64_t)a * (int16_t)(b & 0xFFFF)) >> 16));
}
static inline int32_t signed_multiply_accumulate_32x16t(int32_t sum, int32_t a, uint32_t b)
{
  return (int32_t)((uint32_t)sum + (uint32_t)(int32_t)(((int64_t)a * (int16_t)(b >> 16)) >> 16));
}
static inline uint32_t signed_add_16_and_16(uint32_t a, uint32_t b)
{
  int32_t lo = (int16_t)(a & 0xFFFF) + (int16_t)(b & 0xFFFF);
  int32_t hi = (int16_t)(a >> 16) + (int16_t)(b >> 16);
  if (lo > 32767) lo = 32767; else if (lo < -32768) lo = -32768;
  if (hi > 32767) hi = 32767; else if (hi < -32768) hi = -32768;
  return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu);
}
static inline uint32_t logical_and(uint32_t a, uint32_t b) { return a & b; }
static inline int32_t FRACMUL_SHL(int32_t x, int32_t y, int z)
{
  int64_t p = (int64_t)x * y;
  uint32_t lo = (uint32_t)p;
  int32_t hi = (int32_t)(p >> 32);
  return (int32_t)(((uint32_t)hi << (z + 1)) | (lo >> (31 - z)));
}
#endif
