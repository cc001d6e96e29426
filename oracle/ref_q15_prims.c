/* TEST INFRASTRUCTURE — arm_mult_q15 / arm_add_q15 / arm_sub_q15 for the reference's freq_conv.cpp.
 *
 * These three CMSIS-DSP V1.5.1 routines are declared by the reference (arm_math.h:1898,2412,2468) but their sources are NOT
 * vendored (on a Teensy they come from the core's pre-built libarm_cortexM4l_math.a).  What IS vendored are the saturating
 * primitives they are made of: the portable C definitions of __QADD16 / __QSUB16 (arm_math.h:721-765) and clip_q31_to_q15
 * (arm_math.h:555-560).  oracle/Makefile extracts those definitions by line range into _ref/q15_prims_extract.inc at build
 * time (nothing is copied into the repo); the loops around them follow the documented element-wise form of the CMSIS routines
 * ("saturating add/sub of q15", "(a*b)>>15 saturated to q15").  So for row A6 the saturation arithmetic is the reference's own
 * code and only the three one-line loops are restated. */
#include <stdint.h>
typedef int16_t q15_t;
typedef int32_t q31_t;
typedef int64_t q63_t;
#define CMSIS_INLINE
#define __STATIC_INLINE static inline
static inline int32_t prim_ssat(int32_t val, uint32_t sat)
{ /* SSAT Rd, #sat, Rn (ARMv7-M): clamp to the signed sat-bit range */
  const int32_t max = (int32_t)((1u << (sat - 1u)) - 1u), min = -1 - max;
  return val > max ? max : val < min ? min : val;
}
#define __SSAT(v, s) prim_ssat((int32_t)(v), (uint32_t)(s))
#include "q15_prims_extract.inc" /* clip_q31_to_q15, __QADD16, __QSUB16 from the reference's arm_math.h */

uint32_t ref_qadd16(uint32_t x, uint32_t y) { return __QADD16(x, y); }
uint32_t ref_qsub16(uint32_t x, uint32_t y) { return __QSUB16(x, y); }
int16_t ref_clip_q31_to_q15(int32_t x) { return clip_q31_to_q15(x); }

void arm_mult_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize)
{
  for (uint32_t i = 0; i < blockSize; i++) pDst[i] = clip_q31_to_q15(((q31_t)pSrcA[i] * pSrcB[i]) >> 15);
}
void arm_add_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize)
{
  for (uint32_t i = 0; i < blockSize; i++) pDst[i] = (q15_t)__QADD16((uint32_t)(uint16_t)pSrcA[i], (uint32_t)(uint16_t)pSrcB[i]);
}
void arm_sub_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize)
{
  for (uint32_t i = 0; i < blockSize; i++) pDst[i] = (q15_t)__QSUB16((uint32_t)(uint16_t)pSrcA[i], (uint32_t)(uint16_t)pSrcB[i]);
}
