/* TEST INFRASTRUCTURE — what freq_conv.{h,cpp} needs from arm_math.h, as prototypes: the vendored header is not C++-clean on
 * a 64-bit host (pointer -> int32 casts, arm_math.h:5855).  Same types and signatures as arm_math.h:385-390,1898,2412,2468. */
#ifndef MSDR_STUB_FC_ARM_MATH_H
#define MSDR_STUB_FC_ARM_MATH_H
#include <stdint.h>
typedef int16_t q15_t;
typedef int32_t q31_t;
#ifdef __cplusplus
extern "C" {
#endif
void arm_mult_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize);
void arm_add_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize);
void arm_sub_q15(q15_t *pSrcA, q15_t *pSrcB, q15_t *pDst, uint32_t blockSize);
#ifdef __cplusplus
}
#endif
#endif
