/* TEST INFRASTRUCTURE — host shim for compiling the reference's CMSIS-DSP sources on x86-64.
 *
 * Pre-included (gcc -include) ahead of /root/reference/src/CMSIS_5/arm_math.h.
 * arm_math.h:294 hard-defines ARM_MATH_CM4 and :327 pulls "core_cm4.h" (ARM inline asm).
 * We pre-define that header's include guards so it expands to nothing, and supply the
 * Cortex-M4 DSP intrinsics as portable C.  Semantics follow the reference's own portable
 * fall-backs (arm_math.h:492-501 __PKHBT/__PKHTB, :873-921 __SMLAD/__SMLADX/__QADD/__QSUB),
 * which the reference compiles out because ARM_MATH_DSP is forced at arm_math.h:328.
 * Build with -fwrapv: the ARM instructions wrap mod 2^32.
 */
#ifndef MSDR_CMSIS_SHIM_H
#define MSDR_CMSIS_SHIM_H

#include <stdint.h>

/* neutralise core_cm4.h / cmsis_compiler.h */
#define __CORE_CM4_H_GENERIC
#define __CORE_CM4_H_DEPENDANT
#define __CMSIS_COMPILER_H

#define __STATIC_INLINE static inline
#define __STATIC_FORCEINLINE static inline __attribute__((always_inline))
#define __ASM __asm
#define __INLINE inline
#ifndef __FPU_USED
#define __FPU_USED 1
#endif
#define __FPU_PRESENT 1

static inline uint32_t msdr_shim_clz(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }
#define __CLZ(v) msdr_shim_clz((uint32_t)(v))

/* SSAT Rd,#sat,Rn : saturate to signed `sat`-bit range */
static inline int32_t msdr_shim_ssat(int32_t val, uint32_t sat)
{
  if (sat >= 1u && sat <= 32u) {
    const int32_t max = (int32_t)((1u << (sat - 1u)) - 1u);
    const int32_t min = -1 - max;
    if (val > max) return max;
    if (val < min) return min;
  }
  return val;
}
#define __SSAT(v, s) msdr_shim_ssat((int32_t)(v), (uint32_t)(s))

static inline uint32_t msdr_shim_usat(int32_t val, uint32_t sat)
{
  if (sat <= 31u) {
    const uint32_t max = ((1u << sat) - 1u);
    if (val > (int32_t)max) return max;
    if (val < 0) return 0u;
  }
  return (uint32_t)val;
}
#define __USAT(v, s) msdr_shim_usat((int32_t)(v), (uint32_t)(s))

/* PKHBT Rd,Rn,Rm,LSL #s : Rd = Rn[15:0] | (Rm<<s)[31:16]  (arm_math.h:495) */
#define __PKHBT(ARG1, ARG2, ARG3) ((uint32_t)((((uint32_t)(ARG1)) & 0x0000FFFFu) | ((((uint32_t)(ARG2)) << (ARG3)) & 0xFFFF0000u)))
/* PKHTB Rd,Rn,Rm,ASR #s : Rd = Rn[31:16] | (Rm>>s)[15:0]  (arm_math.h:497) */
#define __PKHTB(ARG1, ARG2, ARG3) ((uint32_t)((((uint32_t)(ARG1)) & 0xFFFF0000u) | (((uint32_t)((int32_t)(ARG2) >> (ARG3))) & 0x0000FFFFu)))

static inline int32_t msdr_lo16(uint32_t x) { return (int32_t)(int16_t)(x & 0xFFFFu); }
static inline int32_t msdr_hi16(uint32_t x) { return (int32_t)(int16_t)(x >> 16); }

/* SMLAD: sum + x.lo*y.lo + x.hi*y.hi, 32-bit wrap (arm_math.h:895-906) */
static inline uint32_t __SMLAD(uint32_t x, uint32_t y, uint32_t sum)
{
  return (uint32_t)(msdr_lo16(x) * msdr_lo16(y)) + (uint32_t)(msdr_hi16(x) * msdr_hi16(y)) + sum;
}
/* SMLADX: sum + x.lo*y.hi + x.hi*y.lo (arm_math.h:909-921) */
static inline uint32_t __SMLADX(uint32_t x, uint32_t y, uint32_t sum)
{
  return (uint32_t)(msdr_lo16(x) * msdr_hi16(y)) + (uint32_t)(msdr_hi16(x) * msdr_lo16(y)) + sum;
}
static inline uint32_t __SMUAD(uint32_t x, uint32_t y)  { return __SMLAD(x, y, 0u); }
static inline uint32_t __SMUADX(uint32_t x, uint32_t y) { return __SMLADX(x, y, 0u); }
static inline uint32_t __SMUSD(uint32_t x, uint32_t y)
{
  return (uint32_t)(msdr_lo16(x) * msdr_lo16(y)) - (uint32_t)(msdr_hi16(x) * msdr_hi16(y));
}
static inline uint32_t __SMUSDX(uint32_t x, uint32_t y)
{
  return (uint32_t)(msdr_lo16(x) * msdr_hi16(y)) - (uint32_t)(msdr_hi16(x) * msdr_lo16(y));
}
static inline uint32_t __SMLSDX(uint32_t x, uint32_t y, uint32_t sum) { return __SMUSDX(x, y) + sum; }
static inline uint64_t __SMLALD(uint32_t x, uint32_t y, uint64_t sum)
{
  return sum + (uint64_t)(int64_t)(msdr_lo16(x) * msdr_lo16(y)) + (uint64_t)(int64_t)(msdr_hi16(x) * msdr_hi16(y));
}
static inline uint64_t __SMLALDX(uint32_t x, uint32_t y, uint64_t sum)
{
  return sum + (uint64_t)(int64_t)(msdr_lo16(x) * msdr_hi16(y)) + (uint64_t)(int64_t)(msdr_hi16(x) * msdr_lo16(y));
}
static inline int32_t msdr_shim_sat64(int64_t v)
{
  if (v > 2147483647LL) return 2147483647;
  if (v < -2147483648LL) return (int32_t)0x80000000;
  return (int32_t)v;
}
static inline int32_t __QADD(int32_t x, int32_t y) { return msdr_shim_sat64((int64_t)x + y); }
static inline int32_t __QSUB(int32_t x, int32_t y) { return msdr_shim_sat64((int64_t)x - y); }
static inline uint32_t __QADD16(uint32_t x, uint32_t y)
{
  int32_t r = msdr_shim_ssat(msdr_lo16(x) + msdr_lo16(y), 16);
  int32_t s = msdr_shim_ssat(msdr_hi16(x) + msdr_hi16(y), 16);
  return ((uint32_t)s << 16) | ((uint32_t)r & 0xFFFFu);
}
static inline uint32_t __QSUB16(uint32_t x, uint32_t y)
{
  int32_t r = msdr_shim_ssat(msdr_lo16(x) - msdr_lo16(y), 16);
  int32_t s = msdr_shim_ssat(msdr_hi16(x) - msdr_hi16(y), 16);
  return ((uint32_t)s << 16) | ((uint32_t)r & 0xFFFFu);
}
static inline uint32_t __SHADD16(uint32_t x, uint32_t y)
{
  int32_t r = (msdr_lo16(x) + msdr_lo16(y)) >> 1;
  int32_t s = (msdr_hi16(x) + msdr_hi16(y)) >> 1;
  return ((uint32_t)s << 16) | ((uint32_t)r & 0xFFFFu);
}
static inline uint32_t __SHSUB16(uint32_t x, uint32_t y)
{
  int32_t r = (msdr_lo16(x) - msdr_lo16(y)) >> 1;
  int32_t s = (msdr_hi16(x) - msdr_hi16(y)) >> 1;
  return ((uint32_t)s << 16) | ((uint32_t)r & 0xFFFFu);
}
static inline uint32_t __ROR(uint32_t v, uint32_t s) { s &= 31u; return s ? ((v >> s) | (v << (32u - s))) : v; }
static inline uint32_t __SXTB16(uint32_t x)
{
  return ((uint32_t)(((int32_t)(x << 24) >> 24) & 0x0000FFFF)) | ((uint32_t)(((int32_t)(x << 8) >> 8) & (int32_t)0xFFFF0000));
}
static inline int32_t __SMMLA(int32_t x, int32_t y, int32_t sum)
{
  return sum + (int32_t)(((int64_t)x * y) >> 32);
}

#endif /* MSDR_CMSIS_SHIM_H */
