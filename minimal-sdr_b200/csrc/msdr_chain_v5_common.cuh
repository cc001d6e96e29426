// msdr_chain_v5_common.cuh — helpers shared by the row-block kernels (msdr_chain_v5.cu: windows up to 128 words, msdr_chain_v5l.cu: the
// 256-tap window): epilogue drain and demodulation on registers, the biquad tile loop, the developer profile.
#pragma once
#include "msdr_chain_common.cuh"
#include "msdr_tc_common.cuh"

namespace msdr {
namespace v5 {

using namespace tc;

constexpr uint32_t kPad = 0xFFFFFFFFu;
// sleep between mbarrier probes per role (nanoseconds; msdr_device.cuh::mbar_wait).  The tensor-memory hand-off between the MMA warp
// and the epilogue is the kernel's tightest loop and probes often; everything that sits behind a ring of slots can afford to find
// out late that its barrier has flipped, and its probes otherwise delay the dependent instruction chains of the biquad warps.
#ifndef MSDR_V5_NS
#define MSDR_V5_NS 1
#endif
constexpr uint32_t kNsLoad = 256 * MSDR_V5_NS, kNsConv = 128 * MSDR_V5_NS, kNsSlot = 128 * MSDR_V5_NS, kNsBq = 128 * MSDR_V5_NS, kNsStore = 256 * MSDR_V5_NS;

// developer profile (MSDR_PROF=1): per-CTA cycle totals, slot = role * 4 + counter
struct Prof {
  long long *base;
  long long acc[4];
  long long t;
  __device__ __forceinline__ Prof(long long *b, int role) : base(b ? b + (size_t)blockIdx.x * 64 + role * 4 : nullptr), acc{0, 0, 0, 0}, t(0) {}
  __device__ __forceinline__ void start() { if (base) t = clock64(); }
  __device__ __forceinline__ void lap(int i) { if (base) { const long long n = clock64(); acc[i] += n - t; t = n; } }
  __device__ __forceinline__ void flush() { if (base && (threadIdx.x & 31) == 0) for (int i = 0; i < 4; ++i) base[i] = acc[i]; }
};

__device__ __forceinline__ uint4 lds128(uint32_t a)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4 &v)
{
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// NQ * 8 samples of one row in place (NQ even): 128-bit words, two register sets alternating (the next eight samples are on their
// way while the recurrence runs)
template <class BQ, int NQ = N / 8>
__device__ __forceinline__ void bq_tile(BQ (&st)[1], uint32_t a)
{
  uint4 v0 = lds128(a), v1;
#pragma unroll 1
  for (int q = 0; q < NQ; q += 2) {
    v1 = lds128(a + 16u * (uint32_t)(q + 1));
    v0.x = bq_word<1>(st, v0.x);
    v0.y = bq_word<1>(st, v0.y);
    v0.z = bq_word<1>(st, v0.z);
    v0.w = bq_word<1>(st, v0.w);
    sts128(a + 16u * (uint32_t)q, v0);
    if (q + 2 < NQ) v0 = lds128(a + 16u * (uint32_t)(q + 2));
    v1.x = bq_word<1>(st, v1.x);
    v1.y = bq_word<1>(st, v1.y);
    v1.z = bq_word<1>(st, v1.z);
    v1.w = bq_word<1>(st, v1.w);
    sts128(a + 16u * (uint32_t)(q + 1), v1);
  }
}

// one branch of eight output columns: the three byte-plane accumulators -> the reference's accumulator mod 2^32
__device__ __forceinline__ void drain8(uint32_t taddr, uint32_t (&acc)[8])
{
  uint32_t a0[8], a1[8], a2[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]), "=r"(a0[4]), "=r"(a0[5]), "=r"(a0[6]), "=r"(a0[7]) : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]), "=r"(a1[4]), "=r"(a1[5]), "=r"(a1[6]), "=r"(a1[7]) : "r"(taddr + (uint32_t)N));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a2[0]), "=r"(a2[1]), "=r"(a2[2]), "=r"(a2[3]), "=r"(a2[4]), "=r"(a2[5]), "=r"(a2[6]), "=r"(a2[7]) : "r"(taddr + 2u * (uint32_t)N));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (a0[j] << 16) + (a1[j] << 8) + a2[j]; // mod 2^32, like the reference accumulator
}

// demodulation switch (Minimal-SDR.ino:589-628) over the 32 packed (I | Q << 16) words a thread holds -> 16 words of int16 pairs
// SSB kinds straight on the packed words p = I | Q << 16 (Minimal-SDR.ino:591-604: the int16 sum wraps, no saturation):
//   USB  I + Q = upper half of p * 65537;   LSB  I - Q = I + ~Q + 1 = upper half of (p ^ 0xFFFF0000) * 65537 + 0x10000
// one LOP3 and one IMAD per sample, one PRMT per pair.  xm / xc: the per-row XOR mask and addend (0 / 0 for USB).
__device__ __forceinline__ void demod_ssb_regs(const uint32_t (&iq)[32], uint32_t xm, uint32_t xc, uint32_t (&out)[16])
{
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t t0 = (iq[2 * j] ^ xm) * 65537u + xc, t1 = (iq[2 * j + 1] ^ xm) * 65537u + xc;
    out[j] = __byte_perm(t0, t1, 0x7632);
  }
}

template <int KIND>
__device__ __forceinline__ void demod_regs(const uint32_t (&iq)[32], int sgn, uint32_t (&out)[16])
{
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) { // eight samples in flight: the envelope kinds are a dependent-latency problem
    int y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = demod_inline<KIND>((int)(short)(iq[c0 + j] & 0xFFFFu), (int)iq[c0 + j] >> 16, sgn);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[c0 / 2 + j] = ((uint32_t)y[2 * j] & 0xFFFFu) | ((uint32_t)y[2 * j + 1] << 16);
  }
}


} // namespace v5
} // namespace msdr
