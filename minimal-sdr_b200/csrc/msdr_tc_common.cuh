// msdr_tc_common.cuh — building blocks of the tensor-core FIR (tcgen05.mma kind::i8), shared by the stand-alone FIR + demod
// kernel (msdr_fir_tc.cu) and the fused chain kernel (msdr_chain_v4.cu).  See msdr_fir_tc.cu for the derivation.
#pragma once
#include "msdr_device.cuh"
#include "msdr_internal.h"

namespace msdr {
namespace tc {

constexpr int M = 128;        // channels per tile = TMEM lanes
constexpr int P = 32;         // output pairs (= window words) per tile
constexpr int N = 2 * P;      // GEMM N = output samples per tile
constexpr int RING_MAX = 8;   // A ring depth in pairs of K-blocks (a pair = 32 window words per row)
constexpr int OW = N + 4;     // staging row pitch, words (packed I/Q per sample, then the demodulated words in place)
constexpr uint32_t kPairBytes = 2 * (M / 8) * 128; // one pair of one plane: 2 K-blocks x 16 row-groups x 128 B = 4 KB
constexpr uint32_t kKStrideA = (M / 8) * 128;      // K-adjacent core matrices of A
constexpr uint32_t kKStrideB = (N / 8) * 128;
constexpr uint32_t kStagingBytes = M * OW * 4;

__host__ __device__ inline uint32_t a_plane_bytes(uint32_t ring) { return ring * kPairBytes; }
// deepest ring (<= RING_MAX) the window allows; the converters run ring - K/32 pairs ahead of the tensor core
__host__ __device__ inline uint32_t ring_for(uint32_t K, uint32_t max_ring) { return max_ring < (uint32_t)RING_MAX ? max_ring : (uint32_t)RING_MAX; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; // descriptor version (sm_100)
  return d;
}
// The MMA warp runs its loop warp-uniformly and ONE elected lane issues (elect.sync inside the asm): descriptors and TMEM
// addresses are then warp-uniform values the compiler keeps in uniform registers.  Issued from inside `if (lane == 0)` every
// operand goes through a R2UR broadcast loop (~100 cycles of single-thread latency per MMA).
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_512(uint32_t *slot)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t tmem)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 window words (32 samples) of one row -> the four byte planes of ring position `pos`, K-block `kb`.
// v[0..3]: the raw words; odd words are negated here (fs/4 mix, Minimal-SDR.ino:550,555).
__device__ __forceinline__ void convert_store(uint8_t *sA, uint32_t plane_bytes, uint32_t pos, uint32_t kb, uint32_t r, const uint4 (&v)[4])
{
  uint32_t el[4], eh[4], ol[4], oh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = v[j].x, b = neg16x2(v[j].y), c = v[j].z, d = neg16x2(v[j].w);
    el[j] = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
    eh[j] = __byte_perm(__byte_perm(a, b, 0x0051), __byte_perm(c, d, 0x0051), 0x5410);
    ol[j] = __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
    oh[j] = __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
  }
  const uint32_t off = ((pos * 2 + kb) * (M / 8) + r / 8) * 128 + (r % 8) * 16;
  *reinterpret_cast<uint4 *>(sA + 0 * plane_bytes + off) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
  *reinterpret_cast<uint4 *>(sA + 1 * plane_bytes + off) = make_uint4(el[0], el[1], el[2], el[3]);
  *reinterpret_cast<uint4 *>(sA + 2 * plane_bytes + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  *reinterpret_cast<uint4 *>(sA + 3 * plane_bytes + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

// All MMAs of one tile (one thread): 2 branches x 4 plane combinations x KS steps of M128 N64 K32.
// qt = ring sequence number of the NEWEST pair of the window; accumulator (branch, combo) lives at TMEM column (4 br + combo) * N.
// The descriptors differ only in their start-address field, so the plane bases are encoded once and offsets are added.
struct IssueCtx {
  uint64_t descA[4], descB[4]; // plane bases: A e_hi, e_lo, o_hi, o_lo;  B I_hi, I_lo, Q_hi, Q_lo
};
__device__ __forceinline__ void issue_init(IssueCtx &c, const uint8_t *sA, uint32_t a_plane, const uint8_t *sB, uint32_t b_plane)
{
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c.descA[i] = make_desc(smem_u32(sA + i * a_plane), kKStrideA, 128);
    c.descB[i] = make_desc(smem_u32(sB + i * b_plane), kKStrideB, 128);
  }
}
// Accumulators per branch: 0 = hi*hi, 1 = hi*lo + lo*hi (two MMAs into the same accumulator), 2 = lo*lo.
constexpr int kAccPerBranch = 3;
__device__ __forceinline__ void umma_i8(uint32_t dcol, uint64_t da, uint64_t db, uint32_t a_signed, uint32_t b_signed, uint32_t acc)
{
  const uint32_t idesc = (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void issue_tile(const IssueCtx &c, uint32_t tmem, uint32_t qt, uint32_t KS, uint32_t ring)
{
  uint32_t pos = (qt - (KS - 1)) % ring;
  for (uint32_t ks = 0; ks < KS; ++ks) {
    const uint64_t aoff = (uint64_t)((pos * kPairBytes) >> 4), boff = (uint64_t)((ks * 2 * kKStrideB) >> 4);
    const uint32_t acc = ks > 0;
#pragma unroll
    for (uint32_t br = 0; br < 2; ++br) {
      const uint64_t a_hi = c.descA[2 * br] + aoff, a_lo = c.descA[2 * br + 1] + aoff; // hi planes are signed bytes, lo planes unsigned
      const uint64_t b_hi = c.descB[2 * br] + boff, b_lo = c.descB[2 * br + 1] + boff;
      const uint32_t d0 = tmem + (br * kAccPerBranch) * N;
      umma_i8(d0, a_hi, b_hi, 1, 1, acc);
      umma_i8(d0 + N, a_hi, b_lo, 1, 0, acc);
      umma_i8(d0 + N, a_lo, b_hi, 0, 1, 1);
      umma_i8(d0 + 2 * N, a_lo, b_lo, 0, 0, acc);
    }
    pos = pos + 1 == ring ? 0 : pos + 1;
  }
}

// The MMAs of ONE branch of a tile (msdr_chain_v5.cu hands the two branches to the epilogue separately).
__device__ __forceinline__ void issue_branch(const IssueCtx &c, uint32_t tmem, uint32_t qt, uint32_t KS, uint32_t ring, uint32_t br)
{
  uint32_t pos = (qt - (KS - 1)) % ring;
  const uint32_t d0 = tmem + (br * kAccPerBranch) * N;
  for (uint32_t ks = 0; ks < KS; ++ks) {
    const uint64_t aoff = (uint64_t)((pos * kPairBytes) >> 4), boff = (uint64_t)((ks * 2 * kKStrideB) >> 4);
    const uint32_t acc = ks > 0;
    const uint64_t a_hi = c.descA[2 * br] + aoff, a_lo = c.descA[2 * br + 1] + aoff; // hi planes are signed bytes, lo planes unsigned
    const uint64_t b_hi = c.descB[2 * br] + boff, b_lo = c.descB[2 * br + 1] + boff;
    umma_i8(d0, a_hi, b_hi, 1, 1, acc);
    umma_i8(d0 + N, a_hi, b_lo, 1, 0, acc);
    umma_i8(d0 + N, a_lo, b_hi, 0, 1, 1);
    umma_i8(d0 + 2 * N, a_lo, b_lo, 0, 0, acc);
    pos = pos + 1 == ring ? 0 : pos + 1;
  }
}

// Epilogue step 1 (thread = TMEM lane = channel row): accumulators -> recombine -> >>15 -> SSAT16 (arm_fir_fast_q15.c:234-238)
// -> packed (I | Q << 16) parked in this thread's staging row.  cvt.pack.sat saturates and packs both branches at once.
__device__ __forceinline__ uint32_t pack_sat_iq(int I, int Q)
{
  uint32_t d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(Q), "r"(I)); // upper half <- sat16(Q), lower half <- sat16(I)
  return d;
}
__device__ __forceinline__ void drain_tile(uint32_t lane_addr, uint32_t *orow)
{
#ifdef MSDR_DRAIN_UNROLLED
#pragma unroll
#else
#pragma unroll 1 // eight passes of eight columns as a loop: 650 instructions of straight-line code per tile and warp otherwise (instruction caches)
#endif
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t acc[2 * kAccPerBranch][8]; // [branch*3 + part][column]
#pragma unroll
    for (int a = 0; a < 2 * kAccPerBranch; ++a) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(acc[a][0]), "=r"(acc[a][1]), "=r"(acc[a][2]), "=r"(acc[a][3]), "=r"(acc[a][4]), "=r"(acc[a][5]), "=r"(acc[a][6]), "=r"(acc[a][7])
                   : "r"(lane_addr + (uint32_t)(a * N + c0)));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t iq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t ai = (acc[0][j] << 16) + (acc[1][j] << 8) + acc[2][j]; // mod 2^32, like the reference accumulator
      const uint32_t aq = (acc[3][j] << 16) + (acc[4][j] << 8) + acc[5][j];
      iq[j] = pack_sat_iq((int)ai >> 15, (int)aq >> 15);
    }
    *reinterpret_cast<uint4 *>(orow + c0) = make_uint4(iq[0], iq[1], iq[2], iq[3]);
    *reinterpret_cast<uint4 *>(orow + c0 + 4) = make_uint4(iq[4], iq[5], iq[6], iq[7]);
  }
}
// Epilogue step 2: demodulation switch (Minimal-SDR.ino:589-628) over the parked row, in place (word c/2 <= c is already consumed).
// One thread owns the row, so the square roots of the envelope kinds are a dependent-latency problem: eight samples are kept in
// flight per batch (the per-sample call of msdr_device.cuh::demod_envelope left the lone epilogue warp of a sub-partition
// waiting ~250 cycles per sample).
template <int KIND>
__device__ __forceinline__ int demod_inline(int I, int Q, int sgn)
{
  if (KIND <= 1) return (int)(short)(I + sgn * Q); // kind 0 (LSB): I - Q, kind 1 (USB): I + Q
  const int s = (int)((uint32_t)(I * I) + (uint32_t)(Q * Q));
  if (KIND == 2) {
    // arm_sqrt_f32 (arm_math.h:5733-5760): sqrtf for in >= 0, else 0.  s < 0 only for I = Q = -32768 (s wraps to -2^31).
    // Computed unconditionally and selected at the end: a branch per sample would serialise the eight square roots in flight.
    // (the select is an AND with an arithmetic mask: a ?: lets the compiler sink the whole chain back under a branch)
    const float r = sqrt_rn_fast(__uint2float_rn((uint32_t)s));
    const int y = __float2int_rz(r);
    const int positive = ((-s) & ~s) >> 31; // -1 for s > 0, else 0
    return (int)(short)(y & positive);
  }
  return (int)(short)(sqrt_q31(s, nullptr) >> 16);
}
template <int KIND>
__device__ __forceinline__ void demod_row_kind(uint32_t *orow, int sgn)
{
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 8) {
    const uint4 v0 = *reinterpret_cast<const uint4 *>(orow + c0), v1 = *reinterpret_cast<const uint4 *>(orow + c0 + 4);
    const uint32_t iq[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    int y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = demod_inline<KIND>((int)(short)(iq[j] & 0xFFFFu), (int)iq[j] >> 16, sgn);
    *reinterpret_cast<uint4 *>(orow + (c0 >> 1)) =
        make_uint4(((uint32_t)y[0] & 0xFFFFu) | ((uint32_t)y[1] << 16), ((uint32_t)y[2] & 0xFFFFu) | ((uint32_t)y[3] << 16),
                   ((uint32_t)y[4] & 0xFFFFu) | ((uint32_t)y[5] << 16), ((uint32_t)y[6] & 0xFFFFu) | ((uint32_t)y[7] << 16));
  }
}
__device__ __forceinline__ void demod_row(uint32_t *orow, int kind)
{
  if (kind <= 1) demod_row_kind<0>(orow, kind ? 1 : -1);
  else if (kind == 2) demod_row_kind<2>(orow, 0);
  else demod_row_kind<3>(orow, 0);
}

} // namespace tc
} // namespace msdr
