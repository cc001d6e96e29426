// msdr_fir_tc.cu — K3 (study): fs/4 mix + FIR pair + demodulation on the 5th-generation tensor cores (tcgen05.mma kind::i8).
//
// Exactness licence: arm_fir_fast_q15 accumulates mod 2^32 (arm_fir_fast_q15.c:132-183), so the accumulator may be evaluated in
// any order and in any decomposition.  int16 = 256 * hi (signed byte) + lo (unsigned byte) on both the samples and the taps:
//   sum c*x = 2^16 * sum ch*xh + 2^8 * (sum ch*xl + sum cl*xh) + sum cl*xl            (mod 2^32)
// Four int8 GEMMs with exact int32 accumulation (|partial| <= K * 2^16 < 2^24) run on the tensor cores with independently
// signed/unsigned operands (verified on hardware by tools/tc_probe/umma_i8_probe.cu); the recombination, >>15, SSAT16 and
// the demodulation switch (Minimal-SDR.ino:589-628) are done literally on the CUDA cores.
//
// Toeplitz form per tile of 128 channels x 64 output samples (polyphase split of DESIGN.md 4.1: I uses the even samples, Q the
// odd ones, P = 32 output pairs):
//   D_b[row = channel, n = 2i + p] = sum_k  A_b[row, k] * B_b[n, k],   b in {I, Q},  k = window word index, K = roundup32(P + KP - 1)
//   A_I[row, k] = byte plane of the sign-folded EVEN sample of word k of the row's window,  A_Q: the ODD sample
//   B_I[2i,   k] = cB[d]   B_I[2i+1, k] = cA[d]   B_Q[2i, k] = cD[d]   B_Q[2i+1, k] = cC[d],   d = k - i + KP - 1 - K + P  (0 if outside [0, KP))
// with cA..cD the expanded sub-filters of msdr_capi.cu::expand_set.  All rows of a tile share one tap table (B is the shared
// operand); the host orders rows by table.
//
// Shared-memory operands are K-major, no swizzle: core matrices of 8 rows x 16 bytes, ordered [k/16][row/8][row%8][k%16];
// descriptor LBO = stride between K-adjacent core matrices, SBO = 128 (row-group stride).
//
// This first version runs the phases of a tile back to back (convert -> MMA -> epilogue); it exists to establish parity
// and the cost structure on hardware.  See DESIGN.md "K3" for measurements and what a pipelined version would add.
#include "msdr_device.cuh"
#include "msdr_internal.h"

namespace msdr {
namespace tc {

constexpr int M = 128;       // channels per tile = TMEM lanes
constexpr int P = 32;        // output pairs per tile
constexpr int N = 2 * P;     // GEMM N = output samples per tile
constexpr int kThreads = 256;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; // descriptor version (sm_100)
  return d;
}

struct TcParams {
  const int16_t *in;     // [rows][stride]; every row has `pad` valid (history or zero) samples BEFORE index 0
  int16_t *out;          // [rows][ostride] demodulated int16
  size_t stride, ostride;
  uint32_t rows;         // multiple of M is NOT required
  uint32_t L;            // samples per row, multiple of 64
  uint32_t K;            // window words per tile, multiple of 32
  const uint8_t *bmat;   // [n_sets][4][N * K] bytes: B_I hi, B_I lo, B_Q hi, B_Q lo in canonical layout
  const uint8_t *row_set; // [rows] table id (uniform inside each block of M rows)
  const uint8_t *row_kind; // [rows] demod kind 0..3
  uint32_t n_row_blocks, n_time_tiles;
};

__global__ void __launch_bounds__(kThreads, 1) fir_demod_tc_kernel(const TcParams p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t K = p.K;
  const uint32_t a_plane = M * K;            // bytes per A plane
  const uint32_t b_plane = N * K;            // bytes per B plane
  uint8_t *sA = smem;                        // 4 planes: e_hi, e_lo, o_hi, o_lo
  uint8_t *sB = smem + 4 * a_plane;          // 4 planes: I_hi, I_lo, Q_hi, Q_lo
  uint32_t *sOut = reinterpret_cast<uint32_t *>(sB + 4 * b_plane); // [M][N/2 + 4] words
  constexpr int OW = N / 2 + 4;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint8_t s_kind[M];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const uint32_t kstrideA = (M / 8) * 128, kstrideB = (N / 8) * 128;
  uint32_t phase = 0;
  int cur_set = -1;

  const uint32_t n_tiles = p.n_row_blocks * p.n_time_tiles;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t rb = tile / p.n_time_tiles, tt = tile - rb * p.n_time_tiles;
    const uint32_t row0 = rb * M;
    const int set = p.row_set[row0];
    if (set != cur_set) { // (re)load the Toeplitz operand of this table
      const uint4 *src = reinterpret_cast<const uint4 *>(p.bmat + (size_t)set * 4 * b_plane);
      uint4 *dst = reinterpret_cast<uint4 *>(sB);
      for (uint32_t i = tid; i < 4 * b_plane / 16; i += kThreads) dst[i] = src[i];
      cur_set = set;
    }
    if (tid < M) s_kind[tid] = (row0 + tid < p.rows) ? p.row_kind[row0 + tid] : 0;

    // ---- convert: raw int16 window -> sign-folded byte planes in UMMA layout.  Task = (row, 16 consecutive window words).
    // window word k of the tile <-> samples 2*(tt*P + P - K + k) + {0,1} of the row (negative indices reach into the padding)
    const long long w_first = (long long)tt * P + P - (long long)K;
    for (uint32_t task = tid; task < M * (K / 16); task += kThreads) {
      const uint32_t r = task % M, kg = task / M; // consecutive threads -> consecutive rows: conflict-free 16-byte stores
      uint32_t w[16];
      if (row0 + r < p.rows) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.in + (size_t)(row0 + r) * p.stride + 2 * (w_first + (long long)kg * 16));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 v = __ldg(src + q);
          w[4 * q + 0] = v.x; w[4 * q + 1] = neg16x2(v.y); w[4 * q + 2] = v.z; w[4 * q + 3] = neg16x2(v.w); // odd words negated (fs/4 mix)
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = 0u;
      }
      uint32_t el[4], eh[4], ol[4], oh[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[4 * q], b = w[4 * q + 1], c = w[4 * q + 2], d = w[4 * q + 3];
        // byte j of each word gathered across the four words
        el[q] = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
        eh[q] = __byte_perm(__byte_perm(a, b, 0x0051), __byte_perm(c, d, 0x0051), 0x5410);
        ol[q] = __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
        oh[q] = __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
      }
      const uint32_t off = (kg * (M / 8) + r / 8) * 128 + (r % 8) * 16;
      *reinterpret_cast<uint4 *>(sA + 0 * a_plane + off) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
      *reinterpret_cast<uint4 *>(sA + 1 * a_plane + off) = make_uint4(el[0], el[1], el[2], el[3]);
      *reinterpret_cast<uint4 *>(sA + 2 * a_plane + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4 *>(sA + 3 * a_plane + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
    fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- MMA: 2 branches x 4 byte-plane combinations x K/32 steps, one thread issues
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (uint32_t br = 0; br < 2; ++br) {
        for (uint32_t combo = 0; combo < 4; ++combo) {
          const uint32_t ah = combo >> 1, bh = combo & 1; // 0 = hi plane (signed), 1 = lo plane (unsigned)
          const uint32_t idesc = (2u << 4) | ((ah == 0 ? 1u : 0u) << 7) | ((bh == 0 ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
          const uint32_t abase = smem_u32(sA + (2 * br + ah) * a_plane), bbase = smem_u32(sB + (2 * br + bh) * b_plane);
          const uint32_t dcol = tmem + (br * 4 + combo) * N;
          for (uint32_t ks = 0; ks < K / 32; ++ks) {
            const uint64_t da = make_desc(abase + ks * 2 * kstrideA, kstrideA, 128);
            const uint64_t db = make_desc(bbase + ks * 2 * kstrideB, kstrideB, 128);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: thread = channel row (TMEM lane); recombine the byte planes, >>15, SSAT16, demodulate
    if (warp < 4) {
      const int kind = s_kind[tid];
      const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t *orow = sOut + (uint32_t)tid * OW;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t acc[8][8]; // [branch*4 + combo][column]
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(acc[a][0]), "=r"(acc[a][1]), "=r"(acc[a][2]), "=r"(acc[a][3]), "=r"(acc[a][4]), "=r"(acc[a][5]), "=r"(acc[a][6]),
                         "=r"(acc[a][7])
                       : "r"(lane_addr + (uint32_t)(a * N + c0)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        int y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // combos: 0 = hi*hi, 1 = hi(x)*lo(c), 2 = lo(x)*hi(c), 3 = lo*lo
          const uint32_t ai = (acc[0][j] << 16) + ((acc[1][j] + acc[2][j]) << 8) + acc[3][j];
          const uint32_t aq = (acc[4][j] << 16) + ((acc[5][j] + acc[6][j]) << 8) + acc[7][j];
          const int I = ssat16((int)ai >> 15), Q = ssat16((int)aq >> 15); // arm_fir_fast_q15.c:234-238
          y[j] = demod_sample(kind, I, Q);
        }
        uint4 o;
        o.x = ((uint32_t)y[0] & 0xFFFFu) | ((uint32_t)y[1] << 16);
        o.y = ((uint32_t)y[2] & 0xFFFFu) | ((uint32_t)y[3] << 16);
        o.z = ((uint32_t)y[4] & 0xFFFFu) | ((uint32_t)y[5] << 16);
        o.w = ((uint32_t)y[6] & 0xFFFFu) | ((uint32_t)y[7] << 16);
        *reinterpret_cast<uint4 *>(orow + c0 / 2) = o;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- coalesced write-back: 128 rows x 128 bytes
    for (uint32_t i = tid; i < M * (N / 8); i += kThreads) {
      const uint32_t r = i / (N / 8), q = i % (N / 8);
      if (row0 + r < p.rows)
        *reinterpret_cast<uint4 *>(p.out + (size_t)(row0 + r) * p.ostride + (size_t)tt * N + q * 8) = *reinterpret_cast<const uint4 *>(sOut + r * OW + q * 4);
    }
    __syncthreads();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

} // namespace tc

uint32_t tc_window_words(uint32_t T) { return ((tc::P + kp_of_taps(T) - 1u) + 31u) & ~31u; }
uint32_t tc_tile_samples() { return tc::N; }
uint32_t tc_tile_rows() { return tc::M; }

// Toeplitz operand of one table for the tensor-core FIR, in the canonical UMMA layout (see header comment).
// cA..cD: expanded sub-filters (length KP each).  out: 4 planes x N*K bytes.
void tc_build_bmat(const int *cA, const int *cB, const int *cC, const int *cD, uint32_t KP, uint32_t K, uint8_t *out)
{
  const uint32_t N = tc::N, P = tc::P, plane = N * K;
  for (uint32_t n = 0; n < N; ++n) {
    const uint32_t i = n >> 1, odd = n & 1;
    for (uint32_t k = 0; k < K; ++k) {
      const long d = (long)k - (long)i + (long)KP - 1 - (long)K + (long)P;
      int ci = 0, cq = 0;
      if (d >= 0 && d < (long)KP) { ci = odd ? cA[d] : cB[d]; cq = odd ? cC[d] : cD[d]; }
      const uint32_t off = ((k / 16) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 16);
      out[0 * plane + off] = (uint8_t)((ci >> 8) & 0xFF); // hi byte (two's complement => signed plane)
      out[1 * plane + off] = (uint8_t)(ci & 0xFF);        // lo byte (unsigned plane)
      out[2 * plane + off] = (uint8_t)((cq >> 8) & 0xFF);
      out[3 * plane + off] = (uint8_t)(cq & 0xFF);
    }
  }
}

cudaError_t launch_fir_demod_tc(const int16_t *in, size_t stride, int16_t *out, size_t ostride, uint32_t rows, uint32_t L, uint32_t K, const uint8_t *bmat,
                                const uint8_t *row_set, const uint8_t *row_kind, cudaStream_t s)
{
  using namespace tc;
  TcParams p{};
  p.in = in; p.out = out; p.stride = stride; p.ostride = ostride; p.rows = rows; p.L = L; p.K = K;
  p.bmat = bmat; p.row_set = row_set; p.row_kind = row_kind;
  p.n_row_blocks = (rows + M - 1) / M;
  p.n_time_tiles = L / N;
  if (p.n_row_blocks == 0 || p.n_time_tiles == 0) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const size_t smem = (size_t)4 * M * K + (size_t)4 * N * K + (size_t)M * (N / 2 + 4) * 4 + 1024;
  e = cudaFuncSetAttribute(fir_demod_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  uint32_t grid = (uint32_t)sms;
  const uint32_t n_tiles = p.n_row_blocks * p.n_time_tiles;
  if (grid > n_tiles) grid = n_tiles;
  fir_demod_tc_kernel<<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}

} // namespace msdr
