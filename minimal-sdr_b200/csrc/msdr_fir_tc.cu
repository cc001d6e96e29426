// msdr_fir_tc.cu — K3: fs/4 mix + FIR pair + demodulation on the 5th-generation tensor cores (tcgen05.mma kind::i8).
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
// Pipeline (one persistent CTA per SM; a work item = 128 rows x a run of consecutive 64-sample tiles):
//   convert warps   raw int16 -> sign-folded byte planes, each window word converted ONCE into a ring of 8 "pairs" (32 words =
//                   two K-blocks of 16) per plane; the window of a tile is the last K/32 pairs, addressed by descriptors
//   MMA warp        one thread: 2 branches x 4 plane combinations x K/32 steps of M128 N64 K32; tcgen05.commit signals the
//                   epilogue (accumulators complete) and the converters (oldest pair no longer read)
//   epilogue warps  4 warps = 128 TMEM lanes = 128 rows: tcgen05.ld, recombine, >>15, SSAT16 -> packed I/Q in registers,
//                   release TMEM (the next tile's MMAs overlap the demodulation), demodulate, coalesced store
#include "msdr_tc_common.cuh"

namespace msdr {
namespace tc {

constexpr int RING = RING_MAX;
constexpr int kThreads = 512; // 16 warps; roles by warp id (warp id % 4 = SM sub-partition)
// warps 0-3: epilogue (TMEM lane quadrant = warp id % 4);  warp 6: MMA;  warps 7,10,11,14,15: convert;  4,5,8,9,12,13: unused here
constexpr int NCONV = 5;
constexpr int kLive = (4 + 1 + NCONV) * 32;
constexpr uint32_t kPlaneBytes = RING * kPairBytes; // one A plane ring: 32 KB

__device__ __forceinline__ bool is_conv_warp(int w) { return w == 7 || w == 10 || w == 11 || w == 14 || w == 15; }
__device__ __forceinline__ int conv_index(int w) { return w == 7 ? 0 : w == 10 ? 1 : w == 11 ? 2 : w == 14 ? 3 : 4; }

struct TcParams {
  const int16_t *in;      // [rows][stride]; every row has 2*(K-P) valid (history or zero) samples BEFORE index 0
  int16_t *out;           // [rows][ostride] demodulated int16
  size_t stride, ostride;
  uint32_t rows;
  uint32_t L;             // samples per row, multiple of 64
  uint32_t K;             // window words per tile, multiple of 32, <= 192
  const uint8_t *bmat;    // [n_sets][4][N * K] bytes: B_I hi, B_I lo, B_Q hi, B_Q lo in canonical layout
  const uint8_t *row_set; // [rows] table id (uniform inside each block of M rows)
  const uint8_t *row_kind; // [rows] demod kind 0..3
  uint32_t n_row_blocks, n_time_tiles, tiles_per_item, items_per_block;
  int *counter;           // work-item counter (zeroed before launch)
};

struct __align__(16) TcCtrl {
  uint64_t a_full[RING];   // convert -> MMA   : pair converted (NCONV arrivals)
  uint64_t blk_free[RING]; // MMA -> convert   : pair no longer read (tcgen05.commit)
  uint64_t tmem_full;      // MMA -> epilogue  : accumulators complete (tcgen05.commit)
  uint64_t tmem_empty;     // epilogue -> MMA  : accumulators drained (4 arrivals)
  uint32_t tmem_base;
  int item[2];             // work item of the even / odd iteration (-1: done)
};

__global__ void __launch_bounds__(kThreads, 1) fir_demod_tc_kernel(const TcParams p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t K = p.K, KS = K / 32; // K-steps per tile = pairs per window
  uint8_t *sA = smem;                                   // 4 plane rings: e_hi, e_lo, o_hi, o_lo
  uint8_t *sB = smem + 4 * kPlaneBytes;                 // 4 planes: I_hi, I_lo, Q_hi, Q_lo
  const uint32_t b_plane = N * K;
  uint32_t *sOut = reinterpret_cast<uint32_t *>(sB + 4 * b_plane); // [M][OW] words
  TcCtrl *tc = reinterpret_cast<TcCtrl *>(reinterpret_cast<uint8_t *>(sOut) + M * OW * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&tc->a_full[i], NCONV); mbar_init(&tc->blk_free[i], 1); }
    mbar_init(&tc->tmem_full, 1);
    mbar_init(&tc->tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc_512(&tc->tmem_base);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tc->tmem_base;
  const bool is_epi = warp < 4, is_mma = warp == 6, is_conv = is_conv_warp(warp);
  if (!(is_epi || is_mma || is_conv)) return;

  IssueCtx ictx;
  issue_init(ictx, sA, kPlaneBytes, sB, b_plane);
  uint32_t q = 0;     // pairs converted so far by this CTA (ring position q % RING, phase q / RING)
  uint32_t ntile = 0; // tiles processed so far by this CTA (TMEM hand-off phases)
  int cur_set = -1;
  uint32_t iter = 0;
  const uint32_t n_items = p.n_row_blocks * p.items_per_block;

  for (;; ++iter) {
    // ---- work item: a row block and a run of consecutive tiles (time-major across row blocks)
    if (is_mma && lane == 0) {
      const int it = atomicAdd(p.counter, 1);
      tc->item[iter & 1] = it < (int)n_items ? it : -1;
    }
    named_bar_sync(1, kLive); // item visible; every role has finished the previous item
    const int item = tc->item[iter & 1];
    if (item < 0) break;
    const uint32_t seg = (uint32_t)item / p.n_row_blocks, rb = (uint32_t)item - seg * p.n_row_blocks;
    const uint32_t tb = seg * p.tiles_per_item, te = min(tb + p.tiles_per_item, p.n_time_tiles);
    const uint32_t row0 = rb * M;
    const uint32_t qbase = q;                 // first (warm-up) pair of this item
    const uint32_t npairs = (te - tb) + KS - 1;

    if (is_conv) {
      const int ctid = conv_index(warp) * 32 + lane;
      const int set = p.row_set[row0];
      if (set != cur_set) { // (re)load the Toeplitz operand of this table; the pipeline is drained at item boundaries
        const uint4 *src = reinterpret_cast<const uint4 *>(p.bmat + (size_t)set * 4 * b_plane);
        uint4 *dst = reinterpret_cast<uint4 *>(sB);
        for (uint32_t i = ctid; i < 4 * b_plane / 16; i += NCONV * 32) dst[i] = __ldg(src + i);
        cur_set = set;
      }
      // pair u covers window words [32 u', 32 u' + 32) of every row, u' = tb - (KS - 1) + u  (negative: the padding in front)
      for (uint32_t u = 0; u < npairs; ++u, ++q) {
        const uint32_t pos = q % RING;
        mbar_wait(&tc->blk_free[pos], ((q / RING) & 1u) ^ 1u);
        const long long w0 = ((long long)tb - (long long)(KS - 1) + (long long)u) * P; // first word of the pair
        for (uint32_t task = ctid; task < M * 2; task += NCONV * 32) {
          const uint32_t r = task % M, kb = task / M; // consecutive threads -> consecutive rows: conflict-free 16-byte stores
          uint4 v[4];
          if (row0 + r < p.rows) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.in + (size_t)(row0 + r) * p.stride + 2 * (w0 + (long long)kb * 16));
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __ldg(src + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = make_uint4(0, 0, 0, 0);
          }
          convert_store(sA, kPlaneBytes, pos, kb, r, v);
        }
        fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&tc->a_full[pos]);
      }
    } else if (is_mma) {
      q += npairs;
      // warp-uniform: every lane waits, one elected lane issues (see umma_i8)
      for (uint32_t t = tb; t < te; ++t, ++ntile) {
        const uint32_t qt = qbase + (t - tb) + KS - 1; // newest pair of this tile's window
        mbar_wait(&tc->a_full[qt % RING], (qt / RING) & 1u);
        mbar_wait(&tc->tmem_empty, (ntile & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_tile(ictx, tmem, qt, KS, RING);
        umma_commit(&tc->tmem_full);
        umma_commit(&tc->blk_free[(qt - (KS - 1)) % RING]); // the oldest pair of the window is not read again
      }
      // the remaining pairs of the last window are free once the last MMAs have completed
      for (uint32_t s = 1; s < KS; ++s) umma_commit(&tc->blk_free[(qbase + npairs - KS + s) % RING]);
      __syncwarp();
    } else { // epilogue: thread = channel row (TMEM lane)
      q += npairs;
      const int kind = (row0 + tid < p.rows) ? p.row_kind[row0 + tid] : 0;
      const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
      uint32_t *orow = sOut + (uint32_t)tid * OW;
      for (uint32_t t = tb; t < te; ++t, ++ntile) {
        mbar_wait(&tc->tmem_full, ntile & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        named_bar_sync(2, 128); // the previous tile's staging rows have been copied out
        drain_tile(lane_addr, orow);
        // accumulators are out of TMEM: hand it back so the next tile's MMAs overlap the demodulation
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tc->tmem_empty);
        demod_row(orow, kind);
        named_bar_sync(2, 128);
        // coalesced write-back: 128 rows x 128 bytes
        for (uint32_t i = tid; i < M * (N / 8); i += 128) {
          const uint32_t r = i / (N / 8), c = i % (N / 8);
          if (row0 + r < p.rows)
            *reinterpret_cast<uint4 *>(p.out + (size_t)(row0 + r) * p.ostride + (size_t)t * N + c * 8) = *reinterpret_cast<const uint4 *>(sOut + r * OW + c * 4);
        }
      }
    }
  }

  // everybody who is still here has left the loop through the barrier above; MMAs are complete (the epilogue consumed them)
  if (warp == 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tmem_dealloc_512(tmem);
  }
}

} // namespace tc

uint32_t tc_window_words_kp(uint32_t KP) { return ((tc::P + KP - 1u) + 31u) & ~31u; }
uint32_t tc_window_words(uint32_t T) { return tc_window_words_kp(kp_of_taps(T)); }
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
                                const uint8_t *row_set, const uint8_t *row_kind, int *counter, cudaStream_t s)
{
  using namespace tc;
  TcParams p{};
  p.in = in; p.out = out; p.stride = stride; p.ostride = ostride; p.rows = rows; p.L = L; p.K = K;
  p.bmat = bmat; p.row_set = row_set; p.row_kind = row_kind; p.counter = counter;
  p.n_row_blocks = (rows + M - 1) / M;
  p.n_time_tiles = L / N;
  if (p.n_row_blocks == 0 || p.n_time_tiles == 0) return cudaSuccess;
  if (K % 32 || K / 32 > RING - 2) return cudaErrorInvalidValue;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  // a work item re-converts K/32 - 1 warm-up pairs; pick the split of the time axis that minimises the critical path
  // ceil(items / SMs) * (tiles + warm-up) of the dynamic schedule
  const uint32_t KS = K / 32;
  uint32_t best_ipb = 1;
  uint64_t best_cost = ~0ull;
  for (uint32_t ipb = 1; ipb <= p.n_time_tiles && ipb <= 64; ++ipb) {
    const uint32_t tpi = (p.n_time_tiles + ipb - 1) / ipb, real_ipb = (p.n_time_tiles + tpi - 1) / tpi;
    const uint64_t waves = ((uint64_t)p.n_row_blocks * real_ipb + (uint32_t)sms - 1) / (uint32_t)sms;
    const uint64_t cost = waves * (tpi + KS - 1 + 2); // + 2: pipeline fill/drain per item
    if (cost < best_cost) { best_cost = cost; best_ipb = ipb; }
  }
  p.tiles_per_item = (p.n_time_tiles + best_ipb - 1) / best_ipb;
  p.items_per_block = (p.n_time_tiles + p.tiles_per_item - 1) / p.tiles_per_item;
  const size_t smem = (size_t)4 * kPlaneBytes + (size_t)4 * N * K + (size_t)M * OW * 4 + sizeof(TcCtrl) + 1024;
  e = cudaFuncSetAttribute(fir_demod_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  uint32_t grid = (uint32_t)sms;
  const uint32_t n_items = p.n_row_blocks * p.items_per_block;
  if (grid > n_items) grid = n_items;
  fir_demod_tc_kernel<<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}

} // namespace msdr
