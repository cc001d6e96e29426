// msdr_chain_v4.cu — K1 (tensor-core form): the fused receive chain with the FIR pair on tcgen05.mma kind::i8.
//
//   int16 IF samples -> [fs/4 mix folded into the byte planes] -> FIR pair as four int8 Toeplitz GEMMs per branch
//   (exact mod 2^32, msdr_fir_tc.cu) -> >>15, SSAT16 -> SSB sum / AM envelope -> biquad cascade -> int16 audio
//
// Reference semantics: Minimal-SDR.ino:546-558 (mix), arm_fir_fast_q15.c:60-329 (FIR), Minimal-SDR.ino:589-628 (demod),
// filter_biquad.cpp:33-82 (biquad).
//
// Same decoupling as msdr_chain_v3.cu — FIR work is produced by any SM, biquad chains are pinned to an SM for the whole
// launch and only ever wait for their own input — but the FIR producers are a tensor-core pipeline, which takes the FIR off
// the integer/FP64 issue slots the serial biquad recurrence needs.  One persistent CTA per SM, 16 warps, roles by warp id
// (warp id % 4 = SM sub-partition):
//
//   warps 0-3   epilogue   TMEM lane quadrant = warp id % 4 (hardware rule): thread = channel row.  tcgen05.ld, recombine the
//                          four byte-plane products, >>15, SSAT16, demodulate, coalesced store of the 128 x 64 tile to `out`
//                          (used as the intermediate buffer), then count the rows into tile_cnt[group][span] (release).
//   warp 4, 5   chains     biquad object 1 / object 2 of channel group g = wave * grid + blockIdx (lane = channel, state in
//                          registers), a two-warp stage pipeline over 128-sample sub-tiles exactly as in v3.  Each chain warp
//                          shares its sub-partition only with one epilogue warp.
//   warp 6      MMA        one thread issues 2 x 4 x K/32 MMAs (M128 N64 K32) per tile; tcgen05.commit -> mbarriers.
//   warps 7,10,11,14,15    convert: raw int16 (global / carried history) -> sign-folded byte planes in a ring, each window
//                          word converted once per work item.
//
// A FIR work item = (row block of 128 channels sharing one tap table, span of 512 samples); items are claimed from a global
// counter in wave-major, then time-major order so that a chain's input is produced while it runs.  Rows are gathered through
// a row map (channels sorted by tap table inside each wave); a chain waits until all rows of its group have been counted.
#include "msdr_chain_common.cuh"
#include "msdr_tc_common.cuh"

namespace msdr {
namespace v4 {

using namespace tc;

constexpr int SPAN = 512;          // samples per FIR work item
constexpr int UNIT = 128;          // samples per readiness counter: a chain starts on a span after its first two tiles
// chain sub-tile ring: one slot being filtered by warp A, one by warp B, one draining to `out`, two landing.  With fewer
// slots the load of a sub-tile can only be issued when it is already needed and its latency (~2 k cycles) is paid per sub-tile.
constexpr int NSLOT = 5;
// sub-tile length p.sub = 128 samples (64 for windows K > 128, where the B operand needs the room); row pitch sub/2 + 4 words
// (4 mod 32: conflict-free row-wise LDS.128)
constexpr int kChainA = 4, kChainB = 5, kMmaWarp = 6, kLoadWarp = 8, kStoreWarp = 9;
constexpr int kThreads = 16 * 32;
constexpr int NCONV = 5;
constexpr int kLive = (4 + 1 + NCONV) * 32;
__host__ __device__ inline uint32_t slot_bytes(uint32_t sub) { return kGroup * (sub / 2 + 4) * 4u; }
constexpr uint32_t kCtrlBytes = 1024;

__device__ __forceinline__ bool is_conv_warp(int w) { return w == 7 || w == 10 || w == 11 || w == 14 || w == 15; }
__device__ __forceinline__ int conv_index(int w) { return w == 7 ? 0 : w == 10 ? 1 : w == 11 ? 2 : w == 14 ? 3 : 4; }

struct __align__(16) Ctrl {
  uint64_t a_full[RING_MAX];   // convert -> MMA   : pair converted (NCONV arrivals)
  uint64_t blk_free[RING_MAX]; // MMA -> convert   : pair no longer read (tcgen05.commit)
  uint64_t tmem_full;          // MMA -> epilogue  : accumulators complete (tcgen05.commit)
  uint64_t tmem_empty;         // epilogue -> MMA  : accumulators drained (4 arrivals)
  uint64_t ld_full[NSLOT];     // load -> chain A  : sub-tile in smem
  uint64_t ab_full[NSLOT];     // chain A -> B     : object 1 done
  uint64_t st_full[NSLOT];     // chain B -> store : object 2 done
  uint64_t slot_free[NSLOT];   // store -> load    : final audio written back, slot reusable
  uint32_t tmem_base;
  int item_rb[2], item_span[2]; // work item of the even / odd iteration (rb < 0: done)
  uint32_t rowmap[M];           // rows of the current item (written by the epilogue threads)
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "Ctrl must fit its smem slot");

uint32_t sub_for(uint32_t K) { return K > 128 ? 64 : 128; }
size_t smem_bytes(uint32_t K, uint32_t ring) { return (size_t)kCtrlBytes + 4u * a_plane_bytes(ring) + 4u * N * K + kStagingBytes + (size_t)NSLOT * slot_bytes(sub_for(K)) + 1024u; }

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
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <class BQ>
__device__ __forceinline__ void chain_span(BQ (&st)[1], uint4 *row, int nq)
{
  uint4 nxt = row[0];
#pragma unroll 1
  for (int q = 0; q < nq; ++q) {
    uint4 v = nxt;
    if (q + 1 < nq) nxt = row[q + 1]; // the load latency hides behind the recurrence
    v.x = bq_word<1>(st, v.x);
    v.y = bq_word<1>(st, v.y);
    v.z = bq_word<1>(st, v.z);
    v.w = bq_word<1>(st, v.w);
    row[q] = v;
  }
}

template <class BQ>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const ChainParams p)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the operand rings want 128-byte alignment; round the dynamic window up to 1 KB to be independent of the static layout
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~(uintptr_t)1023u);
  const uint32_t K = p.tc_K, KS = K / 32, ring = p.tc_ring;
  const uint32_t a_plane = a_plane_bytes(ring), b_plane = N * K;
  Ctrl *pc = reinterpret_cast<Ctrl *>(smem);
  uint8_t *sA = smem + kCtrlBytes;
  uint8_t *sB = sA + 4 * a_plane;
  uint32_t *sOut = reinterpret_cast<uint32_t *>(sB + 4 * b_plane);
  unsigned char *bq_base = reinterpret_cast<unsigned char *>(sOut) + kStagingBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < RING_MAX; ++i) { mbar_init(&pc->a_full[i], NCONV); mbar_init(&pc->blk_free[i], 1); }
    mbar_init(&pc->tmem_full, 1);
    mbar_init(&pc->tmem_empty, 4);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&pc->ld_full[s], 1); mbar_init(&pc->ab_full[s], 1); mbar_init(&pc->st_full[s], 1); mbar_init(&pc->slot_free[s], 1); }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_512(&pc->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = pc->tmem_base;
  const bool is_epi = warp < 4, is_mma = warp == kMmaWarp, is_conv = is_conv_warp(warp);

  if (is_epi || is_mma || is_conv) {
    // ================================================================== FIR + demod producers (tensor-core pipeline)
    Prof prof(p.prof, is_conv ? 0 : is_mma ? 1 : 2);
    IssueCtx ictx;
    issue_init(ictx, sA, a_plane, sB, b_plane);
    uint32_t q = 0;     // pairs converted so far by this CTA (ring position q % ring, phase q / ring)
    uint32_t ntile = 0; // tiles processed so far by this CTA (TMEM hand-off phases)
    int cur_set = -1;
    uint32_t cur_wave = 0, wave_item0 = 0; // claiming thread only
    const uint32_t n_tiles = p.L / N;
    for (uint32_t iter = 0;; ++iter) {
      if (is_mma && lane == 0) {
        const uint32_t it = (uint32_t)atomicAdd(&p.ctrl[0], 1);
        int rb = -1, span = 0;
        if (it < p.n_items) {
          for (;;) { // items are claimed in increasing order: walk the waves forward
            const uint32_t nrb = p.tc_wave_rb0[cur_wave + 1] - p.tc_wave_rb0[cur_wave];
            if (it - wave_item0 < nrb * p.NT) {
              const uint32_t rem = it - wave_item0;
              span = (int)(rem / nrb);
              rb = (int)(p.tc_wave_rb0[cur_wave] + (rem - (uint32_t)span * nrb));
              break;
            }
            wave_item0 += nrb * p.NT;
            ++cur_wave;
          }
        }
        pc->item_rb[iter & 1] = rb;
        pc->item_span[iter & 1] = span;
      }
      prof.start();
      named_bar_sync(1, kLive); // item visible; every producer role has finished the previous item
      prof.lap(3);
      const int rb = pc->item_rb[iter & 1];
      if (rb < 0) break;
      const uint32_t span = (uint32_t)pc->item_span[iter & 1];
      const uint32_t tb = span * (SPAN / N), te = min(tb + SPAN / N, n_tiles);
      const uint32_t qbase = q;
      const uint32_t npairs = (te - tb) + KS - 1;
      const uint4 rbi = __ldg(p.tc_rb + rb); // x: table id, y: first group entry, z: number of group entries
      const uint32_t *rmap = p.tc_rowmap + (size_t)rb * M;

      if (is_conv) {
        const int ctid = conv_index(warp) * 32 + lane;
        if ((int)rbi.x != cur_set) { // (re)load the Toeplitz operand of this table; the pipeline is drained at item boundaries
          const uint4 *src = reinterpret_cast<const uint4 *>(p.tc_bmat + (size_t)rbi.x * 4 * b_plane);
          uint4 *dst = reinterpret_cast<uint4 *>(sB);
          for (uint32_t i = ctid; i < 4 * b_plane / 16; i += NCONV * 32) dst[i] = __ldg(src + i);
          cur_set = (int)rbi.x;
        }
        // this thread's (at most two) conversion tasks: row r, K-block kb of every pair
        const uint32_t task0 = ctid, task1 = ctid + NCONV * 32;
        const uint32_t r0 = task0 % M, kb0 = task0 / M, r1 = task1 % M, kb1 = task1 / M;
        const uint32_t row_a = __ldg(rmap + r0), row_b = task1 < 2 * M ? __ldg(rmap + r1) : 0xFFFFFFFFu;
        const int Hs = (int)p.H;
        auto load16 = [&](uint32_t row, long long s0, uint4 (&v)[4]) { // 32 samples from sample index s0 (may be negative: history)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long s = s0 + 8 * j;
            if (row == 0xFFFFFFFFu || s < -(long long)Hs) v[j] = make_uint4(0, 0, 0, 0); // padding row / beyond the taps' reach
            else if (s >= 0) v[j] = __ldg(reinterpret_cast<const uint4 *>(p.in + (size_t)row * p.stride + s));
            else v[j] = __ldcg(reinterpret_cast<const uint4 *>(p.hist + ((size_t)p.ch0 + row) * p.H + (Hs + s)));
          }
        };
        for (uint32_t u = 0; u < npairs; ++u, ++q) {
          const uint32_t pos = q % ring;
          prof.start();
          mbar_wait(&pc->blk_free[pos], ((q / ring) & 1u) ^ 1u);
          prof.lap(0);
          const long long w0 = ((long long)tb - (long long)(KS - 1) + (long long)u) * P; // first window word of the pair
          if (!(p.ablate & 1u)) {
            uint4 v[4];
            load16(row_a, 2 * (w0 + 16 * (long long)kb0), v);
            convert_store(sA, a_plane, pos, kb0, r0, v);
            if (task1 < 2 * M) {
              load16(row_b, 2 * (w0 + 16 * (long long)kb1), v);
              convert_store(sA, a_plane, pos, kb1, r1, v);
            }
          }
          fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&pc->a_full[pos]);
          prof.lap(1);
        }
      } else if (is_mma) {
        q += npairs;
        if (lane == 0) {
          uint32_t nt = ntile;
          for (uint32_t t = tb; t < te; ++t, ++nt) {
            const uint32_t qt = qbase + (t - tb) + KS - 1; // newest pair of this tile's window
            prof.start();
            mbar_wait(&pc->a_full[qt % ring], (qt / ring) & 1u);
            prof.lap(0);
            mbar_wait(&pc->tmem_empty, (nt & 1u) ^ 1u);
            prof.lap(1);
            tc_fence_after();
            if (!(p.ablate & 1u)) issue_tile(ictx, tmem, qt, KS, ring);
            umma_commit(&pc->tmem_full);
            umma_commit(&pc->blk_free[(qt - (KS - 1)) % ring]); // the oldest pair of the window is not read again
            prof.lap(2);
          }
          for (uint32_t s = 1; s < KS; ++s) umma_commit(&pc->blk_free[(qbase + npairs - KS + s) % ring]);
        }
        ntile += te - tb;
        __syncwarp();
      } else { // epilogue: thread = channel row (TMEM lane)
        q += npairs;
        const uint32_t row = __ldg(rmap + tid);
        pc->rowmap[tid] = row;
        const int kind = row != 0xFFFFFFFFu ? demod_kind_of((int)p.mode[p.ch0 + row], p.am_q31) : 0;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t *orow = sOut + (uint32_t)tid * OW;
        for (uint32_t t = tb; t < te; ++t, ++ntile) {
          prof.start();
          mbar_wait(&pc->tmem_full, ntile & 1u);
          prof.lap(0);
          tc_fence_after();
          named_bar_sync(2, 128); // the previous tile's staging rows have been copied out (and rowmap is complete)
          if (!(p.ablate & 1u)) drain_tile(lane_addr, orow);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pc->tmem_empty); // the next tile's MMAs overlap the demodulation
          prof.lap(1);
          if (!(p.ablate & 1u)) demod_row(orow, kind);
          named_bar_sync(2, 128);
          for (uint32_t i = tid; i < M * (N / 8) && !(p.ablate & 1u); i += 128) { // coalesced write-back: 128 rows x 128 bytes
            const uint32_t r = i / (N / 8), c = i % (N / 8);
            const uint32_t orow_g = pc->rowmap[r];
            if (orow_g != 0xFFFFFFFFu)
              *reinterpret_cast<uint4 *>(p.out + (size_t)orow_g * p.stride + (size_t)t * N + c * 8) = *reinterpret_cast<const uint4 *>(sOut + r * OW + c * 4);
          }
          if ((((t + 1) * N) % UNIT) == 0) { // publish: every row of this block has the unit in `out`
            __threadfence();
            named_bar_sync(2, 128);
            if ((uint32_t)tid < rbi.z) {
              const uint32_t e = __ldg(p.tc_grp + rbi.y + tid);
              atomicAdd(p.tile_cnt + (size_t)(e & 0xFFFFFFu) * p.NU + (t * N) / UNIT, (int)(e >> 24));
            }
          }
          prof.lap(2);
        }
      }
    }
    if (warp == 0 || warp == kMmaWarp || warp == 7) prof.flush();
    if (warp == 0) {
      tc_fence_before();
      tmem_dealloc_512(tmem);
    }
  } else if (warp == kChainA || warp == kChainB) {
    // ================================================================== biquad chain: warp A = object 1, warp B = object 2
    const bool isA = (warp == kChainA);
    const int obj = isA ? 0 : 1;
    const int SUB = (int)p.tc_sub, BW = SUB / 2 + 4;
    const uint32_t kSlotBytes = slot_bytes(p.tc_sub);
    const int nsub = (int)((p.L + SUB - 1) / SUB);
    Prof prof(p.prof, isA ? 3 : 4);
    uint32_t pos = 0; // sub-tiles handled so far by this CTA's chain (ring position; identical in all four chain-side warps)
    for (int g = (int)blockIdx.x; g < (int)p.NG; g += (int)p.W) {
      const uint32_t row = (uint32_t)(g * kGroup + lane), ch = p.ch0 + row;
      const bool active = row < p.C;

      // cascade structure of this lane's object: stages run while bit31 of word 7 says another follows (filter_biquad.cpp:75,79)
      int nst = 1;
      BQ st[1];
      uint32_t fl = 0u;
      if (active) {
        for (int k = 0; k < 3 && (nst == k + 1); ++k)
          if ((uint32_t)__ldcg(p.bq + (size_t)((obj * 4 + k) * 8 + 7) * p.Cpad + ch) & 0x80000000u) nst = k + 2;
      }
      const bool fast = __all_sync(0xffffffffu, nst == 1);
      if (fast && active) bq_load_stage(st[0], fl, p.bq, p.Cpad, obj, 0, ch);

      for (int k = 0; k < nsub; ++k, ++pos) {
        const int slot = (int)(pos % NSLOT);
        const uint32_t phs = (pos / NSLOT) & 1u;
        prof.start();
        mbar_wait(isA ? &pc->ld_full[slot] : &pc->ab_full[slot], phs);
        prof.lap(0);
        const int nq = (int)(min((uint32_t)SUB, p.L - (uint32_t)k * SUB) >> 3);
        uint4 *r0 = reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(bq_base + (uint32_t)slot * kSlotBytes) + (uint32_t)lane * BW);
        if (!(p.ablate & 2u)) {
          if (fast) {
            if (active) chain_span(st, r0, nq);
          } else { // generic cascade: stage-major over the sub-tile like the reference (filter_biquad.cpp:44-79); state in global
            const int nmax = __reduce_max_sync(0xffffffffu, active ? nst : 0);
            for (int j = 0; j < nmax; ++j) {
              if (active && j < nst) {
                BQ gs[1];
                uint32_t gf;
                bq_load_stage(gs[0], gf, p.bq, p.Cpad, obj, j, ch);
                chain_span(gs, r0, nq);
                bq_store_stage(gs[0], gf, p.bq, p.Cpad, obj, j, ch);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(isA ? &pc->ab_full[slot] : &pc->st_full[slot]);
        prof.lap(1);
      }
      if (fast && active) bq_store_stage(st[0], fl, p.bq, p.Cpad, obj, 0, ch);
    }
    prof.flush();
  } else if (warp == kLoadWarp || warp == kStoreWarp) {
    // ================================================================== chain I/O: sub-tiles `out` -> smem ring -> `out`
    // Plain coalesced 16-byte accesses by two dedicated warps (32 per-row bulk copies per direction kept the SM's copy engine
    // busier than the biquad itself).  A warp instruction covers 32 >> lg rows of 1 << lg chunks; all addresses advance by
    // constants, so a sub-tile costs 32 << lg >> 5 loads and stores per lane and little else.
    const bool isLoad = (warp == kLoadWarp);
    const int SUB = (int)p.tc_sub, BW = SUB / 2 + 4;
    const uint32_t kSlotBytes = slot_bytes(p.tc_sub);
    const int nsub = (int)((p.L + SUB - 1) / SUB); // L is a multiple of 128 and SUB is 64 or 128: sub-tiles are always full
    Prof prof(p.prof, isLoad ? 5 : 6);
    const int lg = SUB == 128 ? 4 : 3;           // log2 of the 16-byte chunks per row
    const int rpi = 32 >> lg, nins = kGroup / rpi; // rows per warp instruction, instructions per sub-tile
    const int r0 = lane >> lg, c = lane & ((1 << lg) - 1);
    const size_t gstep = (size_t)rpi * p.stride * 2u;  // bytes between the rows of consecutive instructions
    const uint32_t sstep = (uint32_t)rpi * BW * 4u;
    uint32_t pos = 0;
    for (int g = (int)blockIdx.x; g < (int)p.NG; g += (int)p.W) {
      const int nrows = min(kGroup, (int)p.C - g * kGroup);
      unsigned char *gp0 = reinterpret_cast<unsigned char *>(p.out + ((size_t)g * kGroup + r0) * p.stride) + c * 16;
      int ready = 0; // leading units known to be in `out` for every row of the group
      for (int k = 0; k < nsub; ++k, ++pos) {
        const int slot = (int)(pos % NSLOT);
        const uint32_t phs = (pos / NSLOT) & 1u;
        unsigned char *sp0 = bq_base + (uint32_t)slot * kSlotBytes + (uint32_t)r0 * BW * 4u + c * 16;
        unsigned char *gp = gp0 + (size_t)k * SUB * 2u;
        if (isLoad) {
          prof.start();
          const int need = (k * SUB) / UNIT;
          while (ready <= need) { // all rows of the group have this unit in `out`
            if (lane == 0) {
              const long long t0 = clock64();
              while (ld_relaxed_gpu(p.tile_cnt + (size_t)g * p.NU + ready) < nrows) {
                __nanosleep(64);
                if (clock64() - t0 > kWatchdogCycles) __trap();
              }
              fence_acquire_gpu();
            }
            __syncwarp();
            ++ready;
          }
          prof.lap(0);
          mbar_wait(&pc->slot_free[slot], phs ^ 1u);
          prof.lap(1);
          // weak loads: the acquire fence above ordered them after the producers' stores (and dropped this SM's L1 lines)
          for (int i0 = 0; i0 < nins; i0 += 8) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (r0 + (i0 + j) * rpi < nrows) v[j] = *reinterpret_cast<const uint4 *>(gp + (size_t)(i0 + j) * gstep);
            prof.lap(2);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (r0 + (i0 + j) * rpi < nrows) *reinterpret_cast<uint4 *>(sp0 + (uint32_t)(i0 + j) * sstep) = v[j];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&pc->ld_full[slot]);
          prof.lap(3);
        } else {
          prof.start();
          mbar_wait(&pc->st_full[slot], phs);
          prof.lap(0);
          for (int i0 = 0; i0 < nins; i0 += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (r0 + (i0 + j) * rpi < nrows)
                *reinterpret_cast<uint4 *>(gp + (size_t)(i0 + j) * gstep) = *reinterpret_cast<const uint4 *>(sp0 + (uint32_t)(i0 + j) * sstep);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&pc->slot_free[slot]);
          prof.lap(1);
        }
      }
      if (isLoad) {
        // carry the last H raw samples: hist <- tail of (hist || in[0..L)).  Every span of this group has been counted and
        // only a span's first window reaches back into the history, so nobody reads the old history any more.
        const uint32_t hq = p.H >> 3; // uint4 per history row (<= 33)
        for (int r = 0; r < nrows; ++r) {
          const size_t c = (size_t)g * kGroup + (size_t)r;
          uint4 *hrow = reinterpret_cast<uint4 *>(p.hist + ((size_t)p.ch0 + c) * p.H);
          const uint4 *irow = reinterpret_cast<const uint4 *>(p.in + c * p.stride);
          uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
          const uint32_t i0 = (uint32_t)lane, i1 = (uint32_t)lane + 32u;
          if (p.L >= p.H) {
            const uint4 *src = irow + ((p.L - p.H) >> 3);
            if (i0 < hq) v0 = src[i0];
            if (i1 < hq) v1 = src[i1];
          } else {
            const uint32_t lq = p.L >> 3, keep = hq - lq; // keep = old entries that survive
            if (i0 < hq) v0 = (i0 < keep) ? __ldcg(hrow + i0 + lq) : irow[i0 - keep];
            if (i1 < hq) v1 = (i1 < keep) ? __ldcg(hrow + i1 + lq) : irow[i1 - keep];
          }
          __syncwarp();
          if (i0 < hq) hrow[i0] = v0;
          if (i1 < hq) hrow[i1] = v1;
        }
      }
    }
    prof.flush();
  }
}

} // namespace v4

uint32_t chain_v4_span_samples() { return v4::SPAN; }
uint32_t chain_v4_unit_samples() { return v4::UNIT; }

// largest window the fused kernel can hold next to its other buffers; 0 = the tensor-core form does not apply
bool chain_v4_config(uint32_t K, int smem_max, uint32_t *ring_out)
{
  if (K % 32u || K / 32u < 2u) return false;
  for (uint32_t ring = tc::RING_MAX; ring > K / 32u; --ring) {
    if (v4::smem_bytes(K, ring) <= (size_t)smem_max) { *ring_out = ring; return true; }
  }
  return false;
}

cudaError_t launch_chain_v4(const ChainParams &p_in, cudaStream_t stream, int variant, ChainLaunchInfo *info)
{
  using namespace v4;
  ChainParams p = p_in;
  p.ablate = ((uint32_t)variant >> 4) & 3u;
  p.tc_sub = sub_for(p.tc_K);
  const size_t smem = smem_bytes(p.tc_K, p.tc_ring);
  auto kern = chain_kernel<BqStage>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (info) { info->grid = (int)p.W; info->block = kThreads; info->smem = smem; info->tile = SPAN; }
  kern<<<p.W, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

} // namespace msdr
