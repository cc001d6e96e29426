// msdr_chain_v3.cu — K1 (current): the fused receive chain with FIR producers decoupled from pinned biquad chains.
//
//   int16 IF samples -> [fs/4 mix folded into tap selection] -> FIR pair (Q15, 32-bit wrapping accumulate)
//   -> SSB sum / AM envelope -> biquad cascade (Q2.30 x int16, 14-bit error feedback) -> int16 audio
//
// Reference semantics: Minimal-SDR.ino:546-558 (mix), arm_fir_fast_q15.c:60-329 (FIR), Minimal-SDR.ino:589-628 (demod),
// filter_biquad.cpp:33-82 (biquad).
//
// Why this shape.  The FIR is parallel in time (given a (T-1)-sample halo), the biquad is an exact-arithmetic serial
// recurrence per channel (truncation + saturation + residual feedback: no associative scan).  Measured on B200 one warp
// needs ~75-100 cycles per biquad sample-step (tools/microbench/bqstep.cu), so a channel's chain must never wait for
// anything but its own input, and everything else must be spread over the whole chip.  One persistent CTA per SM:
//
//   warp 0          producer   claims FIR tiles (channel group g of 32, time tile i of 512 samples) from a global counter,
//                              time-major, and streams halo + tile HBM -> smem with cp.async.bulk (TMA engine, UBLKCP).
//   8 FIR warps     FIR        fold the fs/4 oscillator sign in place, then one channel row per warp at a time: lane l
//                              computes 16 consecutive outputs of all four polyphase sub-filters from a rotating
//                              register window (msdr_chain_common.cuh), >>15, SSAT, demod -> smem.
//   warp 11         store      bulk-copies the demodulated tile smem -> `out` (used as the intermediate buffer: it is
//                              rewritten in place by the biquad), then publishes flag[g][i] = epoch (release).
//   chain A, B      chains     the CTA owns channel group g = wave * grid + blockIdx for the WHOLE launch (lane = channel, state
//                              in registers).  A two-warp stage pipeline over 128-sample sub-tiles: warp A acquires the tile
//                              flag, bulk-loads the 32 rows from `out` (L2-hot), runs biquad object 1 in smem and hands the
//                              buffer to warp B, which runs object 2 and bulk-stores the final audio in place.
//
// Warp -> scheduler placement is deliberate (warp id % 4 selects the SM sub-partition, tools/microbench/placement.cu):
// the latency-critical chain warps share sub-partition 0 with the two sleeping I/O warps only (an IMAD-saturating
// neighbour on the same sub-partition costs a chain +30..150 %, on another one nothing); the eight FIR warps own
// sub-partitions 1-3 (with the sleeping store warp) and drive BOTH their IMAD and DFMA pipes.  12 warps keep 170 registers
// per thread available to the FIR inner loop.
//
// FIR tiles of any group are produced by any SM, so FIR work balances for any channel count; chains are pinned, so no
// state ever migrates and FIR warps never wait for a biquad.  Groups are taken in waves of `grid` chains so that a
// chain's inputs are produced while it runs.
#include "msdr_chain_common.cuh"

namespace msdr {
namespace v3 {

constexpr int TT = 512;            // FIR tile, samples
constexpr int NF = 8;              // FIR warps (warp ids 1,2,3,5,6,7,9,10: sub-partitions 1-3)
constexpr int NSLOT = 3;           // chain sub-tile ring
constexpr int R = TT / 64;         // output pairs per FIR lane
constexpr int DW = TT / 2 + 4;     // demod row pitch, words (4 mod 32: conflict-free row-wise LDS.128)
constexpr int kProducerWarp = 0, kChainA = 4, kChainB = 8, kStoreWarp = 11;
constexpr int kThreads = 12 * 32;

struct Tile {
  int grp, tile, t0, len;
  uint32_t flags;
  int pad[3];
};
constexpr uint32_t TF_END = 0x80000000u;

struct __align__(16) Ctrl {
  uint64_t full[2];   // producer -> FIR   : raw tile landed (tx bytes)
  uint64_t empty[2];  // FIR -> producer   : raw tile consumed
  uint64_t dfull[2];  // FIR -> store warp : demodulated tile complete in smem
  uint64_t dfree[2];  // store warp -> FIR : demodulated tile drained
  uint64_t ld_full[NSLOT];   // TMA -> chain A : sub-tile landed
  uint64_t ab_full[NSLOT];   // chain A -> B   : object 1 done
  uint64_t slot_free[NSLOT]; // chain B -> A   : final audio drained, slot reusable
  Tile desc[2];
  Tile ddesc[2];
  int job_ctr[2];
  int pad[2];
  uint32_t rowinfo[2][kGroup]; // setid | mode << 8 | kp4 << 16
};
constexpr uint32_t kCtrlBytes = 640;
static_assert(sizeof(Ctrl) <= kCtrlBytes, "Ctrl must fit its smem slot");

__host__ __device__ inline uint32_t raw_stage_bytes(uint32_t H) { return kGroup * (H + TT) * 2u; }
constexpr uint32_t kDStage = kGroup * DW * 4u;
// chain sub-tile ring: KCH * 32 rows of SUB samples.  KCH = channels per chain lane (a chain owns KCH * 32 channels):
//   KCH = 1, SUB = 256 when channel groups are scarce (one chain per SM already covers them: BASELINE config 3),
//   KCH = 2, SUB = 128 when there are at least two groups per SM: the second channel per lane rides in the issue slots the
//   latency-bound recurrence leaves idle (+56 % chain throughput per SM, measured).
// Row pitch SUB/2 + 4 words = 4 * odd: conflict-free row-wise LDS.128.  Both shapes use 52 KB.
constexpr uint32_t kBqBuf = kGroup * (256 / 2 + 4) * 4u; // == 2 * kGroup * (128 / 2 + 4) * 4 up to 1 KB

size_t smem_bytes(uint32_t H, uint32_t n_sets, uint32_t set_stride_words)
{
  return (size_t)kCtrlBytes + align_up(n_sets * set_stride_words * 4u, 128u) + 2u * raw_stage_bytes(H) + 2u * kDStage + (size_t)NSLOT * (kBqBuf + 1024u);
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// item -> (group, tile): waves of KCH * W groups (= W chains), time-major inside a wave
__device__ __forceinline__ void decode_item(const ChainParams &p, const int KCH, int item, int &grp, int &tile)
{
  const int gw = KCH * (int)p.W; // groups per wave
  const int per_wave = gw * (int)p.NT;
  const int wave = item / per_wave;
  const int rem = item - wave * per_wave;
  const int g0 = wave * gw;
  const int wcur = min(gw, (int)p.NG - g0);
  tile = rem / wcur;
  grp = g0 + (rem - tile * wcur);
}

// one 128-sample sub-tile of one chain: NS stages fused, in place in smem (lane = channel row)
template <int NS, class BQ>
__device__ __forceinline__ void chain_span(BQ (&st)[NS], uint4 *row, int q0, int q1)
{
#pragma unroll 1
  for (int q = q0; q < q1; ++q) {
    uint4 v = row[q];
    v.x = bq_word<NS>(st, v.x);
    v.y = bq_word<NS>(st, v.y);
    v.z = bq_word<NS>(st, v.z);
    v.w = bq_word<NS>(st, v.w);
    row[q] = v;
  }
}

template <int KCH, class BQ>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const ChainParams p)
{
  constexpr int SUB = KCH == 1 ? 256 : 128; // chain sub-tile, samples
  constexpr int BW = SUB / 2 + 4;           // chain buffer row pitch, words
  constexpr uint32_t kSlotBytes = kBqBuf + 1024u;
  extern __shared__ __align__(128) unsigned char smem[];
  Ctrl *pc = reinterpret_cast<Ctrl *>(smem);
  int32_t *s_sets = reinterpret_cast<int32_t *>(smem + kCtrlBytes);
  const uint32_t sets_words = p.sets_in_smem ? p.n_sets * p.set_stride_words : 0u;
  const uint32_t RS = p.H + TT; // raw row pitch, samples
  unsigned char *raw_base = smem + kCtrlBytes + align_up(sets_words * 4u, 128u);
  const uint32_t raw_stage = raw_stage_bytes(p.H);
  unsigned char *d_base = raw_base + 2u * raw_stage;
  unsigned char *bq_base = d_base + 2u * kDStage;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pc->full[s], 1);
      mbar_init(&pc->empty[s], NF);
      mbar_init(&pc->dfull[s], NF);
      mbar_init(&pc->dfree[s], 1);
    }
    for (int q = 0; q < NSLOT; ++q) {
      mbar_init(&pc->ld_full[q], 1);
      mbar_init(&pc->ab_full[q], 1);
      mbar_init(&pc->slot_free[q], 1);
    }
    mbar_fence_init();
  }
  for (uint32_t i = threadIdx.x; i < sets_words; i += blockDim.x) s_sets[i] = p.sets[i];
  __syncthreads();

  const int Hw = (int)(p.H >> 1);

  const bool is_fir = (warp & 3) != 0 && warp != kStoreWarp;
  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ producer
    uint32_t it = 0;
    for (;; ++it) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&p.ctrl[0], 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= (int)p.n_items) break;
      int grp, tile;
      decode_item(p, KCH, item, grp, tile);
      const int nrows = min(kGroup, (int)p.C - grp * kGroup);
      const uint32_t row = (uint32_t)(grp * kGroup + lane); // row of in/out
      const uint32_t ch = p.ch0 + row;                      // channel of the chain object
      uint32_t rinfo = 0;
      if (lane < nrows) {
        const uint32_t set = p.setid[ch];
        rinfo = set | ((uint32_t)p.mode[ch] << 8) | (p.set_kp4[set] << 16);
      }
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->empty[s], ph ^ 1u);
      const int t0 = tile * TT;
      const int len = min(TT, (int)p.L - t0);
      pc->rowinfo[s][lane] = rinfo;
      const uint32_t row_bytes = (p.H + (uint32_t)len) * 2u;
      if (lane == 0) {
        Tile td;
        td.grp = grp; td.tile = tile; td.t0 = t0; td.len = len; td.flags = 0u;
        td.pad[0] = td.pad[1] = td.pad[2] = 0;
        pc->desc[s] = td;
        pc->job_ctr[s] = 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&pc->full[s], (uint32_t)nrows * row_bytes);
      __syncwarp();
      if (lane < nrows) {
        int16_t *dst = reinterpret_cast<int16_t *>(raw_base + (uint32_t)s * raw_stage) + (uint32_t)lane * RS;
        if ((uint32_t)t0 >= p.H) {
          bulk_g2s(dst, p.in + (size_t)row * p.stride + (size_t)(t0 - (int)p.H), row_bytes, &pc->full[s]);
        } else { // halo (partly) from the carried history
          const uint32_t nh = p.H - (uint32_t)t0;
          bulk_g2s(dst, p.hist + (size_t)ch * p.H + (size_t)t0, nh * 2u, &pc->full[s]);
          bulk_g2s(dst + nh, p.in + (size_t)row * p.stride, (uint32_t)(t0 + len) * 2u, &pc->full[s]);
        }
      }
    }
    { // end marker travels through the same ring
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->empty[s], ph ^ 1u);
      if (lane == 0) {
        pc->desc[s].flags = TF_END;
        mbar_arrive(&pc->full[s]);
      }
    }
  } else if (is_fir) {
    // ------------------------------------------------------------------ FIR + demod warps
    const int fidx = warp - 1 - (warp >> 2); // 0..NF-1
    const int ftid = fidx * 32 + lane;
    for (uint32_t it = 0;; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->full[s], ph);
      const Tile td = pc->desc[s];
      if (td.flags & TF_END) {
        mbar_wait(&pc->dfree[s], ph ^ 1u);
        if (ftid == 0) pc->ddesc[s] = td;
        __syncwarp();
        if (lane == 0) mbar_arrive(&pc->dfull[s]);
        break;
      }
      const int nrows = min(kGroup, (int)p.C - td.grp * kGroup);
      uint32_t *rawW = reinterpret_cast<uint32_t *>(raw_base + (uint32_t)s * raw_stage);
      mbar_wait(&pc->dfree[s], ph ^ 1u);
      if (ftid == 0) pc->ddesc[s] = td;
      uint32_t *dW = reinterpret_cast<uint32_t *>(d_base + (uint32_t)s * kDStage);
      for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(&pc->job_ctr[s], 1);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= nrows) break;
        const uint32_t ri = pc->rowinfo[s][job];
        const uint32_t set = ri & 0xFFu;
        const int kind = demod_kind_of((int)((ri >> 8) & 0xFFu), p.am_q31);
        { // fold the fs/4 oscillator sign into this row: samples with n % 4 in {2,3} are negated (Minimal-SDR.ino:550,555),
          // i.e. every odd word.  Only this warp reads the row, so a warp-level sync is enough.
          uint4 *pw = reinterpret_cast<uint4 *>(rawW + (uint32_t)job * (RS >> 1));
          const int q4 = (int)((p.H + (uint32_t)td.len) >> 3); // uint4 per row
          for (int c4 = lane; c4 < q4; c4 += 32) {
            uint4 v = pw[c4];
            v.y = neg16x2(v.y);
            v.w = neg16x2(v.w);
            pw[c4] = v;
          }
          __syncwarp();
        }
        // taps normally sit in shared memory; many long tables (e.g. ten 256-tap sets) stay in global memory and stream
        // through L1.  One generic pointer keeps a single copy of the FIR code in the instruction cache.
        const int4 *cf = p.sets_in_smem ? reinterpret_cast<const int4 *>(s_sets + set * p.set_stride_words)
                                        : reinterpret_cast<const int4 *>(p.sets + (size_t)set * p.set_stride_words);
        if (!(p.ablate & 1u))
          fir_demod_row<R>(rawW + (uint32_t)job * (RS >> 1), cf, (int)(ri >> 16), Hw, lane, td.len, kind, dW + (uint32_t)job * DW);
      }
      fence_proxy_async_smem(); // raw[s] and d[s] were written through the generic proxy; the TMA engine touches both next
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&pc->dfull[s]);
        mbar_arrive(&pc->empty[s]);
      }
    }
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------------ store warp: demodulated tile -> `out`, publish
    for (uint32_t it = 0;; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->dfull[s], ph);
      const Tile td = pc->ddesc[s];
      if (td.flags & TF_END) break;
      const uint32_t row = (uint32_t)(td.grp * kGroup + lane);
      if (row < p.C) {
        const uint32_t *drow = reinterpret_cast<const uint32_t *>(d_base + (uint32_t)s * kDStage) + (uint32_t)lane * DW;
        bulk_s2g(p.out + (size_t)row * p.stride + (size_t)td.t0, drow, (uint32_t)td.len * 2u);
      }
      bulk_commit();
      bulk_wait<0>();          // the writes have been performed, not only the smem reads
      fence_proxy_async_all(); // async-proxy writes -> ordered before the generic-proxy release below
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        st_release_gpu(p.tile_flags + (size_t)td.grp * p.NT + td.tile, (int)p.epoch);
        mbar_arrive(&pc->dfree[s]);
      }
    }
  } else if (warp == kChainA || warp == kChainB) {
    // ------------------------------------------------------------------ biquad chain: warp A = object 1, warp B = object 2
    // KCH = 2 channels per lane: the recurrence is latency-bound (~45 cycles per step for ~16 instructions), so a second,
    // independent channel per lane rides in the idle issue slots.  A chain therefore owns a SUPER-group of 2 x 32 channels.
    const bool isA = (warp == kChainA);
    const int obj = isA ? 0 : 1;
    const int nsub = (int)((p.L + SUB - 1) / SUB); // the last sub-tile may be short
    const int NSG = ((int)p.NG + KCH - 1) / KCH;
    uint32_t pos = 0; // sub-tiles handled so far by this CTA's chain (ring position; identical in A and B)
    for (int sg = (int)blockIdx.x; sg < NSG; sg += (int)p.W) {
      uint32_t row[KCH], ch[KCH];
      bool active[KCH];
      int16_t *orow[KCH];
      int nrows_total = 0, ngrp = 0;
#pragma unroll
      for (int h = 0; h < KCH; ++h) {
        const int g = sg * KCH + h;
        row[h] = (uint32_t)(g * kGroup + lane);
        ch[h] = p.ch0 + row[h];
        active[h] = g < (int)p.NG && row[h] < p.C;
        orow[h] = p.out + (size_t)row[h] * p.stride;
        if (g < (int)p.NG) { nrows_total += min(kGroup, (int)p.C - g * kGroup); ++ngrp; }
      }

      // cascade structure of this lane's object: stages run while bit31 of word 7 says another follows (filter_biquad.cpp:75,79)
      int nst[KCH];
      BQ st[KCH][1];
      uint32_t fl[KCH];
      bool lane_fast = true;
#pragma unroll
      for (int h = 0; h < KCH; ++h) {
        nst[h] = 1;
        fl[h] = 0u;
        if (active[h]) {
          for (int k = 0; k < 3 && (nst[h] == k + 1); ++k)
            if ((uint32_t)__ldcg(p.bq + (size_t)((obj * 4 + k) * 8 + 7) * p.Cpad + ch[h]) & 0x80000000u) nst[h] = k + 2;
        }
        lane_fast = lane_fast && nst[h] == 1;
      }
      const bool fast = __all_sync(0xffffffffu, lane_fast);
#pragma unroll
      for (int h = 0; h < KCH; ++h)
        if (fast && active[h]) bq_load_stage(st[h][0], fl[h], p.bq, p.Cpad, obj, 0, ch[h]);

      // buffer layout of a ring slot: [KCH][32 rows][BW words]
      auto run_object = [&](uint32_t *slot_base, int nq) {
        if (p.ablate & 2u) return;
        uint4 *r0 = reinterpret_cast<uint4 *>(slot_base + (uint32_t)lane * BW);
        uint4 *r1 = reinterpret_cast<uint4 *>(slot_base + (uint32_t)((KCH - 1) * kGroup + lane) * BW);
        if (KCH == 2 && fast && active[0] && active[KCH - 1]) {
          // two independent recurrences interleaved word by word
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {
            uint4 v0 = r0[q], v1 = r1[q];
            v0.x = bq_word<1>(st[0], v0.x); v1.x = bq_word<1>(st[KCH - 1], v1.x);
            v0.y = bq_word<1>(st[0], v0.y); v1.y = bq_word<1>(st[KCH - 1], v1.y);
            v0.z = bq_word<1>(st[0], v0.z); v1.z = bq_word<1>(st[KCH - 1], v1.z);
            v0.w = bq_word<1>(st[0], v0.w); v1.w = bq_word<1>(st[KCH - 1], v1.w);
            r0[q] = v0; r1[q] = v1;
          }
        } else if (fast) {
          if (active[0]) chain_span<1>(st[0], r0, 0, nq);
          if (KCH == 2 && active[KCH - 1]) chain_span<1>(st[KCH - 1], r1, 0, nq);
        } else { // generic cascade: stage-major over the sub-tile like the reference (filter_biquad.cpp:44-79); state in global
#pragma unroll
          for (int h = 0; h < KCH; ++h) {
            const int nmax = __reduce_max_sync(0xffffffffu, active[h] ? nst[h] : 0);
            for (int j = 0; j < nmax; ++j) {
              if (active[h] && j < nst[h]) {
                BQ gs[1];
                uint32_t gf;
                bq_load_stage(gs[0], gf, p.bq, p.Cpad, obj, j, ch[h]);
                chain_span<1>(gs, h ? r1 : r0, 0, nq);
                bq_store_stage(gs[0], gf, p.bq, p.Cpad, obj, j, ch[h]);
              }
            }
          }
        }
      };

      if (isA) {
        int ready = 0; // leading FIR tiles known to be in `out` for every group of the super-group
        int ji = 0;    // next sub-tile to load
        // wait (or poll once) until FIR tile t of all member groups is published; on success order the async-proxy reads
        // after the acquire
        auto tile_ready = [&](int t, bool blocking) -> bool {
          bool advanced = false;
          while (ready <= t) {
            int ok = 0;
            if (lane == 0) {
              const long long t0 = clock64();
              for (;;) {
                ok = 1;
                for (int h = 0; h < ngrp; ++h)
                  if (ld_acquire_gpu(p.tile_flags + (size_t)(sg * KCH + h) * p.NT + ready) != (int)p.epoch) ok = 0;
                if (ok || !blocking) break;
                __nanosleep(100);
                if (clock64() - t0 > kWatchdogCycles) __trap();
              }
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) return false;
            ++ready;
            advanced = true;
          }
          if (advanced) fence_proxy_async_all();
          return true;
        };
        for (int k = 0; k < nsub; ++k) {
          // keep NSLOT - 2 sub-tiles in flight beyond the one being processed (one slot is with warp B)
          while (ji < nsub && ji <= k + NSLOT - 2) {
            const bool must = (ji == k);
            const uint32_t pj = pos + (uint32_t)(ji - k);
            const int sj = (int)(pj % NSLOT);
            const uint32_t phj = (pj / NSLOT) & 1u;
            if (!tile_ready((ji * SUB) / TT, must)) break;
            if (must) mbar_wait(&pc->slot_free[sj], phj ^ 1u);
            else if (!mbar_test_wait(&pc->slot_free[sj], phj ^ 1u)) break;
            const uint32_t lenj = min((uint32_t)SUB, p.L - (uint32_t)ji * SUB);
            if (lane == 0) mbar_arrive_expect_tx(&pc->ld_full[sj], (uint32_t)nrows_total * lenj * 2u);
            __syncwarp();
            uint32_t *sb = reinterpret_cast<uint32_t *>(bq_base + (uint32_t)sj * kSlotBytes);
#pragma unroll
            for (int h = 0; h < KCH; ++h)
              if (active[h]) bulk_g2s(sb + (uint32_t)(h * kGroup + lane) * BW, orow[h] + (size_t)ji * SUB, lenj * 2u, &pc->ld_full[sj]);
            ++ji;
          }
          const int slot = (int)(pos % NSLOT);
          const uint32_t phs = (pos / NSLOT) & 1u;
          mbar_wait(&pc->ld_full[slot], phs);
          run_object(reinterpret_cast<uint32_t *>(bq_base + (uint32_t)slot * kSlotBytes), (int)(min((uint32_t)SUB, p.L - (uint32_t)k * SUB) >> 3));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pc->ab_full[slot]);
          ++pos;
        }
      } else {
        for (int k = 0; k < nsub; ++k) {
          const int slot = (int)(pos % NSLOT);
          const uint32_t phs = (pos / NSLOT) & 1u;
          mbar_wait(&pc->ab_full[slot], phs);
          uint32_t *sb = reinterpret_cast<uint32_t *>(bq_base + (uint32_t)slot * kSlotBytes);
          const uint32_t lenk = min((uint32_t)SUB, p.L - (uint32_t)k * SUB);
          run_object(sb, (int)(lenk >> 3));
          fence_proxy_async_smem();
#pragma unroll
          for (int h = 0; h < KCH; ++h)
            if (active[h]) bulk_s2g(orow[h] + (size_t)k * SUB, sb + (uint32_t)(h * kGroup + lane) * BW, lenk * 2u);
          bulk_commit();
          if (k > 0) { // the previous sub-tile's copy has finished reading its slot by now
            bulk_wait_read<1>();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pc->slot_free[(pos - 1u) % NSLOT]);
          }
          ++pos;
        }
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pc->slot_free[(pos - 1u) % NSLOT]);
      }

#pragma unroll
      for (int h = 0; h < KCH; ++h)
        if (fast && active[h]) bq_store_stage(st[h][0], fl[h], p.bq, p.Cpad, obj, 0, ch[h]);
      if (!isA) {
        // carry the last H raw samples: hist <- tail of (hist || in[0..L)).  Every FIR tile of these groups has been
        // published (warp B consumed them all), so nobody reads the old history any more.
        const uint32_t hq = p.H >> 3; // uint4 per history row (<= 33)
        for (int r = 0; r < ngrp * kGroup; ++r) {
          const size_t c = (size_t)sg * KCH * kGroup + (size_t)r;
          if (c >= p.C) break;
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
    if (!isA) bulk_wait<0>();
  }
}

} // namespace v3

uint32_t chain_tile_samples() { return v3::TT; }

cudaError_t launch_chain_v3(const ChainParams &p_in, cudaStream_t stream, int variant, ChainLaunchInfo *info)
{
  using namespace v3;
  ChainParams p = p_in;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;

  p.NG = (p.C + kGroup - 1) / kGroup;
  p.NT = (p.L + TT - 1) / TT;
  p.n_items = p.NG * p.NT;
  p.TPS = p.S = 0;
  p.ablate = ((uint32_t)variant >> 4) & 3u; // variant bits 4,5: ablation study

  int smem_max = 0;
  e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  p.sets_in_smem = 1;
  size_t smem = smem_bytes(p.H, p.n_sets, p.set_stride_words);
  if (smem > (size_t)smem_max) {
    p.sets_in_smem = 0;
    smem = smem_bytes(p.H, 0, 0);
  }
  // variant bit 0: biquad products on DFMA (FP64 pipe) instead of IMAD.HI (integer pipe); with the chain warps alone on
  // their sub-partition the integer form has the shorter recurrence (tools/microbench/placement.cu)
  // channels per chain lane: 1.  The two-channel form (variant bit 3) raises a chain's own throughput by 56 % but doubles the
  // copy skeleton per sub-tile and loses to the one-channel form on both measured shapes (4096 x 64 blocks: 43 vs 81 Gsamples/s,
  // 65536 x 128 blocks: 93 vs 102), because the FIR warps are the other half of the bound.
  const int kch = (variant & 8) ? 2 : 1;
  auto kern = kch == 2 ? ((variant & 1) ? chain_kernel<2, BqStageD> : chain_kernel<2, BqStage>)
                       : ((variant & 1) ? chain_kernel<1, BqStageD> : chain_kernel<1, BqStage>);
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  // every CTA must be resident: chains spin on flags that FIR warps of other CTAs publish
  uint32_t grid = (uint32_t)sms;
  const uint32_t nsg = (p.NG + (uint32_t)kch - 1) / (uint32_t)kch;
  const uint32_t need = p.n_items > nsg ? p.n_items : nsg;
  if (grid > need) grid = need;
  p.W = grid;
  if (info) { info->grid = (int)grid; info->block = kThreads; info->smem = smem; info->tile = TT; }
  kern<<<grid, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_chain(const ChainParams &p, cudaStream_t stream, int variant, ChainLaunchInfo *info)
{
  return launch_chain_v3(p, stream, variant, info);
}

} // namespace msdr
