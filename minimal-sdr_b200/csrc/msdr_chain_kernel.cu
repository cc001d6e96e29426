// msdr_chain_kernel.cu — K1: the fused receive chain for sm_100a.
//
//   int16 IF samples -> [fs/4 mix folded into tap selection] -> FIR pair (Q15, 32-bit wrapping accumulate)
//   -> SSB sum / AM envelope -> biquad cascade (Q2.30 x int16, 14-bit error feedback) -> int16 audio
//
// Reference semantics: Minimal-SDR.ino:546-558 (mix), arm_fir_fast_q15.c:60-329 (FIR), Minimal-SDR.ino:589-628
// (demod), filter_biquad.cpp:33-82 (biquad).  Bit-exact by construction, see DESIGN.md "Exactness".
//
// Structure (one persistent CTA per SM, warp-specialised, mbarrier pipelines):
//
//   warp 0        producer  claims (channel-group, time-segment) work items from a global counter and streams
//                           their tiles HBM -> shared memory with cp.async.bulk (TMA engine), one 1-D bulk
//                           copy per channel row covering the (T-1)-sample halo plus the tile.
//   warps 1..NF   FIR       fold the fs/4 oscillator sign into the samples in place, then each warp takes one
//                           channel row at a time: lane l computes 2R consecutive outputs of all four
//                           polyphase sub-filters (I/Q x even/odd output phase) from a sliding register window,
//                           taps broadcast from shared memory; epilogue does >>15, SSAT, demod; result goes to
//                           a second shared buffer.
//   warp NF+1     biquad    lane = channel: serial recurrence over the tile in shared memory, in place, state in
//                           registers for the whole segment; then one bulk copy per row shared -> HBM.
//
// The only serial dependence across time is the biquad state.  Work items are ordered segment-major and the
// biquad warp hands its state to the CTA that owns the next segment of the same channel group through global
// memory + a release/acquire flag, so FIR work balances over all SMs for any channel count.
#include "msdr_chain_common.cuh"

namespace msdr {

struct TileDesc {
  int grp;   // channel group
  int seg;   // segment of the launch this tile belongs to
  int t0;    // first sample of the tile (within the launch)
  int len;   // samples (multiple of 128, <= TT)
  uint32_t flags;
  int pad[3];
};
enum : uint32_t { TF_SEG_FIRST = 1u, TF_SEG_LAST = 2u, TF_LAUNCH_LAST = 4u, TF_END = 0x80000000u };

struct __align__(16) PipeCtrl {
  uint64_t full[2];   // producer -> FIR : raw tile landed (tx bytes)
  uint64_t empty[2];  // FIR -> producer : raw tile consumed
  uint64_t dfull[2];  // FIR -> biquad   : demodulated tile ready
  uint64_t dfree[2];  // biquad -> FIR   : demodulated tile buffer drained to HBM
  TileDesc desc[2];
  TileDesc ddesc[2];
  int job_ctr[2];
  int pad[2];
  uint32_t rowinfo[2][kGroup]; // per row: setid | mode << 8 | kp4 << 16
};
static_assert(sizeof(PipeCtrl) <= 512, "PipeCtrl must fit its smem slot");
constexpr uint32_t kCtrlBytes = 512;

size_t chain_smem_bytes(uint32_t H, uint32_t n_sets, uint32_t set_stride_words, int tile)
{
  const uint32_t sets_bytes = align_up(n_sets * set_stride_words * 4u, 128u);
  const uint32_t raw_stage = kGroup * (H + (uint32_t)tile) * 2u;
  const uint32_t d_stage = kGroup * ((uint32_t)tile / 2u + 4u) * 4u;
  return (size_t)kCtrlBytes + sets_bytes + 2u * raw_stage + 2u * d_stage;
}

template <int TT, int NF, class BQ>
__global__ void __launch_bounds__((NF + 2) * 32, 1) chain_kernel(const ChainParams p)
{
  constexpr int R = TT / 64;      // output pairs per lane
  constexpr int DW = TT / 2 + 4;  // demod-buffer row pitch in words (== 4 mod 32: conflict-free LDS.128 by row)
  extern __shared__ __align__(128) unsigned char smem[];
  PipeCtrl *pc = reinterpret_cast<PipeCtrl *>(smem);
  int32_t *s_sets = reinterpret_cast<int32_t *>(smem + kCtrlBytes);
  const uint32_t sets_words = p.n_sets * p.set_stride_words;
  const uint32_t RS = p.H + TT; // raw row pitch, samples
  unsigned char *raw_base = smem + kCtrlBytes + align_up(sets_words * 4u, 128u);
  const uint32_t raw_stage = kGroup * RS * 2u;
  unsigned char *d_base = raw_base + 2u * raw_stage;
  constexpr uint32_t d_stage = kGroup * DW * 4u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pc->full[s], 1);
      mbar_init(&pc->empty[s], NF);
      mbar_init(&pc->dfull[s], NF);
      mbar_init(&pc->dfree[s], 1);
    }
    mbar_fence_init();
  }
  for (uint32_t i = threadIdx.x; i < sets_words; i += blockDim.x) s_sets[i] = p.sets[i];
  __syncthreads();

  const int Hw = (int)(p.H >> 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t it = 0;
    for (;;) {
      int item = 0;
      if (lane == 0) item = atomicAdd(&p.ctrl[0], 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= (int)p.n_items) break;
      const int seg = item / (int)p.NG, grp = item - seg * (int)p.NG; // segment-major order
      const int tile_begin = seg * (int)p.TPS;
      const int tile_end = min(tile_begin + (int)p.TPS, (int)p.NT);
      const int nrows = min(kGroup, (int)p.C - grp * kGroup);
      const uint32_t row = (uint32_t)(grp * kGroup + lane); // row of in/out
      const uint32_t ch = p.ch0 + row;                      // channel of the chain
      uint32_t rinfo = 0;
      if (lane < nrows) {
        const uint32_t set = p.setid[ch];
        rinfo = set | ((uint32_t)p.mode[ch] << 8) | (p.set_kp4[set] << 16);
      }
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1u;
        mbar_wait(&pc->empty[s], ph ^ 1u);
        const int t0 = tile * TT;
        const int len = min(TT, (int)p.L - t0);
        pc->rowinfo[s][lane] = rinfo;
        const uint32_t row_bytes = (p.H + (uint32_t)len) * 2u;
        if (lane == 0) {
          TileDesc td;
          td.grp = grp; td.seg = seg; td.t0 = t0; td.len = len;
          td.flags = (tile == tile_begin ? TF_SEG_FIRST : 0u) | (tile + 1 == tile_end ? TF_SEG_LAST : 0u) |
                     (tile + 1 == (int)p.NT ? TF_LAUNCH_LAST : 0u);
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
  } else if (warp <= NF) {
    // ------------------------------------------------------------------ FIR + demod warps
    const int ftid = (warp - 1) * 32 + lane;
    for (uint32_t it = 0;; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->full[s], ph);
      const TileDesc td = pc->desc[s];
      if (td.flags & TF_END) {
        mbar_wait(&pc->dfree[s], ph ^ 1u);
        if (ftid == 0) pc->ddesc[s] = td;
        __syncwarp();
        if (lane == 0) mbar_arrive(&pc->dfull[s]);
        break;
      }
      const int nrows = min(kGroup, (int)p.C - td.grp * kGroup);
      uint32_t *rawW = reinterpret_cast<uint32_t *>(raw_base + (uint32_t)s * raw_stage);
      // fold the fs/4 oscillator sign: samples with n % 4 in {2,3} are negated (Minimal-SDR.ino:550,555),
      // i.e. every odd word of the row
      {
        const int q4 = (int)((p.H + (uint32_t)td.len) >> 3); // uint4 per row
        for (int r = warp - 1; r < nrows; r += NF) {
          uint4 *pw = reinterpret_cast<uint4 *>(rawW + (uint32_t)r * (RS >> 1));
          for (int c4 = lane; c4 < q4; c4 += 32) {
            uint4 v = pw[c4];
            v.y = neg16x2(v.y);
            v.w = neg16x2(v.w);
            pw[c4] = v;
          }
        }
      }
      named_bar_sync(1, NF * 32);
      mbar_wait(&pc->dfree[s], ph ^ 1u);
      if (ftid == 0) pc->ddesc[s] = td;
      uint32_t *dW = reinterpret_cast<uint32_t *>(d_base + (uint32_t)s * d_stage);
      for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(&pc->job_ctr[s], 1);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= nrows) break;
        const uint32_t ri = pc->rowinfo[s][job];
        const uint32_t set = ri & 0xFFu;
        const int kind = demod_kind_of((int)((ri >> 8) & 0xFFu), p.am_q31);
        const int4 *cf = reinterpret_cast<const int4 *>(s_sets + set * p.set_stride_words);
        fir_demod_row<R>(rawW + (uint32_t)job * (RS >> 1), cf, (int)(ri >> 16), Hw, lane, td.len, kind, dW + (uint32_t)job * DW);
      }
      fence_proxy_async_smem(); // raw[s] was rewritten through the generic proxy; its next writer is the TMA engine
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&pc->dfull[s]);
        mbar_arrive(&pc->empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ biquad warp (lane = channel)
    int *flags = p.ctrl + 1;
    BQ st[2];
    uint32_t fl[2] = {0u, 0u};
    int nst0 = 1, nst1 = 1;
    bool fast = true;
    for (uint32_t it = 0;; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(&pc->dfull[s], ph);
      const TileDesc td = pc->ddesc[s];
      if (td.flags & TF_END) break;
      const uint32_t row = (uint32_t)(td.grp * kGroup + lane);
      const uint32_t ch = p.ch0 + row;
      const bool active = row < p.C;
      uint32_t *drow = reinterpret_cast<uint32_t *>(d_base + (uint32_t)s * d_stage) + (uint32_t)lane * DW;

      if (td.flags & TF_SEG_FIRST) {
        if (td.seg > 0) { // state hand-off from the CTA that ran the previous segment of this group
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flags + td.grp) < td.seg) {
              __nanosleep(64);
              if (clock64() - t0 > kWatchdogCycles) __trap();
            }
          }
          __syncwarp();
        }
        nst0 = nst1 = 1;
        if (active) {
          // stages run while bit31 of word 7 says another follows (filter_biquad.cpp:75,79)
          for (int k = 0; k < 3 && (nst0 == k + 1); ++k)
            if ((uint32_t)__ldcg(p.bq + (size_t)((0 * 4 + k) * 8 + 7) * p.Cpad + ch) & 0x80000000u) nst0 = k + 2;
          for (int k = 0; k < 3 && (nst1 == k + 1); ++k)
            if ((uint32_t)__ldcg(p.bq + (size_t)((1 * 4 + k) * 8 + 7) * p.Cpad + ch) & 0x80000000u) nst1 = k + 2;
        }
        fast = __all_sync(0xffffffffu, nst0 == 1 && nst1 == 1);
        if (fast && active) {
          bq_load_stage(st[0], fl[0], p.bq, p.Cpad, 0, 0, ch);
          bq_load_stage(st[1], fl[1], p.bq, p.Cpad, 1, 0, ch);
        }
      }

      if (fast) {
        if (active) {
          uint4 *dp = reinterpret_cast<uint4 *>(drow);
#pragma unroll 1
          for (int q = 0; q < (td.len >> 3); ++q) {
            uint4 v = dp[q];
            v.x = bq_word<2>(st, v.x);
            v.y = bq_word<2>(st, v.y);
            v.z = bq_word<2>(st, v.z);
            v.w = bq_word<2>(st, v.w);
            dp[q] = v;
          }
        }
      } else {
        // generic cascade: stage-major over the tile like the reference (filter_biquad.cpp:44-79)
        for (int obj = 0; obj < 2; ++obj) {
          const int myn = obj ? nst1 : nst0;
          for (int k = 0; k < 4; ++k) {
            const bool has = active && k < myn;
            if (!__any_sync(0xffffffffu, has)) break;
            if (has) {
              BQ g[1];
              uint32_t gf;
              bq_load_stage(g[0], gf, p.bq, p.Cpad, obj, k, ch);
              uint4 *dp = reinterpret_cast<uint4 *>(drow);
#pragma unroll 1
              for (int q = 0; q < (td.len >> 3); ++q) {
                uint4 v = dp[q];
                v.x = bq_word<1>(g, v.x);
                v.y = bq_word<1>(g, v.y);
                v.z = bq_word<1>(g, v.z);
                v.w = bq_word<1>(g, v.w);
                dp[q] = v;
              }
              bq_store_stage(g[0], gf, p.bq, p.Cpad, obj, k, ch);
            }
          }
        }
      }

      // drain the finished tile to HBM: one bulk copy per channel row
      fence_proxy_async_smem();
      if (active) bulk_s2g(p.out + (size_t)row * p.stride + (size_t)td.t0, drow, (uint32_t)td.len * 2u);
      bulk_commit();
      // the buffer is reusable as soon as the copy engine has READ it (not when the data has landed in HBM)
      bulk_wait_read<0>();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pc->dfree[s]);

      if (td.flags & TF_SEG_LAST) {
        if (fast && active) {
          bq_store_stage(st[0], fl[0], p.bq, p.Cpad, 0, 0, ch);
          bq_store_stage(st[1], fl[1], p.bq, p.Cpad, 1, 0, ch);
        }
        if (td.flags & TF_LAUNCH_LAST) {
          // carry the last H raw samples: hist <- tail of (hist || in[0..L))
          const int nrows = min(kGroup, (int)p.C - td.grp * kGroup);
          const uint32_t hq = p.H >> 3; // uint4 per history row (<= 33)
          for (int r = 0; r < nrows; ++r) {
            const size_t c = (size_t)(td.grp * kGroup + r);
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
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(flags + td.grp, td.seg + 1);
      }
    }
    bulk_wait<0>();
  }
}

cudaError_t launch_chain_handoff(const ChainParams &p_in, cudaStream_t stream, int variant, ChainLaunchInfo *info)
{
  constexpr int TT = 512, NF = 8;
  ChainParams p = p_in;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;

  p.NG = (p.C + kGroup - 1) / kGroup;
  p.NT = (p.L + TT - 1) / TT;
  // enough work items to balance the SMs, never finer than one tile
  uint32_t S = (8u * (uint32_t)sms + p.NG - 1) / p.NG;
  if (S > p.NT) S = p.NT;
  if (S < 1) S = 1;
  p.TPS = (p.NT + S - 1) / S;
  p.S = (p.NT + p.TPS - 1) / p.TPS;
  p.n_items = p.NG * p.S;

  const size_t smem = chain_smem_bytes(p.H, p.n_sets, p.set_stride_words, TT);
  // variant bit 0: biquad products on IMAD.HI (integer pipe) instead of DFMA (FP64 pipe)
  auto kern = (variant & 1) ? chain_kernel<TT, NF, BqStageD> : chain_kernel<TT, NF, BqStage>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (NF + 2) * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  // every CTA must be resident: the biquad hand-off spins on flags owned by other CTAs
  uint32_t grid = (uint32_t)sms * (uint32_t)per_sm;
  if (grid > p.n_items) grid = p.n_items;
  if (info) { info->grid = (int)grid; info->block = (NF + 2) * 32; info->smem = smem; info->tile = TT; }
  kern<<<grid, (NF + 2) * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

} // namespace msdr
