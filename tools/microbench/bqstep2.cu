// bqstep2.cu — shorter recurrence for the biquad sample-step (one warp, isolated), verified against msdr::bq_step.
// The step's critical path is  y[n-1] -> a1 product (+ early + res) -> >>14 -> clamp -> <<16.
//   mode 0  msdr::bq_step as shipped: residual added through an opaque IMAD, so the loop-carried cycles are
//           IMAD.HI(a1) -> SHF -> I2IP  and  IMAD.HI(a1) -> LOP3 -> IMAD
//   mode 1  P : the same arithmetic with the residual added by a plain add: ptxas seeds ONE IMAD.HI chain through all five
//           products with it (cycle = LOP3 + 5 dependent IMAD.HI)
//   mode 7  H : msdr::bq_step(BqStageH): feed-forward products as exact DFMA.RM, a1/a2 products IMAD.HI
//   mode 2  Q : a1 product split into 16-bit halves, no IMAD.HI on the path            SHF, VIMNMX x2, IMAD, SHF, IADD
//   mode 3  Q': same with cvt.sat.s16.s32 for the clamp                                SHF, I2I.SAT, IMAD, SHF, IADD
//   mode 4  P': like P but the >>14 folded: clamp(sum, -2^29, 2^29-1) has no cheap single op; uses shf + pack (same as 1) with
//           residual/feed-forward chain on plain IMAD halves (all five products split)   -- integer pipe relief only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../minimal-sdr_b200/csrc -o bqstep2 bqstep2.cu
#include <cstdio>
#include <vector>
#include "msdr_device.cuh"
using namespace msdr;

#define STEPS 8192

__device__ __forceinline__ int pack_hi_sat(int v)
{
  int d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0)); // upper half <- sat16(v), lower half <- 0
  return d;
}
__device__ __forceinline__ int sat16_cvt(int v)
{
  int d;
  asm("cvt.sat.s16.s32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}

struct StQ { // split coefficients, plain-int history
  int b0, b1, b2, a1, a2, a1h, a1l;
  int x1, x2, y1s, y2s, y1; // x*, y*s: << 16 ; y1: plain
  int res;
};

template <int MODE>
__device__ __forceinline__ int step(BqStage &s, StQ &q, int xs)
{
  if (MODE == 0) return bq_step(s, xs);
  if (MODE == 1) {
    int e;
    asm("mul.hi.s32 %0, %1, %2;" : "=r"(e) : "r"(s.b0), "r"(xs));
    asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(s.b1), "r"(s.x1));
    asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(s.b2), "r"(s.x2));
    asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(s.a2), "r"(s.y2));
    const int pre = e + s.res;
    const int sum = smlaw_s(pre, s.a1, s.y1);
    const int ys = pack_hi_sat(sum >> 14);
    s.res = sum & 0x3FFF;
    s.x2 = s.x1; s.x1 = xs;
    s.y2 = s.y1; s.y1 = ys;
    return ys;
  }
  // MODE 2, 3
  int e;
  asm("mul.hi.s32 %0, %1, %2;" : "=r"(e) : "r"(q.b0), "r"(xs));
  asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(q.b1), "r"(q.x1));
  asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(q.b2), "r"(q.x2));
  asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(e) : "r"(q.a2), "r"(q.y2s));
  const int pre = e + q.res;
  const int t = q.a1l * q.y1;            // 16 x 16 bit, exact
  const int pre2 = q.a1h * q.y1 + pre;   // wraps like the reference's 32-bit add
  const int sum = pre2 + (t >> 16);
  const int y = MODE == 2 ? ssat16(sum >> 14) : sat16_cvt(sum >> 14);
  q.res = sum & 0x3FFF;
  q.x2 = q.x1; q.x1 = xs;
  q.y2s = q.y1s; q.y1s = y << 16; q.y1 = y;
  return y << 16;
}

// recurrence with the feed-forward sum e supplied: no IMAD.HI at all (a1, a2 products split into 16-bit halves)
struct StR { int a1h, a1l, a2h, a2l; int y1, y2, res; };
__device__ __forceinline__ int step_rec(StR &r, int e)
{
  const int t2 = r.a2l * r.y2;                     // off the critical path: y[n-2] is one step old
  const int pre = r.a2h * r.y2 + e + (t2 >> 16);
  const int t1 = r.a1l * r.y1;
  const int p1 = r.a1h * r.y1 + pre;
  const int sum = p1 + (t1 >> 16) + r.res;
  const int y = ssat16(sum >> 14);
  r.res = sum & 0x3FFF;
  r.y2 = r.y1; r.y1 = y;
  return y;
}
__device__ __forceinline__ int step_rec_hi(BqStage &s, int e) // same with IMAD.HI (what msdr::bq_step does per sample)
{
  int pre;
  asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(pre) : "r"(s.a2), "r"(s.y2), "r"(e));
  pre += s.res;
  const int sum = smlaw_s(pre, s.a1, s.y1);
  const int y = ssat16(sum >> 14);
  s.res = sum & 0x3FFF;
  s.y2 = s.y1; s.y1 = y << 16;
  return y;
}

// recurrence with the feed-forward sum supplied, residual kept off the IMAD.HI chain (what msdr::rec_step does)
struct StB { int a1, a2, y1, y2, res; };
__device__ __forceinline__ int step_rec_b(StB &r, int e)
{
  int t, pre;
  asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(r.a2), "r"(r.y2), "r"(e));
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(pre) : "r"(r.res), "r"(kBqOne), "r"(t));
  const int sum = smlaw_s(pre, r.a1, r.y1);
  int ys;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(ys) : "r"(sum >> 14), "r"(0));
  r.res = sum & 0x3FFF;
  r.y2 = r.y1; r.y1 = ys;
  return ys;
}

template <int MODE>
__global__ void k(long long *cyc, int *sink, int seed, int amp, int lanes)
{
  if ((int)threadIdx.x >= lanes) return;
  BqStage s;
  s.b0 = 236552419 + seed; s.b1 = 473104839; s.b2 = 236552419; s.a1 = 175469220 * 5; s.a2 = -47937074 * 9;
  s.x1 = s.x2 = s.y1 = s.y2 = 0; s.res = 0;
  StQ q;
  q.b0 = s.b0; q.b1 = s.b1; q.b2 = s.b2; q.a1 = s.a1; q.a2 = s.a2; q.a1h = s.a1 >> 16; q.a1l = s.a1 & 0xFFFF;
  q.x1 = q.x2 = q.y1s = q.y2s = q.y1 = 0; q.res = 0;
  StR r;
  r.a1h = s.a1 >> 16; r.a1l = s.a1 & 0xFFFF; r.a2h = s.a2 >> 16; r.a2l = s.a2 & 0xFFFF; r.y1 = r.y2 = r.res = 0;
  BqStageH sh;
  bq_set_coefs(sh, s.b0, s.b1, s.b2, s.a1, s.a2);
  sh.x1 = sh.x2 = bq_d_from_int(0); sh.y1 = sh.y2 = 0; sh.res = 0;
  int fx1 = 0, fx2 = 0;
  uint32_t x = threadIdx.x * 977u + seed;
  int acc = 0;
  StB rb; rb.a1 = s.a1; rb.a2 = s.a2; rb.y1 = rb.y2 = rb.res = 0;
  long long t0 = clock64();
  if (MODE == 10) { // batched hybrid: the feed-forward sums of 8 samples first (DFMA), then their 8 recurrence steps
#pragma unroll 1
    for (int n = 0; n < STEPS; n += 8) {
      int e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x = x * 1664525u + 1013904223u;
        const int xin = ((int)x >> 16) >> amp;
        const double xD = bq_d_from_int(xin);
        e[j] = __double2loint(__fma_rd(sh.b0, xD, sh.m0)) + __double2loint(__fma_rd(sh.b1, sh.x1, sh.m1)) + __double2loint(__fma_rd(sh.b2, sh.x2, sh.m2));
        sh.x2 = sh.x1; sh.x1 = xD;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = acc * 31 + (step_rec_b(rb, e[j]) >> 16);
    }
  } else
#pragma unroll 8
  for (int n = 0; n < STEPS; ++n) {
    x = x * 1664525u + 1013904223u;
    const int xin = ((int)x >> 16) >> amp; // signed 16-bit, scaled
    int v;
    if (MODE <= 3) v = step<MODE>(s, q, xin << 16) >> 16;
    if (MODE == 4) { // verification: feed-forward sum computed here the way a helper warp would, recurrence split
      const int xs = xin << 16;
      const int e = __mulhi(s.b0, xs) + __mulhi(s.b1, fx1) + __mulhi(s.b2, fx2);
      fx2 = fx1; fx1 = xs;
      v = step_rec(r, e);
    }
    if (MODE == 7) v = bq_step(sh, xin) >> 16;        // hybrid: feed-forward on DFMA, recurrence on IMAD.HI
    if (MODE == 5) v = step_rec(r, (int)x >> 3);      // timing: e supplied
    if (MODE == 6) v = step_rec_hi(s, (int)x >> 3);   // timing: e supplied, IMAD.HI recurrence
    acc = acc * 31 + v;
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, long long *d_cyc, int *d_sink, int sms, std::vector<int> *ref, int lanes = 32)
{
  for (int amp : {0, 6}) { // full-scale input saturates the output regularly; amp 64 never does
    k<MODE><<<sms, 32>>>(d_cyc, d_sink, 1, amp, lanes);
    cudaDeviceSynchronize();
    cudaMemset(d_sink, 0, (size_t)sms * 32 * sizeof(int));
    k<MODE><<<sms, 32>>>(d_cyc, d_sink, 2, amp, lanes);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    std::vector<int> out((size_t)sms * 32);
    cudaMemcpy(h.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), d_sink, out.size() * sizeof(int), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    std::vector<int> &r = ref[amp == 0 ? 0 : 1];
    const char *verdict = "reference";
    if (MODE == 0 && lanes == 32) r = out; else if (lanes == 32) verdict = (r == out) ? "identical to reference" : "MISMATCH"; else verdict = "";
    printf("%-52s lanes %2d  >>%d  cycles/step %6.1f   %s\n", name, lanes, amp, avg / STEPS, verdict);
  }
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long *d_cyc; int *d_sink;
  cudaMalloc(&d_cyc, 1024 * sizeof(long long)); cudaMalloc(&d_sink, (size_t)sms * 1024 * sizeof(int));
  std::vector<int> ref[2];
  run<0>("0 bq_step (residual via opaque IMAD)", d_cyc, d_sink, sms, ref);
  run<1>("1 P  (plain add: 5 IMAD.HI serialised)", d_cyc, d_sink, sms, ref);
  run<2>("2 Q  (SHF,VIMNMX,VIMNMX,IMAD,SHF,IADD)", d_cyc, d_sink, sms, ref);
  run<3>("3 Q' (SHF,I2I.SAT,IMAD,SHF,IADD)", d_cyc, d_sink, sms, ref);
  run<4>("4 split recurrence, feed-forward in loop (verify)", d_cyc, d_sink, sms, ref);
  run<10>("10 hybrid, batched: 8 feed-forward sums, then 8 recurrence steps", d_cyc, d_sink, sms, ref);
  run<7>("7 hybrid: 3 DFMA feed-forward + 2 IMAD.HI recurrence", d_cyc, d_sink, sms, ref);
  { std::vector<int> dummy[2]; dummy[0] = dummy[1] = std::vector<int>();
    run<5>("5 split recurrence, e supplied (timing only)", d_cyc, d_sink, sms, dummy, 31);
    run<6>("6 IMAD.HI recurrence, e supplied (timing only)", d_cyc, d_sink, sms, dummy, 31); }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
