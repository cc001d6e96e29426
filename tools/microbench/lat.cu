// lat.cu — dependent-issue latency of the instructions on the biquad recurrence (one warp per SM, clock64 around an unrolled chain).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <vector>
#define ITERS 2048
#define U 8

template <int MODE>
__global__ void k(long long *cyc, int *sink, int a, int one, int sh, int zero)
{
  int d = threadIdx.x * 977 + a, c = a * 3 + 1, e = a + 5, f = a ^ 0x1234, g = a ^ 0x777, h = a ^ 0x999, i2 = a + 77;
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, dm = 0.999 + 1e-9 * one, da = 1e-3 * one;
  long long t0 = clock64();
  for (int n = 0; n < ITERS; ++n) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == 0) asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c));            // multiplicand -> result
      if (MODE == 1) asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(c));            // addend -> result
      if (MODE == 2) asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero));
      if (MODE == 3) asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh));
      if (MODE == 4) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(d) : "r"(one), "r"(c));
      if (MODE == 5) asm volatile("and.b32 %0, %0, %1;" : "+r"(d) : "r"(c));
      if (MODE == 6) { asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh)); asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero)); }
      if (MODE == 7) { asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh)); asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero));
                       asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c)); }
      if (MODE == 8) { asm volatile("max.s32 %0, %0, %1;" : "+r"(d) : "r"(c)); asm volatile("min.s32 %0, %0, %1;" : "+r"(d) : "r"(e)); }
      if (MODE == 9) { asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c));           // the cycle + 4 independent IMAD.HI
                       asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(a), "r"(e)); asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(c), "r"(e));
                       asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(e), "r"(a)); asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(a), "r"(a)); }
      if (MODE == 20) { // 4 independent IMAD.HI chains (throughput, one warp)
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c)); asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(f) : "r"(a), "r"(c));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(g) : "r"(a), "r"(c)); asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(h) : "r"(a), "r"(c)); }
      if (MODE == 21) { // 4 independent DFMA chains
        asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x0) : "d"(dm), "d"(da)); asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x1) : "d"(dm), "d"(da));
        asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x2) : "d"(dm), "d"(da)); asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x3) : "d"(dm), "d"(da)); }
      if (MODE == 22) asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x0) : "d"(dm), "d"(da)); // DFMA latency
      if (MODE == 23) { // 4 independent IMAD chains
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(d) : "r"(one), "r"(c)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(f) : "r"(one), "r"(c));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(g) : "r"(one), "r"(c)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(h) : "r"(one), "r"(c)); }
      if (MODE == 24) { // 2 IMAD + 2 SHF chains, independent (mixed pipes)
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(d) : "r"(one), "r"(c)); asm volatile("shr.s32 %0, %0, %1;" : "+r"(f) : "r"(zero));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(g) : "r"(one), "r"(c)); asm volatile("shr.s32 %0, %0, %1;" : "+r"(h) : "r"(zero)); }
      if (MODE == 25) { // 4 independent SHF chains
        asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(zero)); asm volatile("shr.s32 %0, %0, %1;" : "+r"(f) : "r"(zero));
        asm volatile("shr.s32 %0, %0, %1;" : "+r"(g) : "r"(zero)); asm volatile("shr.s32 %0, %0, %1;" : "+r"(h) : "r"(zero)); }
      if (MODE == 26) { // the y cycle (SHF, I2IP, IMAD.HI) + 1 independent IMAD.HI + 3 independent DFMA
        asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh)); asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c));
        asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(a), "r"(e));
        asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x0) : "d"(dm), "d"(da)); asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x1) : "d"(dm), "d"(da));
        asm volatile("fma.rm.f64 %0, %1, %0, %2;" : "+d"(x2) : "d"(dm), "d"(da)); }
      if (MODE == 27) { // the y cycle + 1 independent IMAD.HI
        asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh)); asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c));
        asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(f) : "r"(a), "r"(e)); }
      if (MODE == 28) { // the y cycle + 4 independent IMAD.HI
        asm volatile("shr.s32 %0, %0, %1;" : "+r"(d) : "r"(sh)); asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(d) : "r"(zero));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(d) : "r"(a), "r"(c));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(f) : "r"(a), "r"(e)); asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(g) : "r"(a), "r"(e));
        asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(h) : "r"(a), "r"(e)); asm volatile("mad.hi.s32 %0, %1, %0, %2;" : "+r"(i2) : "r"(a), "r"(e)); }
      if (MODE >= 30 && MODE <= 35) { // the biquad step built up piece by piece: d = sum, f = ys (y << 16), g = y2s, h = res
        int pre = c, e2 = e;
        if (MODE >= 33) { // three feed-forward products of a changing input i2
          i2 = i2 * 1664525 + 1013904223;
          if (MODE == 35) { // on the FP64 pipe
            const double xD = __hiloint2double(0x41310000 + (i2 >> 16), 0);
            e2 = __double2loint(__fma_rd(dm, xD, da)) + __double2loint(__fma_rd(dm, x1, da)) + __double2loint(__fma_rd(dm, x2, da));
            x2 = x1; x1 = xD;
          } else {
            int t0, t1, t2;
            asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(t0) : "r"(a), "r"(i2));
            asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(t1) : "r"(c), "r"(i2 ^ 0x55));
            asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(t2) : "r"(e), "r"(i2 ^ 0x77));
            e2 = t0 + t1 + t2;
          }
        }
        if (MODE >= 32) asm volatile("mad.hi.s32 %0, %1, %2, %3;" : "=r"(e2) : "r"(c), "r"(g), "r"(e2)); // a2 * y2 + e
        if (MODE == 31) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(pre) : "r"(h), "r"(one), "r"(c));
        if (MODE >= 32) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(pre) : "r"(h), "r"(one), "r"(e2));
        asm volatile("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(f), "r"(pre)); // sum = a1 * y1 + pre
        g = f;
        int sh14;
        asm volatile("shr.s32 %0, %1, 14;" : "=r"(sh14) : "r"(d));
        asm volatile("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(f) : "r"(sh14), "r"(zero));
        if (MODE >= 31) asm volatile("and.b32 %0, %1, 0x3fff;" : "=r"(h) : "r"(d));
      }
      if (MODE == 11) asm volatile("cvt.sat.s16.s32 %0, %0;" : "+r"(d));
      if (MODE == 12) { long long w; asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(w) : "r"(a), "r"(d)); d = (int)(w >> 32) + c; }
      if (MODE == 13) asm volatile("mul.hi.s32 %0, %1, %0;" : "+r"(d) : "r"(a));
      if (MODE == 14) asm volatile("mul.hi.u32 %0, %1, %0;" : "+r"(d) : "r"(a));
      if (MODE == 15) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(c), "r"(e));
    }
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = d + f + g + h + i2 + (int)(x0 + x1 + x2 + x3);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, long long *d_cyc, int *d_sink)
{
  k<MODE><<<8, 32>>>(d_cyc, d_sink, 12345, 1, 14, 0); cudaDeviceSynchronize();
  k<MODE><<<8, 32>>>(d_cyc, d_sink, 54321, 1, 14, 0); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 8; ++i) avg += (double)h[i]; avg /= 8;
  printf("%-60s %6.2f cycles per link\n", name, avg / ((double)ITERS * U));
}

int main()
{
  long long *d_cyc; int *d_sink; cudaMalloc(&d_cyc, 64 * 8); cudaMalloc(&d_sink, 8 * 32 * 4);
  run<0>("IMAD.HI  multiplicand -> result", d_cyc, d_sink);
  run<1>("IMAD.HI  addend -> result", d_cyc, d_sink);
  run<13>("mul.hi.s32 (no addend)", d_cyc, d_sink);
  run<14>("mul.hi.u32 (no addend)", d_cyc, d_sink);
  run<12>("mul.wide.s32 + hi word + add", d_cyc, d_sink);
  run<2>("I2IP.S16.S32.SAT (cvt.pack.sat)", d_cyc, d_sink);
  run<11>("cvt.sat.s16.s32", d_cyc, d_sink);
  run<3>("SHF (shr.s32 by register)", d_cyc, d_sink);
  run<4>("IMAD (mad.lo)", d_cyc, d_sink);
  run<5>("LOP3 (and)", d_cyc, d_sink);
  run<15>("PRMT", d_cyc, d_sink);
  run<8>("VIMNMX max + min (2 links)", d_cyc, d_sink);
  run<6>("SHF + I2IP (2 links)", d_cyc, d_sink);
  run<7>("SHF + I2IP + IMAD.HI (3 links: the y cycle)", d_cyc, d_sink);
  run<9>("IMAD.HI link + 4 independent IMAD.HI", d_cyc, d_sink);
  run<20>("4 independent IMAD.HI (per group of 4)", d_cyc, d_sink);
  run<21>("4 independent DFMA (per group of 4)", d_cyc, d_sink);
  run<22>("DFMA latency", d_cyc, d_sink);
  run<23>("4 independent IMAD (per group)", d_cyc, d_sink);
  run<24>("2 IMAD + 2 SHF independent (per group)", d_cyc, d_sink);
  run<25>("4 independent SHF (per group)", d_cyc, d_sink);
  run<27>("y cycle + 1 independent IMAD.HI", d_cyc, d_sink);
  run<26>("y cycle + 1 independent IMAD.HI + 3 DFMA", d_cyc, d_sink);
  run<28>("y cycle + 4 independent IMAD.HI", d_cyc, d_sink);
  run<30>("step A: y cycle only (IMAD.HI a1, SHF, I2IP)", d_cyc, d_sink);
  run<31>("step B: A + residual (LOP3, opaque IMAD)", d_cyc, d_sink);
  run<32>("step C: B + IMAD.HI a2*y2", d_cyc, d_sink);
  run<33>("step D: C + 3 feed-forward IMAD.HI (the full step)", d_cyc, d_sink);
  run<35>("step E: C + 3 feed-forward DFMA (hybrid step)", d_cyc, d_sink);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
