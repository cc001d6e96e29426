// placement.cu — does warp id % 4 select the SM sub-partition, and how much does an IMAD-saturating neighbour slow a
// latency-bound biquad chain warp?  13-warp CTAs, one per SM; "chain" warps run the 1-stage biquad recurrence,
// "load" warps run independent IMAD chains (FIR stand-in), others exit.
#include <cstdio>
#include "msdr_device.cuh"
using namespace msdr;
#define STEPS 4096

template <int F64>
__global__ void __launch_bounds__(416) k(long long *cyc, int *sink, unsigned chain_mask, unsigned load_mask, int seed)
{
  const int warp = threadIdx.x >> 5;
  const bool is_chain = (chain_mask >> warp) & 1u, is_load = (load_mask >> warp) & 1u;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (is_chain) {
    BqStage si; BqStageD sd; BqStageH sh;
    si.b0 = 236552419 + seed; si.b1 = 473104839; si.b2 = 236552419; si.a1 = 175469220; si.a2 = -47937074;
    si.x1 = si.x2 = si.y1 = si.y2 = 0; si.res = 0;
    bq_set_coefs(sd, si.b0, si.b1, si.b2, si.a1, si.a2);
    sd.x1 = sd.x2 = sd.y1 = sd.y2 = bq_d_from_int(0); sd.res = 0;
    bq_set_coefs(sh, si.b0, si.b1, si.b2, si.a1, si.a2);
    sh.x1 = sh.x2 = bq_d_from_int(0); sh.y1 = sh.y2 = 0; sh.res = 0;
    int x = (threadIdx.x * 977 + seed) & 0x7fff, acc = 0;
    long long t0 = clock64();
#pragma unroll 8
    for (int n = 0; n < STEPS; ++n) {
      x = (x * 75 + 74) & 0x7fff;
      if (F64 == 2) { acc ^= bq_step(sh, x); }
      else if (F64) { int y; bq_step(sd, bq_d_from_int(x), y); acc ^= y; }
      else { acc ^= bq_step(si, x << 16); }
    }
    long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) { cyc[blockIdx.x * 16 + warp] = t1 - t0; atomicAdd((int *)&done, 1); }
  } else if (is_load) {
    int acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    int a = seed + threadIdx.x, b = 7;
    const int nchain = __popc(chain_mask);
    while (done < nchain) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

template <int F64>
void run(const char *name, unsigned chain_mask, unsigned load_mask, long long *d_cyc, int *d_sink, int sms)
{
  cudaMemset(d_cyc, 0, sms * 16 * sizeof(long long));
  k<F64><<<sms, 416>>>(d_cyc, d_sink, chain_mask, load_mask, 1);
  cudaDeviceSynchronize();
  k<F64><<<sms, 416>>>(d_cyc, d_sink, chain_mask, load_mask, 2);
  cudaDeviceSynchronize();
  static long long h[148 * 16 + 64];
  cudaMemcpy(h, d_cyc, sms * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; int n = 0;
  for (int i = 0; i < sms * 16; ++i) if (h[i]) { avg += (double)h[i]; ++n; }
  printf("%-58s %s chain warps %08x load warps %08x : %6.1f cycles/step\n", name, F64 == 2 ? "hyb" : F64 ? "f64" : "int", chain_mask, load_mask, avg / n / STEPS);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long *d_cyc; int *d_sink;
  cudaMalloc(&d_cyc, sms * 16 * sizeof(long long)); cudaMalloc(&d_sink, (size_t)sms * 416 * sizeof(int));
  const unsigned fir9 = 0x0EEE;  // warps 1,2,3,5,6,7,9,10,11
#define BOTH(name, cm, lm) run<1>(name, cm, lm, d_cyc, d_sink, sms); run<0>(name, cm, lm, d_cyc, d_sink, sms); run<2>(name, cm, lm, d_cyc, d_sink, sms);
  BOTH("one chain warp (4), nothing else", 0x010, 0);
  BOTH("two chain warps 4,8 (same wid%4)", 0x110, 0);
  BOTH("two chain warps 4,5 (different wid%4)", 0x030, 0);
  BOTH("four chain warps 0,4,8,12 (same wid%4)", 0x1111, 0);
  BOTH("four chain warps 0,1,2,3", 0x000F, 0);
  BOTH("chains 4,8 + IMAD load on 1,2,3,5,6,7,9,10,11", 0x110, fir9);
  BOTH("chains 4,8 + IMAD load on 0,12 (same wid%4)", 0x110, 0x1001);
  BOTH("chains 4,8 + IMAD load on 1,5,9 only", 0x110, 0x0222);
  BOTH("chain 4 + IMAD load on all other 12 warps", 0x010, 0x1FEF);
  BOTH("chains 1,2 + IMAD load on 5,6,9,10 (same wid%4 as chains)", 0x006, 0x0660);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
