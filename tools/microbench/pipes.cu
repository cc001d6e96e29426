// pipes.cu — issue-rate microbenchmarks for the instruction mix of the receive chain on B200 (sm_100a).
// Measures warp-instructions per clock per SM for IMAD / IMAD.HI / IMAD.WIDE / DFMA / FFMA / IDP / ALU ops, selected
// mixes (does DFMA or FFMA co-issue with IMAD?), and legacy mma.sync s8.  Results feed DESIGN.md's INT-issue roofline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define NCH 16   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024) k(long long *cyc, int *sink, int a0, int b0)
{
  int acc[NCH];
  double dacc[NCH];
  float facc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { acc[i] = threadIdx.x + i; dacc[i] = 1.0 + i; facc[i] = 1.0f + i; }
  int a = a0 + threadIdx.x, b = b0;
  double da = 1.0000001, db = 0.5;
  float fa = 1.0000001f, fb = 0.5f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
      if (MODE == 1) asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
      if (MODE == 2) { long long w; asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(w) : "r"(acc[i]), "r"(b)); acc[i] = (int)(w >> 32) ^ (int)w; }
      if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[i]) : "d"(da), "d"(db));
      if (MODE == 4) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[i]) : "d"(da), "d"(db)); }
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(fa), "f"(fb));
      if (MODE == 6) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(facc[i]) : "f"(fa), "f"(fb)); }
      if (MODE == 7) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
      if (MODE == 8) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
      if (MODE == 9) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[(i + 8) % NCH]) : "r"(a), "r"(b)); }
      if (MODE == 10) asm volatile("max.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a));
      if (MODE == 11) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(acc[i]) : "r"(a));
      if (MODE == 12) asm volatile("shr.s32 %0, %0, 3;" : "+r"(acc[i]));
      if (MODE == 13) { asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[(i + 8) % NCH]) : "r"(a), "r"(b)); }
      if (MODE == 14) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(acc[(i + 8) % NCH]) : "r"(a), "r"(b)); }
      if (MODE == 15) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b)); asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[(i + 8) % NCH]) : "r"(a), "r"(b)); }
    }
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i] + (int)dacc[i] + (int)facc[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// legacy tensor path: mma.sync m16n8k32 s8, 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(1024) k_mma(long long *cyc, int *sink, int a0)
{
  int c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
  unsigned a[4] = {(unsigned)a0, (unsigned)a0 + 1, (unsigned)a0 + 2, (unsigned)a0 + 3}, b[2] = {(unsigned)a0 * 3, (unsigned)a0 * 5};
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_iter_per_chain, int threads, long long *d_cyc, int *d_sink, int sms)
{
  k<MODE><<<sms, threads>>>(d_cyc, d_sink, 3, 7);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms, threads>>>(d_cyc, d_sink, 3, 7);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024];
  cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
  const double warp_instr = (double)(threads / 32) * ITER * NCH * instr_per_iter_per_chain;
  printf("%-28s threads/SM %4d  cycles %10.0f  warp-instr/clk/SM %6.3f  lane-ops/clk/SM %7.1f  (%.3f ms, eff clock %.0f MHz)\n", name, threads, avg,
         warp_instr / avg, 32.0 * warp_instr / avg, ms, avg / (ms * 1e3));
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, sms);
  long long *d_cyc; int *d_sink;
  cudaMalloc(&d_cyc, 1024 * sizeof(long long)); cudaMalloc(&d_sink, (size_t)sms * 1024 * sizeof(int));
  for (int threads : {256, 512, 1024}) {
    run<0>("IMAD (mad.lo.s32)", 1, threads, d_cyc, d_sink, sms);
    run<1>("IMAD.HI (mad.hi.s32)", 1, threads, d_cyc, d_sink, sms);
    run<2>("IMAD.WIDE+xor", 2, threads, d_cyc, d_sink, sms);
    run<3>("DFMA", 1, threads, d_cyc, d_sink, sms);
    run<4>("IMAD + DFMA 1:1", 2, threads, d_cyc, d_sink, sms);
    run<5>("FFMA", 1, threads, d_cyc, d_sink, sms);
    run<6>("IMAD + FFMA 1:1", 2, threads, d_cyc, d_sink, sms);
    run<7>("IDP.4A (dp4a)", 1, threads, d_cyc, d_sink, sms);
    run<8>("IDP.2A (dp2a.lo)", 1, threads, d_cyc, d_sink, sms);
    run<9>("IMAD + LOP3 1:1", 2, threads, d_cyc, d_sink, sms);
    run<10>("IMNMX (max.s32)", 1, threads, d_cyc, d_sink, sms);
    run<11>("PRMT", 1, threads, d_cyc, d_sink, sms);
    run<12>("SHF (shr.s32)", 1, threads, d_cyc, d_sink, sms);
    run<13>("IMAD.HI + LOP3 1:1", 2, threads, d_cyc, d_sink, sms);
    run<14>("IMAD + IMAD.HI 1:1", 2, threads, d_cyc, d_sink, sms);
    run<15>("IMAD + IDP.4A 1:1", 2, threads, d_cyc, d_sink, sms);
  }
  for (int threads : {128, 256, 512, 1024}) {
    k_mma<<<sms, threads>>>(d_cyc, d_sink, 3);
    cudaDeviceSynchronize();
    k_mma<<<sms, threads>>>(d_cyc, d_sink, 3);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    const double mma = (double)(threads / 32) * ITER * 8;
    printf("mma.sync m16n8k32 s8         threads/SM %4d  cycles %10.0f  mma/clk/SM %6.3f  int8 MAC/clk/SM %8.1f\n", threads, avg, mma / avg, mma * 16 * 8 * 32 / avg);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
