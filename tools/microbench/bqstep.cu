// bqstep.cu — how many cycles does ONE warp need per biquad sample-step?  The biquad is a serial recurrence per channel,
// so for few-channel workloads (BASELINE C3: 4096 channels = 128 warps on 148 SMs) chip throughput is
// channels * clock / cycles_per_step, whatever the FIR does.  Variants: integer (IMAD.HI) / FP64 (DFMA.RM) products,
// 1 or 2 fused stages, with/without the feed-forward terms.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../minimal-sdr_b200/csrc -o bqstep bqstep.cu
#include <cstdio>
#include "msdr_device.cuh"
using namespace msdr;

#define STEPS 8192

template <int MODE>
__global__ void k(long long *cyc, int *sink, int seed)
{
  BqStage si[2];
  BqStageD sd[2];
  for (int k2 = 0; k2 < 2; ++k2) {
    si[k2].b0 = 236552419 + seed; si[k2].b1 = 473104839; si[k2].b2 = 236552419; si[k2].a1 = 175469220; si[k2].a2 = -47937074;
    si[k2].x1 = si[k2].x2 = si[k2].y1 = si[k2].y2 = 0; si[k2].res = 0;
    bq_set_coefs(sd[k2], si[k2].b0, si[k2].b1, si[k2].b2, si[k2].a1, si[k2].a2);
    sd[k2].x1 = sd[k2].x2 = sd[k2].y1 = sd[k2].y2 = bq_d_from_int(0); sd[k2].res = 0;
  }
  int x = (threadIdx.x * 977 + seed) & 0x7fff, acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 8
  for (int n = 0; n < STEPS; ++n) {
    x = (x * 75 + 74) & 0x7fff; // cheap input generator (2 ALU/FMA ops)
    if (MODE == 0) { int v = bq_step(si[0], x << 16); v = bq_step(si[1], v); acc ^= v; }            // int, 2 stages
    if (MODE == 1) { int v = bq_step(si[0], x << 16); acc ^= v; }                                    // int, 1 stage
    if (MODE == 2) { int y; double v = bq_step(sd[0], bq_d_from_int(x), y); bq_step(sd[1], v, y); acc ^= y; } // f64, 2 stages
    if (MODE == 3) { int y; bq_step(sd[0], bq_d_from_int(x), y); acc ^= y; }                         // f64, 1 stage
    if (MODE == 4) { // f64, recurrence only (feed-forward sum supplied): 2 DFMA
      BqStageD &s = sd[0];
      const int early = x + s.res + bq_term_d(s.a2, s.y2);
      const int sum = early + bq_term_d(s.a1, s.y1);
      const int y = ssat16(sum >> 14); s.res = sum & 0x3FFF;
      s.y2 = s.y1; s.y1 = bq_d_from_int(y); acc ^= y;
    }
    if (MODE == 5) { // int, recurrence only: 2 IMAD.HI
      BqStage &s = si[0];
      const int early = x + s.res + __mulhi(s.a2, s.y2);
      const int sum = smlaw_s(early, s.a1, s.y1);
      const int y = ssat16(sum >> 14); s.res = sum & 0x3FFF;
      s.y2 = s.y1; s.y1 = y << 16; acc ^= y;
    }
    if (MODE == 6) { // f64 recurrence only, two independent stages interleaved (ILP 2)
      for (int q = 0; q < 2; ++q) {
        BqStageD &s = sd[q];
        const int early = x + s.res + bq_term_d(s.a2, s.y2);
        const int sum = early + bq_term_d(s.a1, s.y1);
        const int y = ssat16(sum >> 14); s.res = sum & 0x3FFF;
        s.y2 = s.y1; s.y1 = bq_d_from_int(y); acc ^= y;
      }
    }
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads, long long *d_cyc, int *d_sink, int sms)
{
  k<MODE><<<sms, threads>>>(d_cyc, d_sink, 1);
  cudaDeviceSynchronize();
  k<MODE><<<sms, threads>>>(d_cyc, d_sink, 2);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
  printf("%-44s warps/SM %2d  cycles/step %7.1f  -> steps/clk/SM %6.3f lanes\n", name, threads / 32, avg / STEPS, 1.0 * threads * STEPS / avg);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long *d_cyc; int *d_sink;
  cudaMalloc(&d_cyc, 1024 * sizeof(long long)); cudaMalloc(&d_sink, (size_t)sms * 1024 * sizeof(int));
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<0>("int (IMAD.HI) 2 stages fused", threads, d_cyc, d_sink, sms);
    run<1>("int (IMAD.HI) 1 stage", threads, d_cyc, d_sink, sms);
    run<2>("f64 (DFMA.RM) 2 stages fused", threads, d_cyc, d_sink, sms);
    run<3>("f64 (DFMA.RM) 1 stage", threads, d_cyc, d_sink, sms);
    run<4>("f64 recurrence only (2 DFMA)", threads, d_cyc, d_sink, sms);
    run<5>("int recurrence only (2 IMAD.HI)", threads, d_cyc, d_sink, sms);
    run<6>("f64 recurrence only x2 interleaved", threads, d_cyc, d_sink, sms);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
