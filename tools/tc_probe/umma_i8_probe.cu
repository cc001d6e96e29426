// umma_i8_probe.cu — pins down the tcgen05.mma kind::i8 conventions the tensor-core FIR relies on, on real hardware:
//   * shared-memory matrix descriptor for K-major, no-swizzle operands (which of LBO / SBO is the K-direction stride),
//   * independent signedness of A and B (a_format / b_format in the instruction descriptor),
//   * accumulate predicate, TMEM addressing and the tcgen05.ld 32x32b shape.
// One CTA, D[128 x 64] (+)= A[128 x K] * B[64 x K]^T with K = 64 (two MMAs of K = 32), int8 operands, int32 accumulators.
// Operand layout in smem: core matrices of 8 rows x 16 bytes (128 contiguous bytes), ordered [k16][row8][8][16].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_probe umma_i8_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);              // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading dimension byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride dimension byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  return d;                                             // base offset 0, swizzle none (layout_type 0)
}

__global__ void __launch_bounds__(128) probe(const int8_t *gA, const int8_t *gB, int32_t *gD, int hyp, int a_signed, int b_signed)
{
  __shared__ __align__(1024) uint8_t sA[M * K];
  __shared__ __align__(1024) uint8_t sB[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  // canonical layout: element (row r, byte k) -> [(k/16)][(r/8)][r%8][k%16]
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    sA[((k / 16) * (M / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 16)] = (uint8_t)gA[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    sB[((k / 16) * (N / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 16)] = (uint8_t)gB[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    const uint32_t kstrideA = (M / 8) * 128, kstrideB = (N / 8) * 128, rstride = 128;
    const uint32_t idesc = (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 32; ++ks) {
      const uint32_t aaddr = smem_u32(sA) + ks * 2 * kstrideA, baddr = smem_u32(sB) + ks * 2 * kstrideB;
      const uint64_t da = hyp == 0 ? make_desc(aaddr, kstrideA, rstride) : make_desc(aaddr, rstride, kstrideA);
      const uint64_t db = hyp == 0 ? make_desc(baddr, kstrideB, rstride) : make_desc(baddr, rstride, kstrideB);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // thread t = TMEM lane t = row t of D; 64 columns in chunks of 8
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) gD[tid * N + c0 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main()
{
  int8_t *hA = (int8_t *)malloc(M * K), *hB = (int8_t *)malloc(N * K);
  srand(7);
  for (int i = 0; i < M * K; ++i) hA[i] = (int8_t)(rand() & 0xFF);
  for (int i = 0; i < N * K; ++i) hB[i] = (int8_t)(rand() & 0xFF);
  int8_t *dA, *dB; int32_t *dD;
  cudaMalloc(&dA, M * K); cudaMalloc(&dB, N * K); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K, cudaMemcpyHostToDevice);
  int32_t *hD = (int32_t *)malloc(M * N * 4);
  for (int hyp = 0; hyp < 2; ++hyp)
    for (int as = 0; as < 2; ++as)
      for (int bs = 0; bs < 2; ++bs) {
        cudaMemset(dD, 0xEE, M * N * 4);
        probe<<<1, 128>>>(dA, dB, dD, hyp, as, bs);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("hyp %d a_signed %d b_signed %d : CUDA error %s\n", hyp, as, bs, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n) {
            int32_t ref = 0;
            for (int k = 0; k < K; ++k) {
              const int a = as ? (int)hA[m * K + k] : (int)(uint8_t)hA[m * K + k];
              const int b = bs ? (int)hB[n * K + k] : (int)(uint8_t)hB[n * K + k];
              ref += a * b;
            }
            if (ref != hD[m * N + n]) ++bad;
          }
        printf("hypothesis %d (LBO = %s stride)  A %s  B %s : %ld / %d mismatches  D[0][0..3] = %d %d %d %d\n", hyp, hyp == 0 ? "K-direction" : "row-group",
               as ? "s8" : "u8", bs ? "s8" : "u8", bad, M * N, hD[0], hD[1], hD[2], hD[3]);
      }
  return 0;
}
