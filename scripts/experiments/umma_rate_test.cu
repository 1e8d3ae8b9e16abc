// Experiment: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, bf16) as a function of N, the swizzle
// mode of the K-major operands, a row shift of the A descriptor, and the number of independent TMEM accumulators the
// MMAs rotate over. One CTA per SM, one issuing thread, REPS MMAs, clock64 around issue + commit + wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate_test scripts/experiments/umma_rate_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) rate_kernel(int N, int rowb, int shift_rows, int nacc, int reps, int a_span_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 96 * 1024, sBar = sB + 32 * 1024, sSlot = sBar + 8;
  const int tid = threadIdx.x;
  for (uint32_t i = tid; i < (96 + 32) * 1024 / 16; i += 128) *reinterpret_cast<uint4*>(raw + (base - smem_u32(raw)) + i * 16) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sSlot));
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t sbo = (uint64_t)(8 * rowb) >> 4, layout = rowb == 128 ? 2ull : 4ull;
    const uint64_t hi = (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
    const uint64_t db = (uint64_t)((sB & 0x3FFFFu) >> 4) | hi;
    const int ksteps = rowb / 32;
    // descriptors are loop-invariant: 8 unrolled MMAs per iteration, accumulators rotate (a = i % nacc), A windows rotate
    // over four 128-row blocks, K step alternates, so the issue path is as short as it can be
    uint64_t das[8];
    uint32_t ds[8];
    for (int i = 0; i < 8; ++i) {
      const uint32_t a_start = sA + (uint32_t)(shift_rows + (i & 3) * 128) * rowb;
      das[i] = ((uint64_t)((a_start & 0x3FFFFu) >> 4) | hi) + (uint64_t)((i & 1) * 2 % (ksteps * 2));
      ds[i] = tmem + (uint32_t)((i % nacc) * N);
    }
    t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(ds[i]),
                     "l"(das[i]), "l"(db), "r"(idesc), "r"(1u)
                     : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sBar) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  const int reps = 4096;
  printf("clk per MMA (M=128, K=16 bf16), %d MMAs, 148 CTAs\n", reps);
  printf("%5s %5s %6s %5s %8s\n", "N", "rowB", "shift", "nacc", "clk/MMA");
  const int Ns[] = {32, 64, 128, 256};
  for (int rowb : {128, 64})
    for (int N : Ns)
      for (int shift : {0, 3})
        for (int nacc : {1, 2, 4}) {
          if (nacc * N > 512) continue;
          rate_kernel<<<148, 128, 140 * 1024>>>(N, rowb, shift, nacc, reps, 512, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("%5d %5d %6d %5d %8.1f\n", N, rowb, shift, nacc, (double)c / reps);
        }
  return 0;
}
