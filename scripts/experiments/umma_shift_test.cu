// Experiment: does tcgen05.mma read a K-major swizzled operand correctly when the descriptor start address is offset by
// whole rows (not a multiple of the 8-row swizzle atom)? The slab is written with the absolute-address XOR pattern TMA
// uses. Prints, per row shift, whether D == A[shift..shift+127] * B^T exactly.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_shift_test scripts/experiments/umma_shift_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB>
__global__ void __launch_bounds__(128) shift_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int rows, int shift_rows, int base_off_mode) {
  constexpr int KE = ROWB / 2;        // K elements per row
  constexpr int N = 32;
  extern __shared__ __align__(1024) uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gbase = raw + (base - smem_u32(raw));
  const uint32_t sA = base, sB = base + 64 * 1024, sBar = sB + 8192, sSlot = sBar + 8;
  const int tid = threadIdx.x;
  // fill A slab (rows x ROWB) and B tile (N x ROWB) with the absolute-address swizzle
  for (int idx = tid; idx < rows * (ROWB / 16); idx += 128) {
    const int r = idx / (ROWB / 16), c = idx % (ROWB / 16);
    const uint32_t addr = (uint32_t)r * ROWB;                       // offset of the row inside the 1024-aligned slab
    const uint32_t sw = ROWB == 128 ? ((addr >> 7) & 7) : ((addr >> 7) & 3);
    *reinterpret_cast<uint4*>(gbase + addr + ((c ^ sw) << 4)) = *reinterpret_cast<const uint4*>(A + (size_t)r * KE + c * 8);
  }
  for (int idx = tid; idx < N * (ROWB / 16); idx += 128) {
    const int r = idx / (ROWB / 16), c = idx % (ROWB / 16);
    const uint32_t addr = (uint32_t)r * ROWB;
    const uint32_t sw = ROWB == 128 ? ((addr >> 7) & 7) : ((addr >> 7) & 3);
    *reinterpret_cast<uint4*>(gbase + 64 * 1024 + addr + ((c ^ sw) << 4)) = *reinterpret_cast<const uint4*>(B + (size_t)r * KE + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(sSlot));
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t sbo = (uint64_t)(8 * ROWB) >> 4, layout = ROWB == 128 ? 2ull : 4ull;
    const uint32_t a_start = sA + (uint32_t)shift_rows * ROWB;
    uint64_t boff = 0;
    if (base_off_mode == 1) boff = (uint64_t)((a_start >> 7) & 7);
    const uint64_t da = (uint64_t)((a_start & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (boff << 49) | (layout << 61);
    const uint64_t db = (uint64_t)((sB & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
    for (int kk = 0; kk < ROWB / 32; ++kk) {
      const uint32_t acc = kk != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                   "l"(da + kk * 2), "l"(db + kk * 2), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sBar) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

template <int ROWB>
int run(int mode) {
  constexpr int KE = ROWB / 2, ROWS = 320, N = 32;
  std::vector<__nv_bfloat16> hA(ROWS * KE), hB(N * KE);
  std::vector<float> fA(ROWS * KE), fB(N * KE);
  for (int r = 0; r < ROWS; ++r)
    for (int k = 0; k < KE; ++k) { fA[r * KE + k] = (float)((r * 7 + k * 3) % 17 - 8); hA[r * KE + k] = __float2bfloat16(fA[r * KE + k]); }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < KE; ++k) { fB[n * KE + k] = (float)((n * 5 + k) % 13 - 6); hB[n * KE + k] = __float2bfloat16(fB[n * KE + k]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(shift_kernel<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int bad_total = 0;
  for (int shift = 0; shift <= 20; ++shift) {
    shift_kernel<ROWB><<<1, 128, 100 * 1024>>>(dA, dB, dD, ROWS, shift, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ROWB=%d mode=%d shift=%d: CUDA error %s\n", ROWB, mode, shift, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(128 * 32);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < KE; ++k) ref += fA[(i + shift) * KE + k] * fB[n * KE + k];
        if (ref != hD[i * 32 + n]) ++bad;
      }
    printf("ROWB=%d base_offset_mode=%d shift=%2d rows: %s (%d mismatches)\n", ROWB, mode, shift, bad ? "WRONG" : "exact", bad);
    bad_total += bad;
  }
  return bad_total != 0;
}

int main() {
  int rc = 0;
  rc |= run<128>(0);
  rc |= run<128>(1);
  rc |= run<64>(0);
  rc |= run<64>(1);
  return 0;
}
