// All-gather of the packed output records over peer memory: the one exchange step of this path (one record of
// 444-992 bytes per frame from every rank to every rank; the reference's counterpart is nn.DataParallel's gather,
// scripts/test.py:159). One process per GPU; every rank owns a WINDOW in its own HBM that its peers map through CUDA IPC:
//
//   window = data[world][bytes_per_rank] | arrived[world] | consumed[world]            (32-bit step counters)
//
// all_gather(step s) is two kernels on the caller's stream, no host synchronisation and no NCCL call:
//   scatter  : CTA group q waits until peer q has consumed step s-1 out of its window (consumed[q] >= s-1 in MY window, written
//              by q), then stores this rank's record into slot `rank` of q's window with 16-byte stores over NVLink (or into
//              its own window for q == rank), fences system-wide and publishes arrived[rank] = s in q's window
//   collect  : CTA r waits for arrived[r] >= s in the local window, copies slot r into the caller's output (so the window can
//              be rewritten), and the last CTA publishes consumed[rank] = s in every peer's window
// The stores ARE the transfer: nothing is staged, and the record leaves for its destination as soon as the kernel that
// follows the forward runs. Spins are bounded (a peer that never arrives sets an error word instead of hanging the GPU).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/hrp_b200.h"
#include "common.h"

struct hrp_p2p {
  int rank = 0, world = 1, device = 0;
  size_t bytes = 0;                 // per rank, multiple of 16
  uint8_t* window = nullptr;        // local window (cudaMalloc)
  std::vector<uint8_t*> peer;       // peer[q] = base of q's window as mapped here (peer[rank] = window)
  uint8_t** d_peer = nullptr;       // the same table on the device
  unsigned int* d_state = nullptr;  // [0] error word, [1..world] per-peer CTA counters of scatter, [world+1] counter of collect
  unsigned int step = 0;
  bool connected = false;
};

namespace hrp {
namespace {

constexpr int PG_THREADS = 256;
constexpr int PG_SPLIT = 4;                    // CTAs per destination in the scatter
constexpr unsigned int PG_SPIN_LIMIT = 1u << 24;   // x ~200 ns back-off: a few seconds

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 of the CTA waits until *flag >= want (step counters only grow); false on time-out
__device__ __forceinline__ bool wait_at_least(const unsigned int* flag, unsigned int want) {
  for (unsigned int i = 0; i < PG_SPIN_LIMIT; ++i) {
    if ((int)(ld_acquire_sys(flag) - want) >= 0) return true;
    __nanosleep(200);
  }
  return false;
}

__global__ void __launch_bounds__(PG_THREADS)
p2p_scatter_kernel(const uint4* __restrict__ src, size_t n16, uint8_t* const* __restrict__ peer, size_t bytes, int rank, int world,
                   unsigned int step, unsigned int* __restrict__ state) {
  const int q = blockIdx.x / PG_SPLIT, part = blockIdx.x % PG_SPLIT;
  uint8_t* win_q = peer[q];
  unsigned int* my_consumed = reinterpret_cast<unsigned int*>(peer[rank] + (size_t)world * bytes) + world;   // written by the peers
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = (step <= 1 || wait_at_least(my_consumed + q, step - 1)) ? 1 : 0;   // q has copied step-1 out of its window
    if (!ok) atomicExch(state, 1u);
  }
  __syncthreads();
  if (ok) {
    uint4* dst = reinterpret_cast<uint4*>(win_q + (size_t)rank * bytes);
    for (size_t i = (size_t)part * PG_THREADS + threadIdx.x; i < n16; i += (size_t)PG_SPLIT * PG_THREADS) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // the last of the PG_SPLIT CTAs of destination q publishes the arrival
    if (atomicAdd(state + 1 + q, 1u) == PG_SPLIT - 1) {
      atomicExch(state + 1 + q, 0u);
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned int*>(win_q + (size_t)world * bytes) + rank, step);
    }
  }
}

__global__ void __launch_bounds__(PG_THREADS)
p2p_collect_kernel(uint4* __restrict__ dst, size_t n16, uint8_t* const* __restrict__ peer, size_t bytes, int rank, int world,
                   unsigned int step, unsigned int* __restrict__ state) {
  const int r = blockIdx.x;
  const uint8_t* win = peer[rank];
  const unsigned int* arrived = reinterpret_cast<const unsigned int*>(win + (size_t)world * bytes);
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = wait_at_least(arrived + r, step) ? 1 : 0;
    if (!ok) atomicExch(state, 2u);
  }
  __syncthreads();
  if (ok && dst != nullptr) {
    const uint4* s = reinterpret_cast<const uint4*>(win + (size_t)r * bytes);
    uint4* d = dst + (size_t)r * n16;
    for (size_t i = threadIdx.x; i < n16; i += PG_THREADS) d[i] = s[i];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(state + 1 + world, 1u) == (unsigned int)world - 1) {
    atomicExch(state + 1 + world, 0u);
    for (int q = 0; q < world; ++q)          // every slot of my window has been copied out: the peers may write step + 1
      st_release_sys(reinterpret_cast<unsigned int*>(peer[q] + (size_t)world * bytes) + world + rank, step);
  }
}

}  // namespace
}  // namespace hrp

using namespace hrp;

extern "C" int hrp_p2p_create(int rank, int world, size_t bytes_per_rank, int device, hrp_p2p** out) {
  if (!out || world < 1 || world > 64 || rank < 0 || rank >= world || bytes_per_rank == 0)
    return fail(HRP_ERR_INVALID, "hrp_p2p_create: bad argument (rank %d of %d, %zu bytes)", rank, world, bytes_per_rank);
  hrp_p2p* g = new (std::nothrow) hrp_p2p();
  if (!g) return fail(HRP_ERR_NOMEM, "hrp_p2p_create: out of memory");
  g->rank = rank; g->world = world; g->device = device;
  g->bytes = (bytes_per_rank + 15) / 16 * 16;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaError_t e = cudaSetDevice(device);
  const size_t total = (size_t)world * g->bytes + (size_t)2 * world * sizeof(unsigned int);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->window), total);
  if (e == cudaSuccess) e = cudaMemset(g->window, 0, total);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->d_peer), (size_t)world * sizeof(uint8_t*));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->d_state), (size_t)(world + 2) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(g->d_state, 0, (size_t)(world + 2) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    const int rs = fail(HRP_ERR_CUDA, "hrp_p2p_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    hrp_p2p_destroy(g);
    return rs;
  }
  g->peer.assign(world, nullptr);
  g->peer[rank] = g->window;
  *out = g;
  return HRP_OK;
}

extern "C" int hrp_p2p_handle(hrp_p2p* g, void* handle64) {
  if (!g || !handle64) return fail(HRP_ERR_INVALID, "hrp_p2p_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  HRP_CUDA(cudaIpcGetMemHandle(&h, g->window));
  std::memcpy(handle64, &h, 64);
  return HRP_OK;
}

extern "C" int hrp_p2p_connect(hrp_p2p* g, const void* handles) {
  if (!g || !handles) return fail(HRP_ERR_INVALID, "hrp_p2p_connect: null argument");
  if (g->connected) return fail(HRP_ERR_STATE, "hrp_p2p_connect: already connected");
  int prev = 0;
  cudaGetDevice(&prev);
  HRP_CUDA(cudaSetDevice(g->device));
  for (int q = 0; q < g->world; ++q) {
    if (q == g->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)q * 64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaSetDevice(prev);
      return fail(HRP_ERR_CUDA, "hrp_p2p_connect: cannot map the window of rank %d: %s", q, cudaGetErrorString(e));
    }
    g->peer[q] = static_cast<uint8_t*>(p);
  }
  const cudaError_t e = cudaMemcpy(g->d_peer, g->peer.data(), (size_t)g->world * sizeof(uint8_t*), cudaMemcpyHostToDevice);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(HRP_ERR_CUDA, "hrp_p2p_connect: %s", cudaGetErrorString(e));
  g->connected = true;
  return HRP_OK;
}

extern "C" int hrp_p2p_all_gather(hrp_p2p* g, const void* src, size_t bytes, void* dst, void* stream) {
  if (!g || !src || !dst) return fail(HRP_ERR_INVALID, "hrp_p2p_all_gather: null argument");
  if (!g->connected && g->world > 1) return fail(HRP_ERR_STATE, "hrp_p2p_all_gather: hrp_p2p_connect has not been called");
  if (bytes != g->bytes) return fail(HRP_ERR_INVALID, "hrp_p2p_all_gather: %zu bytes per rank, the window was created for %zu (multiples of 16)", bytes, g->bytes);
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) return fail(HRP_ERR_INVALID, "hrp_p2p_all_gather: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (g->world == 1 && !g->connected) {
    HRP_CUDA(cudaMemcpy(g->d_peer, g->peer.data(), sizeof(uint8_t*), cudaMemcpyHostToDevice));
    g->connected = true;
  }
  const unsigned int step = ++g->step;
  const size_t n16 = bytes / 16;
  p2p_scatter_kernel<<<g->world * PG_SPLIT, PG_THREADS, 0, st>>>(static_cast<const uint4*>(src), n16, g->d_peer, g->bytes, g->rank, g->world, step, g->d_state);
  HRP_CHECK_LAUNCH("p2p_scatter_kernel");
  p2p_collect_kernel<<<g->world, PG_THREADS, 0, st>>>(static_cast<uint4*>(dst), n16, g->d_peer, g->bytes, g->rank, g->world, step, g->d_state);
  HRP_CHECK_LAUNCH("p2p_collect_kernel");
  return HRP_OK;
}

extern "C" int hrp_p2p_status(hrp_p2p* g) {
  if (!g) return fail(HRP_ERR_INVALID, "hrp_p2p_status: null argument");
  unsigned int err = 0;
  HRP_CUDA(cudaMemcpy(&err, g->d_state, sizeof(err), cudaMemcpyDeviceToHost));     // synchronises with the gathers in flight
  if (err) return fail(HRP_ERR_STATE, "hrp_p2p_all_gather: a peer did not %s in time (rank %d of %d)", err == 1 ? "consume the previous step" : "deliver its record",
                       g->rank, g->world);
  return HRP_OK;
}

extern "C" void hrp_p2p_destroy(hrp_p2p* g) {
  if (!g) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < (int)g->peer.size(); ++q)
    if (q != g->rank && g->peer[q]) cudaIpcCloseMemHandle(g->peer[q]);
  if (g->window) cudaFree(g->window);
  if (g->d_peer) cudaFree(g->d_peer);
  if (g->d_state) cudaFree(g->d_state);
  cudaGetLastError();
  cudaSetDevice(prev);
  delete g;
}
