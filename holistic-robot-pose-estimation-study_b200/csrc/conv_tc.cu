// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::f16 with bf16 operands, or kind::tf32) with the
// accumulator in TMEM, weights streamed by the TMA bulk-copy engine, activations fetched by TMA tensor copies (im2col
// on the fly: one {channels, W-run, rows, images} box per kernel tap, zero-filled outside the image) into the
// 128-/64-byte-swizzled K-major operand layout the MMA reads, results returned by TMA tensor stores.
//
// Replaces every nn.Conv2d / nn.ConvTranspose2d (+ folded BatchNorm2d, + residual add, + ReLU) the reference runs in
// the two backbones and the deconv head (lib/models/backbones/HRnet.py:28-98,247-265,499-570; Resnet.py:57-139;
// lib/models/full_net.py:214-238,353-355) in the HRP_PREC_BF16 / HRP_PREC_TF32 families, except the 3x3/s1 Cin==Cout
// layers that conv_slab.cu / conv_block.cu take.
//
// GEMM view (same ConvArgs descriptor as the fp32 family): M = B*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin with
// k = (r*KW + s)*Cin + c. Persistent CTAs (one or two per SM) walk 128 x BLOCK_N tiles:
//   warps 0-3  only for shapes whose tile is not a box of the input: cp.async im2col gather (thread t serves chunk j of
//              eight rows of its warp's 32; zero-filled outside the image), XOR-swizzled by hand
//   warp 4     allocates TMEM; issues the tcgen05.mma stream into one of two accumulator buffers and commits ring
//              slots back (all lanes walk the loop, one elected lane issues: tc_ptx.h, elect_one)
//   warp 5     loader: pre-swizzled weight tiles with cp.async.bulk and the activation boxes with
//              cp.async.bulk.tensor.4d, both on the stage's mbarrier (complete_tx)
//   warps 6..  4 or 8 epilogue warps drain the other accumulator buffer: + bias, + residual, ReLU, convert. Dense NHWC
//              outputs go through a shared-memory tile already in the TMA swizzle and leave (and the residual arrives)
//              as tensor copies; strided outputs (transposed-conv phases) and the NCHW fp32 heatmap logits keep a
//              register / cp.async path.
// Stages hand over through mbarriers: full[s] (expect_tx, + 128 producer arrivals in gather mode), empty[s]
// (tcgen05.commit), acc_full / acc_empty per accumulator buffer.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "sa_state.h"
#include "tc_ptx.h"

namespace hrp {
namespace {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_ROW_BYTES = 128;                          // one swizzle-128B row of K
constexpr int TC_A_STAGE = TC_BLOCK_M * TC_ROW_BYTES;      // 16 KB
constexpr int TC_PRODUCERS = 128;
// threads: 4 producer warps + MMA warp + weight-loader warp + EPI epilogue warps (4, or 8 = two per TMEM lane quarter)
constexpr int tc_threads(int epi_warps) { return 192 + 32 * epi_warps; }
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_LAG = 6;
constexpr int TC_SMEM_LIMIT = 227 * 1024;

struct TcParams {
  alignas(64) unsigned char tmap[128];   // CUtensorMap of the NHWC input (TMA mode)
  alignas(64) unsigned char tmap_out[128];   // epi_tma: output as [M][ld_out], box {chunk columns, 128 rows}
  alignas(64) unsigned char tmap_res[128];   // epi_tma: residual, same geometry
  int pdl_early;     // HRP_PDL_EARLY: let the next kernel start launching right after this one's prologue
  int epi_tma;       // 1: the epilogue stages the tile in swizzled shared memory and moves it with TMA (residual in, result out)
  int chunk_bytes;   // epi_tma: bytes of one staged row chunk (128, or 64 when block_n*elem == 64)
  int n_chunks;      // epi_tma: block_n*elem / chunk_bytes
  int a_res;         // activation-resident walk (1x1, TMA mode, one CTA = whole M tiles): the num_kb activation k-blocks of an M tile are loaded
                     // ONCE into the A halves of stages 0..num_kb-1 and every N tile of that M tile multiplies them; the ring carries weights only
  int epi_dual;      // epi_tma, TMA operand loads, EPI == 4, two staging buffers: warps 0-3 are a second epilogue group (odd tiles)
  int epi_wide;      // epi_tma, 2-byte families: drain 32 columns per tcgen05.ld with the residual / bias loads in flight before the wait
  int res_prefetch;  // epi_tma, two staging buffers: request the residual of tile li+1 at the end of tile li
  int n_stg;         // epi_tma: staging buffers (2 when shared memory allows: the store of tile i overlaps tile i+1)
  ConvArgs a;
  int tma;           // 1: activations arrive by cp.async.bulk.tensor (one thread), 0: cp.async gather (four warps)
  int row_bytes;     // bytes of K per operand row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B, Cin*elem == 64 in TMA mode)
  int M;             // B*Ho*Wo
  int Ktot;          // KH*KW*Cin (elements)
  int num_kb;        // ceil(Ktot / elements per 128-byte row)
  int block_n;       // multiple of 16, <= 256, divides Cout
  int stages;
  int lag;           // cp.async groups kept in flight per producer thread (< stages)
  int tmem_cols;     // power of two >= max(32, block_n)
  int tiles_n, total_tiles;
  int stg_off;       // byte offset of the epilogue staging tile (after the operand stages)
  int bar_off;       // byte offset of the barrier block
  int bias_off;      // byte offset of the layer's bias vector in shared memory (Cout floats, filled once per CTA)
  int cpr_log;       // log2 of 16-byte chunks per staged tile row (block_n * elem / 16)
  int round_tf32;    // round fp32 NHWC outputs to TF32 (nearest, ties away) so the next MMA sees exact operands
  int cluster;       // CTAs per thread-block cluster (1, 2 or 4): they work on `cluster` consecutive M tiles of the SAME N tile in lockstep, and
                     // every weight tile is fetched from L2 once per cluster (each CTA multicasts its 1/cluster of the rows to all)
  int total_units;   // work units a cluster walks: ceil(M tiles / cluster) * tiles_n  (cluster == 1: the tiles themselves)
  int x3;            // TF32 only: 3xTF32 -- operands split into hi + lo TF32 halves, D += Ahi Whi + Alo Whi + Ahi Wlo (fp32-grade products)
};

using namespace tc;

// ---- the kernel -------------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory"); }   // the epilogue warps

template <int N>
__device__ __forceinline__ void cp_async_wait_dyn(int n) {      // cp.async.wait_group takes an immediate
  if constexpr (N == 0) {
    cp_async_wait<0>();
  } else {
    if (n >= N) cp_async_wait<N>(); else cp_async_wait_dyn<N - 1>(n);
  }
}

// 3xTF32 operand split of one 16-byte chunk (4 fp32) in shared memory: hi = the top 19 bits (exactly what kind::tf32 reads),
// written back in place; lo = x - hi (exact in fp32), itself cut to TF32 (error 2^-22 |x|), written `lo_off` bytes further.
__device__ __forceinline__ void split_chunk(uint32_t addr, uint32_t lo_off) {
  uint32_t v[4], h[4], l[4];
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = v[e] & 0xffffe000u;
    l[e] = __float_as_uint(__uint_as_float(v[e]) - __uint_as_float(h[e])) & 0xffffe000u;
  }
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + lo_off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}
// gather mode: the eight chunks this producer thread copied into a stage (conv_tc_kernel, producers)
__device__ __forceinline__ void split_own_chunks(uint32_t row0, uint32_t sw_even, uint32_t sw_odd, uint32_t lo_off) {
#pragma unroll
  for (int i = 0; i < 8; ++i) split_chunk(row0 + ((i & 1) ? sw_odd : sw_even) + (uint32_t)i * (4 * TC_ROW_BYTES), lo_off);
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Persistent, warp-specialised: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (N tile fastest, so
// co-resident CTAs share the activation rows they gather). The operand ring and both pipelines run across tile
// boundaries: producers are already gathering tile i+1 while the MMA warp finishes tile i and the epilogue warps drain
// the other TMEM accumulator buffer.
template <bool TF32, int EPI, bool F16>
__global__ void __launch_bounds__(tc_threads(EPI), EPI == 4 ? 2 : 1)
conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int KB = TC_ROW_BYTES / ESZ;     // K elements per k-block (64 bf16 / 32 tf32)
  constexpr int CE = 16 / ESZ;               // elements per 16-byte chunk
  constexpr int UK = 32 / ESZ;               // K elements per tcgen05.mma (16 bf16 / 8 tf32)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  // 3xTF32: every stage holds the operand tile twice, [hi | lo]. The activation tile arrives as raw fp32; warps 0-3 split
  // it in place (hi = the 19 bits the MMA reads, lo = the remainder, itself exact in TF32 to 2^-22) before the MMA warp sees it
  const bool x3 = TF32 && p.x3 != 0;
  const int a_half = TC_BLOCK_M * p.row_bytes, b_half = p.block_n * p.row_bytes;
  const int a_stage = x3 ? 2 * a_half : a_half, b_stage = x3 ? 2 * b_half : b_half;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + (uint32_t)(S * a_stage);
  const uint32_t stg = smem_base + (uint32_t)p.stg_off;   // epilogue staging tile: 128 rows x (block_n*ESZ + 16) bytes
  const uint32_t sBar = smem_base + (uint32_t)p.bar_off;  // full[S], empty[S], acc_full[2], acc_empty[2], tmem slot, row offsets
  const uint32_t bar_full = sBar, bar_empty = sBar + 8u * TC_MAX_STAGES, bar_accf = sBar + 16u * TC_MAX_STAGES,
                 bar_acce = bar_accf + 16u, tmem_slot = bar_acce + 16u, s_rowoff = tmem_slot + 16u,
                 bar_split = sBar + 16u * TC_MAX_STAGES + 48u + 16u * TC_BLOCK_M + 16u;   // 3xTF32, TMA mode: stage s has been split

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ConvArgs& a = p.a;
  // work units: unit u = (group of `cluster` consecutive M tiles, N tile), N tile fastest; CTA rank r of the cluster takes
  // M tile group*cluster + r. Without clusters a unit is a tile and a "cluster" is one CTA.
  const int CS = p.cluster;
  const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
  // activation-resident walk: a CTA takes whole M tiles (blockIdx.x, + gridDim.x, ...) and visits their N tiles in turn; in unit
  // numbering (u = m * tiles_n + n) that is u+1 inside an M tile and a jump of (gridDim.x - 1) M tiles after its last N tile
  const bool a_res = p.a_res != 0;
  const int unit0 = a_res ? (int)blockIdx.x * p.tiles_n : (int)blockIdx.x / CS, unit_step = (int)gridDim.x / CS;
  const int res_jump = ((int)gridDim.x - 1) * p.tiles_n;
  auto next_unit = [&](int u) { return a_res ? (u + 1 + (((u + 1) % p.tiles_n == 0) ? res_jump : 0)) : u + unit_step; };
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
#define HRP_TILE_M(u) (((u) / p.tiles_n) * CS + crank)
#define HRP_TILE_N(u) ((u) % p.tiles_n)

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8u * s, p.tma ? 1 : TC_PRODUCERS + 1);
      mbar_init(bar_empty + 8u * s, (uint32_t)p.cluster);     // one tcgen05.commit per CTA that received the stage's weight tile
      mbar_init(bar_split + 8u * s, p.a_res ? 1u : (uint32_t)TC_PRODUCERS);   // a_res: [0] = activations landed, [1] = activations consumed
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accf + 8u * i, 1); mbar_init(bar_acce + 8u * i, 1); }
    for (int i = 0; i < 2; ++i) mbar_init(sBar + 16u * TC_MAX_STAGES + 48u + 16u * TC_BLOCK_M + 8u * i, 1);   // residual boxes landed (epi_tma)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5 && lane == 0 && p.tma) asm volatile("prefetch.tensormap [%0];" ::"l"(p.tmap) : "memory");
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  if (p.cluster > 1) cluster_sync_all(); else __syncthreads();   // peers multicast into this CTA's stages and arrive on its barriers
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (p.pdl_early) pdl_trigger();
  if (warp != 4) pdl_wait();      // every warp that touches global memory (the MMA warp only reads shared memory / TMEM)

  // TMA mode leaves warps 0-3 without a role (no gather, no 3xTF32 split); with the TMA epilogue and two staging buffers they
  // become a SECOND EPILOGUE GROUP (p.epi_dual): group g drains accumulator buffer g into staging buffer g, i.e. the CTA's
  // even / odd tiles, each group with its own named barrier and its own storing thread. The drain is a chain of
  // fixed-latency instructions of in-order warps (ncu: issue slots 28 % busy, every epilogue warp ~80 % of its time inside
  // the drain), so twice the warps per SM is what raises the rate; shared memory, TMEM and registers stay as they are.
  const bool dual = p.epi_dual != 0;
  if (warp < 4 && !dual) {
    if (!p.tma) {
    // ===== producers: im2col gather. One warp instruction covers 4 tile rows x 128 contiguous bytes (lane = 8*row + chunk),
    // so every request touches 4 cache lines instead of 32; a thread serves chunk `j` of 8 rows of its warp's 32.
    // Per tile, lane l first describes tile row warp*32+l as {byte offset of tap (0,0), bitmask of in-image taps} in
    // shared memory; the gather loop then costs one AND, one select and one add per 16-byte copy. ======================
    const int j = lane & 7, rsub = lane >> 3;
    const uint32_t dst0 = (uint32_t)(warp * 32 + rsub) * TC_ROW_BYTES;               // gather mode: row_bytes == 128
    const uint32_t sw_even = ((uint32_t)j ^ (uint32_t)rsub) << 4, sw_odd = ((uint32_t)j ^ (uint32_t)(4 | rsub)) << 4;
    const uint32_t s_info = s_rowoff + 8u * TC_BLOCK_M;       // 128 x {int32 offset, uint32 tap mask}
    const uint8_t* in8 = static_cast<const uint8_t*>(a.in);
    const int lag = p.lag;
    const int hw_o = a.Ho * a.Wo;
    const int Cin = a.Cin, KW = a.KW, Wi = a.Wi, Ktot = p.Ktot, num_kb = p.num_kb;
    int it = 0;
    int s = 0, as = 0;                                         // ring slot being filled / slot whose copies are awaited (lag behind)
    uint32_t ph = 1;                                           // parity to wait on for empty[s] (first pass: slots start free)
    for (int u = unit0; u < p.total_units; u = next_unit(u)) {
      const int m0 = HRP_TILE_M(u) * TC_BLOCK_M;
      const int b0 = m0 / hw_o;                                // first image this tile touches
      const uint8_t* tile_base = in8 + (size_t)b0 * a.Hi * a.Wi * a.Cin * ESZ;
      {
        const int m = m0 + warp * 32 + lane;
        const bool ok = m < p.M;
        const int mm = ok ? m : m0;
        const int ox = mm % a.Wo, t1 = mm / a.Wo, oy = t1 % a.Ho, b = t1 / a.Ho;
        const int iy0 = oy * a.stride - a.pad_h, ix0 = ox * a.stride - a.pad_w;
        uint32_t xmask = 0, mask = 0;
        for (int fs = 0; fs < a.KW; ++fs) xmask |= ((ix0 + fs >= 0 && ix0 + fs < a.Wi) ? 1u : 0u) << fs;
        for (int fr = 0; fr < a.KH; ++fr)
          if (iy0 + fr >= 0 && iy0 + fr < a.Hi) mask |= xmask << (fr * a.KW);
        if (!ok) mask = 0;
        const int off0 = (((b - b0) * a.Hi + iy0) * a.Wi + ix0) * a.Cin * ESZ;
        __syncwarp();                                          // previous tile's descriptors have been consumed
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_info + 8u * (uint32_t)(warp * 32 + lane)), "r"(off0), "r"(mask) : "memory");
        __syncwarp();
      }
      int off0[8];
      uint32_t tmask[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(off0[i]), "=r"(tmask[i]) : "r"(s_info + 8u * (uint32_t)(warp * 32 + i * 4 + rsub)));
      int c = j * CE, fr = 0, fs = 0, kthr = j * CE;
      while (c >= Cin) { c -= Cin; if (++fs == KW) { fs = 0; ++fr; } }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        if (it >= S) mbar_wait_relaxed(bar_empty + 8u * s, ph);
        if (kthr < Ktot) {
          const uint32_t dst_e = sA + (uint32_t)(s * a_stage) + dst0 + sw_even, dst_o = sA + (uint32_t)(s * a_stage) + dst0 + sw_odd;
          const uint32_t tapbit = 1u << (fr * KW + fs);
          const int dtap = ((fr * Wi + fs) * Cin + c) * ESZ;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = (tmask[i] & tapbit) != 0;
            const int off = ok ? off0[i] + dtap : 0;
            cp_async16(((i & 1) ? dst_o : dst_e) + (uint32_t)i * (4 * TC_ROW_BYTES), tile_base + off, ok ? 16u : 0u);
          }
        }
        kthr += KB;
        c += KB;
        while (c >= Cin) { c -= Cin; if (++fs == KW) { fs = 0; ++fr; } }
        cp_async_commit();
        if (++s == S) { s = 0; ph ^= 1u; }
        if (it >= lag) {
          cp_async_wait_dyn<TC_MAX_LAG>(lag);
          if (x3) split_own_chunks(sA + (uint32_t)(as * a_stage) + dst0, sw_even, sw_odd, (uint32_t)a_half);
          fence_proxy_async();
          mbar_arrive(bar_full + 8u * as);
          if (++as == S) as = 0;
        }
      }
    }
    cp_async_wait<0>();
    for (int k = max(it - lag, 0); k < it; ++k) {
      if (x3) split_own_chunks(sA + (uint32_t)(as * a_stage) + dst0, sw_even, sw_odd, (uint32_t)a_half);
      fence_proxy_async();
      mbar_arrive(bar_full + 8u * as);
      if (++as == S) as = 0;
    }
    } else if (x3) {
      // ===== 3xTF32 splitters (TMA mode): the stage's activation tile has landed -> hi in place, lo beside it ==========
      const uint32_t n16 = (uint32_t)a_half >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int u = unit0; u < p.total_units; u = next_unit(u))
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(bar_full + 8u * s, ph);
          const uint32_t base = sA + (uint32_t)(s * a_stage);
          for (uint32_t i = (uint32_t)tid; i < n16; i += TC_PRODUCERS) split_chunk(base + (i << 4), (uint32_t)a_half);
          fence_proxy_async();                                  // generic-proxy writes -> visible to tcgen05.mma
          mbar_arrive(bar_split + 8u * s);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
    }
  } else if (warp == 4) {
    // ===== MMA issuer: all lanes walk the loops (uniform operands), one elected lane issues (tc_ptx.h: elect_one) ========
    const bool leader = elect_one();
    constexpr uint32_t FMT = TF32 ? 2u : (F16 ? 0u : 1u);       // operand format: TF32 / IEEE half / bf16
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                           ((uint32_t)(TC_BLOCK_M >> 4) << 24);
    const uint32_t dhi = umma_desc_hi(p.row_bytes);
    const int kbe = p.row_bytes / ESZ;                          // K elements per k-block
    int li = 0, s = 0;
    uint32_t ph = 0;
    const int num_kb = p.num_kb;
    int mj = 0;                                                  // a_res: M tiles this CTA has started
    for (int u = unit0; u < p.total_units; u = next_unit(u), ++li) {
      const int buf = li & 1;
      if (li >= 2) mbar_wait(bar_acce + 8u * buf, ((li >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.block_n);
      const bool first_n = a_res && (u % p.tiles_n) == 0, last_n = a_res && ((u + 1) % p.tiles_n) == 0;
      if (first_n) { mbar_wait(bar_split, (uint32_t)(mj & 1)); ++mj; }     // this M tile's activation k-blocks have landed
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(((x3 && p.tma) ? bar_split : bar_full) + 8u * s, ph);
        tc_fence_after();
        const int kleft = p.Ktot - kb * kbe;
        const int nk = (kleft >= kbe ? kbe : kleft) / UK;
        const uint32_t aa = sA + (uint32_t)((a_res ? kb : s) * a_stage), bb = sB + (uint32_t)(s * b_stage);
        if (x3) {
          for (int kk = 0; kk < nk; ++kk)
            if (leader) {                                        // small terms first, the leading product last
              umma<TF32>(tmem_d, umma_desc_at(aa + (uint32_t)a_half + 32u * kk, dhi), umma_desc_at(bb + 32u * kk, dhi), idesc, (kb | kk) != 0);
              umma<TF32>(tmem_d, umma_desc_at(aa + 32u * kk, dhi), umma_desc_at(bb + (uint32_t)b_half + 32u * kk, dhi), idesc, 1u);
              umma<TF32>(tmem_d, umma_desc_at(aa + 32u * kk, dhi), umma_desc_at(bb + 32u * kk, dhi), idesc, 1u);
            }
        } else
        for (int kk = 0; kk < nk; ++kk)
          if (leader) umma<TF32>(tmem_d, umma_desc_at(aa + 32u * kk, dhi), umma_desc_at(bb + 32u * kk, dhi), idesc, (kb | kk) != 0);
        if (leader) {
          if (CS > 1) umma_commit_mc(bar_empty + 8u * s, cmask);    // ... in every CTA whose loader writes into this CTA's stage
          else umma_commit(bar_empty + 8u * s);                  // frees the stage once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(bar_accf + 8u * buf);   // accumulator complete
          if (last_n && kb == num_kb - 1) umma_commit(bar_split + 8u);   // a_res: the resident activations may be replaced
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===== loader: weights [kb][Cout][row_bytes] pre-swizzled, one bulk copy per k-block; in TMA mode also the
    // activation tile: one 4-D tensor copy {channels, Wo-run, rows, images} per k-block, zero-filled outside the image.
    // All lanes walk the loops, one elected lane issues (uniform operands; tc_ptx.h: elect_one) ============================
    {
      const bool leader = elect_one();
      const size_t kb_stride = (size_t)a.Cout * p.row_bytes * (x3 ? 2 : 1);    // 3xTF32 image: [kb][hi | lo][Cout][row]
      const int kbe = p.row_bytes / ESZ;
      const uint32_t tx = (uint32_t)b_stage + (p.tma ? (uint32_t)a_half : 0u);
      int it = 0, s = 0;
      uint32_t ph = 1;
      const uint32_t part = (uint32_t)b_half / (uint32_t)CS;     // this CTA's share of the rows of a multicast weight tile
      int mj = 0;
      for (int u = unit0; u < p.total_units; u = next_unit(u)) {
        const int n0 = HRP_TILE_N(u) * p.block_n;
        const int m0 = HRP_TILE_M(u) * TC_BLOCK_M;
        const int x0 = (m0 % a.Wo) * a.stride - a.pad_w, t1 = m0 / a.Wo, y0 = (t1 % a.Ho) * a.stride - a.pad_h, b0 = t1 / a.Ho;
        const uint8_t* wsrc = static_cast<const uint8_t*>(a.w) + (size_t)n0 * p.row_bytes;
        int c = 0, fr = 0, fs = 0;
        if (a_res && HRP_TILE_N(u) == 0) {
          // all k-blocks of this M tile's activations, once, into the A halves of stages 0..num_kb-1 (1x1 conv: one tap)
          if (mj >= 1) mbar_wait(bar_split + 8u, (uint32_t)((mj - 1) & 1));     // the previous M tile's last MMA has read them
          if (leader) {
            mbar_arrive_expect_tx(bar_split, (uint32_t)(p.num_kb * a_half));
            for (int kb = 0; kb < p.num_kb; ++kb) tma_load_4d(sA + (uint32_t)(kb * a_stage), p.tmap, kb * kbe, x0, y0, b0, bar_split);
          }
          ++mj;
        }
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          if (it >= S) mbar_wait(bar_empty + 8u * s, ph);
          const uint32_t bar = bar_full + 8u * s;
          if (leader) {
            mbar_arrive_expect_tx(bar, a_res ? (uint32_t)b_stage : tx);
            if (p.tma && !a_res) tma_load_4d(sA + (uint32_t)(s * a_stage), p.tmap, c, x0 + fs, y0 + fr, b0, bar);
            if (CS > 1) bulk_g2s_mc(sB + (uint32_t)(s * b_stage) + (uint32_t)crank * part, wsrc + kb * kb_stride + (size_t)crank * part, part, bar, cmask);
            else bulk_g2s(sB + (uint32_t)(s * b_stage), wsrc + kb * kb_stride, (uint32_t)b_half, bar);
            if (x3) bulk_g2s(sB + (uint32_t)(s * b_stage + b_half), wsrc + kb * kb_stride + (size_t)a.Cout * p.row_bytes, (uint32_t)b_half, bar);
          }
          if (p.tma) {
            c += kbe;
            if (c >= a.Cin) { c = 0; if (++fs == a.KW) { fs = 0; ++fr; } }
          }
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue warps 6..: TMEM -> registers -> (+bias, +residual, ReLU) -> global ================================
    // a warp may only touch TMEM lanes 32*(warp%4)..+31; thread (et, eh) owns tile row `et`, column half `eh`
    constexpr int EPI_T = 32 * EPI, EPI_H = EPI / 4;
    const int grp = (dual && warp < 4) ? 1 : 0;                        // dual mode (EPI == 4): warps 6-9 = group 0, warps 0-3 = group 1
    const int et = (warp & 3) * 32 + lane, eh = dual ? 0 : (warp - 6) >> 2, eid = eh * 128 + et;
    const int tile_inc = dual ? 2 : 1;
    auto epi_bar = [&]() {
      if (dual) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      else epi_barrier<EPI>();
    };
    const int c_lo = (p.block_n / EPI_H) * eh, c_hi = c_lo + p.block_n / EPI_H;   // block_n / EPI_H is a multiple of 16 (block_n >= 32)
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t pitch = (uint32_t)p.block_n * ESZ + 16u;            // odd multiple of 16 B: conflict-free both ways
    const uint32_t cpr_log = (uint32_t)p.cpr_log;                       // log2(16-byte chunks per tile row)
    const uint32_t cpr = 1u << cpr_log;
    const uint32_t my = stg + (uint32_t)et * pitch;
    const bool has_res = a.res != nullptr;
    // the layer's bias vector in shared memory, once per CTA: a global load per four columns inside the drain loop keeps
    // an in-order epilogue warp on the long scoreboard (measured on conv_roll.cu: the epilogue, not the MMAs, set the pace)
    const uint32_t s_bias = smem_base + (uint32_t)p.bias_off;
    for (int i = eid; i < a.Cout / 4; i += EPI_T) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias) + i);
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_bias + 16u * (uint32_t)i), "f"(b4.x), "f"(b4.y), "f"(b4.z), "f"(b4.w) : "memory");
    }
    epi_bar();
    int li = grp;
    for (int u = grp ? next_unit(unit0) : unit0; u < p.total_units; u = dual ? next_unit(next_unit(u)) : next_unit(u), li += tile_inc) {
      const int buf = li & 1;
      const int m0 = HRP_TILE_M(u) * TC_BLOCK_M, n0 = HRP_TILE_N(u) * p.block_n;
      // the row's pixel coordinates: three integer divisions per tile that only the register / cp.async epilogues need (the TMA
      // epilogue addresses the [M][Cout] matrix by m0 / n0 alone; they were 16 % of its instructions)
      bool row_ok = true;
      int ox = 0, oy = 0, b = 0, y = 0, x = 0;
      size_t pix = 0;
      if (!p.epi_tma || a.sa_partial != nullptr || a.out_nchw) {
        const int m = m0 + et;
        row_ok = m < p.M;
        const int mm = row_ok ? m : 0;
        ox = mm % a.Wo;
        const int t1 = mm / a.Wo;
        oy = t1 % a.Ho; b = t1 / a.Ho;
        y = oy * a.out_sy + a.out_oy; x = ox * a.out_sx + a.out_ox;
        pix = ((size_t)b * a.Ho_full + y) * a.Wo_full + x;
      }
      const uint32_t t_row = t_lane + (uint32_t)(buf * p.block_n);
      if (a.sa_partial != nullptr) {
        // The heatmap head without the heatmap: this tile is 128 pixels x the 64 depth bins of ONE keypoint. Every thread
        // owns a pixel, reduces its 64 logits to an online-softmax state (max, sum of exponentials, first moments in
        // x, y and depth), the tile's 128 states merge by shuffles and one shared-memory hop, and five floats leave
        // the SM instead of 32 KB of logits (integral.py:102-208; softargmax.cu finishes from the partials).
        mbar_wait(bar_accf + 8u * buf, (li >> 1) & 1);
        tc_fence_after();
        SaState st{-INFINITY, 0.f, 0.f, 0.f, 0.f};
        float vals[64];
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = lds_f4(s_bias + 4u * (uint32_t)(n0 + c0) + 16u * (uint32_t)q);
            vals[c0 + q * 4 + 0] = __uint_as_float(v[q * 4 + 0]) + bq.x; vals[c0 + q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + bq.y;
            vals[c0 + q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + bq.z; vals[c0 + q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + bq.w;
          }
        }
        tc_fence_before();
        if (row_ok) {
          float mx = vals[0];
#pragma unroll
          for (int d = 1; d < 64; ++d) mx = fmaxf(mx, vals[d]);
          float l = 0.f, sz = 0.f;
#pragma unroll
          for (int d = 0; d < 64; ++d) { const float e = __expf(vals[d] - mx); l += e; sz += (float)d * e; }
          st.m = mx; st.l = l; st.sz = sz; st.sx = (float)ox * l; st.sy = (float)oy * l;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { SaState o = sa_shfl_xor(st, off); sa_merge(st, o); }
        const uint32_t sa_smem = s_rowoff + 256u * (uint32_t)grp;  // 4 warps x 5 floats per epilogue group (the row-offset table is idle in this mode)
        if (lane == 0) {
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sa_smem + 32u * (uint32_t)(warp & 3)), "f"(st.m), "f"(st.l), "f"(st.sx), "f"(st.sy) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa_smem + 32u * (uint32_t)(warp & 3) + 16u), "f"(st.sz) : "memory");
        }
        epi_bar();
        if (eid == 0) {
          mbar_arrive(bar_acce + 8u * buf);
          SaState t{-INFINITY, 0.f, 0.f, 0.f, 0.f};
          for (int w = 0; w < 4; ++w) {
            SaState o;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.m), "=f"(o.l), "=f"(o.sx), "=f"(o.sy) : "r"(sa_smem + 32u * (uint32_t)w));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o.sz) : "r"(sa_smem + 32u * (uint32_t)w + 16u));
            sa_merge(t, o);
          }
          const int hw = a.Ho * a.Wo, chunks = hw / TC_BLOCK_M;
          const int bb = m0 / hw, tile_in_img = (m0 - bb * hw) / TC_BLOCK_M, kp = n0 >> 6;
          float* dst = a.sa_partial + ((size_t)(bb * (a.Cout >> 6) + kp) * chunks + tile_in_img) * 5;
          dst[0] = t.m; dst[1] = t.l; dst[2] = t.sx; dst[3] = t.sy; dst[4] = t.sz;
        }
        epi_bar();                                        // the shared-memory hop is reused by the next tile
        continue;
      }
      if (a.out_nchw) {
        // fp32 [B,Cout,Ho,Wo] (heatmap logits for the integral layer): lanes hold adjacent pixels, stores are coalesced
        mbar_wait(bar_accf + 8u * buf, (li >> 1) & 1);
        tc_fence_after();
        float* op = static_cast<float*>(a.out) + (((size_t)b * a.Cout + n0) * a.Ho_full + y) * a.Wo_full + x;
        const size_t cs = (size_t)a.Ho_full * a.Wo_full;
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = lds_f4(s_bias + 4u * (uint32_t)(n0 + c0) + 16u * (uint32_t)q);
            float f0 = __uint_as_float(v[q * 4 + 0]) + bq.x, f1 = __uint_as_float(v[q * 4 + 1]) + bq.y;
            float f2 = __uint_as_float(v[q * 4 + 2]) + bq.z, f3 = __uint_as_float(v[q * 4 + 3]) + bq.w;
            if (a.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
            op[(size_t)(c0 + q * 4 + 0) * cs] = f0; op[(size_t)(c0 + q * 4 + 1) * cs] = f1;
            op[(size_t)(c0 + q * 4 + 2) * cs] = f2; op[(size_t)(c0 + q * 4 + 3) * cs] = f3;
          }
        }
        tc_fence_before();
        epi_bar();
        if (eid == 0) mbar_arrive(bar_acce + 8u * buf);
        continue;
      }
      if (p.epi_tma) {
        // NHWC, dense output: the tile lives in shared memory as n_chunks boxes of [128 rows][chunk_bytes], 16-byte units
        // XOR-swizzled exactly as the TMA unit expects (SWIZZLE_128B / 64B), so threads that each own one tile ROW write
        // without bank conflicts. One thread pulls the residual boxes in (as soon as the previous store has finished
        // reading this buffer) and pushes the finished boxes out; rows beyond M are clipped by the TMA unit.
        const int sb = p.n_stg == 2 ? (li & 1) : 0;
        const uint32_t tile_stg = stg + (uint32_t)sb * (uint32_t)(p.n_chunks * TC_BLOCK_M * p.chunk_bytes);
        const uint32_t bar_res = sBar + 16u * TC_MAX_STAGES + 48u + 16u * TC_BLOCK_M + 8u * (uint32_t)sb;
        const uint32_t cbytes = (uint32_t)p.chunk_bytes, chunk_sz = (uint32_t)TC_BLOCK_M * cbytes;
        const uint32_t swz = cbytes == 128 ? ((uint32_t)et & 7u) : (((uint32_t)et >> 1) & 3u);
        // with two staging buffers the residual of tile li+1 is requested while tile li is drained, so that short-K tiles
        // (1x1 expansions: four MMAs per tile) do not sit out a load latency each. res_prefetch == 2: requested at the START
        // of tile li (after waiting for the store of tile li-1, issued a moment ago, to finish reading that buffer: a short
        // stall that buys the load the whole drain as head start); == 1: at the end of tile li (below), no stall
        const bool prefetch = has_res && p.res_prefetch != 0;
        const bool early = prefetch && p.res_prefetch == 2;
        if (eid == 0) {                                        // the store that last read this buffer has finished reading
          // (dual mode: a group's storing thread has only its own stores in flight, all from this one buffer)
          if (p.n_stg == 2 && !early && !dual) bulk_wait_read1(); else bulk_wait_read0();
          if (has_res && (!prefetch || li == 0)) {
            mbar_arrive_expect_tx(bar_res, (uint32_t)p.n_chunks * chunk_sz);
            for (int k = 0; k < p.n_chunks; ++k)
              tma_load_2d(tile_stg + (uint32_t)k * chunk_sz, p.tmap_res, a.out_coff + n0 + k * (int)(cbytes / ESZ), m0, bar_res);
          }
          const int next = next_unit(u);
          if (early && next < p.total_units) {
            const int m1 = HRP_TILE_M(next) * TC_BLOCK_M, n1 = HRP_TILE_N(next) * p.block_n;
            const uint32_t stg1 = stg + (uint32_t)(sb ^ 1) * (uint32_t)(p.n_chunks * TC_BLOCK_M * p.chunk_bytes);
            const uint32_t bar1 = bar_res - 8u * (uint32_t)sb + 8u * (uint32_t)(sb ^ 1);
            mbar_arrive_expect_tx(bar1, (uint32_t)p.n_chunks * chunk_sz);
            for (int k = 0; k < p.n_chunks; ++k)
              tma_load_2d(stg1 + (uint32_t)k * chunk_sz, p.tmap_res, a.out_coff + n1 + k * (int)(cbytes / ESZ), m1, bar1);
          }
        }
        epi_bar();
        mbar_wait(bar_accf + 8u * buf, (li >> 1) & 1);
        tc_fence_after();
        if (has_res) mbar_wait(bar_res, p.n_stg == 2 ? ((li >> 1) & 1) : (li & 1));
        bool drained = false;
        if constexpr (!TF32) {
          // 2-byte families, 32 columns per TMEM load: the drain is a chain of latencies for an in-order warp (tcgen05.ld,
          // the bias vector, one residual load per 16-byte unit, each waited for in turn: ~250 clk per 16 columns for ~90
          // issue slots). Here one tcgen05.ld.x32, the four residual units and the first half of the bias are all in
          // flight before the single wait, and there are half as many round trips. Same arithmetic per element.
          if (p.epi_wide) {
            drained = true;
            const uint32_t cb_log = cbytes == 128 ? 7u : 6u;
            for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(t_row + (uint32_t)c0, v);
              const uint32_t byte0 = (uint32_t)c0 * 2u, ck = byte0 >> cb_log, u0 = (byte0 & (cbytes - 1u)) >> 4;
              const uint32_t rowp = tile_stg + ck * chunk_sz + (uint32_t)et * cbytes;
              uint32_t addr[4], w[4][4];
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) addr[uu] = rowp + (((u0 + (uint32_t)uu) ^ swz) << 4);
              if (has_res) {
#pragma unroll
                for (int uu = 0; uu < 4; ++uu)
                  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[uu][0]), "=r"(w[uu][1]), "=r"(w[uu][2]), "=r"(w[uu][3]) : "r"(addr[uu]));
              }
              const uint32_t bp = s_bias + 4u * (uint32_t)(n0 + c0);
              float4 bq[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) bq[q] = lds_f4(bp + 16u * (uint32_t)q);
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                f[q * 4 + 0] = __uint_as_float(v[q * 4 + 0]) + bq[q].x; f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + bq[q].y;
                f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + bq[q].z; f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + bq[q].w;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) bq[q] = lds_f4(bp + 64u + 16u * (uint32_t)q);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                f[16 + q * 4 + 0] = __uint_as_float(v[16 + q * 4 + 0]) + bq[q].x; f[16 + q * 4 + 1] = __uint_as_float(v[16 + q * 4 + 1]) + bq[q].y;
                f[16 + q * 4 + 2] = __uint_as_float(v[16 + q * 4 + 2]) + bq[q].z; f[16 + q * 4 + 3] = __uint_as_float(v[16 + q * 4 + 3]) + bq[q].w;
              }
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) {
                float* ff = f + uu * 8;
                if (has_res) {
                  float r[8];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 t2 = unpack2<F16>(w[uu][e]);
                    r[2 * e] = t2.x; r[2 * e + 1] = t2.y;
                  }
                  if (!a.res_after_act) {                        // uniform branches, not per-element predicates
#pragma unroll
                    for (int e = 0; e < 8; ++e) ff[e] += r[e];
                    if (a.relu) {
#pragma unroll
                      for (int e = 0; e < 8; ++e) ff[e] = fmaxf(ff[e], 0.f);
                    }
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) ff[e] = (a.relu ? fmaxf(ff[e], 0.f) : ff[e]) + r[e];
                  }
                } else if (a.relu) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) ff[e] = fmaxf(ff[e], 0.f);
                }
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack2<F16>(ff[2 * e], ff[2 * e + 1]);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr[uu]), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
              }
            }
          }
        }
        for (int c0 = c_lo; c0 < c_hi && !drained; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = lds_f4(s_bias + 4u * (uint32_t)(n0 + c0) + 16u * (uint32_t)q);
            f[q * 4 + 0] = __uint_as_float(v[q * 4 + 0]) + bq.x;
            f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + bq.y;
            f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + bq.z;
            f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + bq.w;
          }
          // 16 columns = (16*ESZ)/16 sixteen-byte units starting at unit u0 of chunk ck
          const uint32_t byte0 = (uint32_t)c0 * ESZ, ck = byte0 / cbytes, u0 = (byte0 - ck * cbytes) >> 4;
          const uint32_t rowp = tile_stg + ck * chunk_sz + (uint32_t)et * cbytes;
          constexpr int UNITS = 16 * ESZ / 16;
#pragma unroll
          for (int uu = 0; uu < UNITS; ++uu) {
            const uint32_t addr = rowp + (((u0 + (uint32_t)uu) ^ swz) << 4);
            float* ff = f + uu * (16 / ESZ);
            if (has_res) {
              uint32_t w[4];
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
              float r[16 / ESZ];
              if constexpr (TF32) {
#pragma unroll
                for (int e = 0; e < 4; ++e) r[e] = __uint_as_float(w[e]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = unpack2<F16>(w[e]);
                  r[2 * e] = t2.x; r[2 * e + 1] = t2.y;
                }
              }
#pragma unroll
              for (int e = 0; e < 16 / ESZ; ++e) {
                float t = ff[e];
                if (!a.res_after_act) t += r[e];
                if (a.relu) t = fmaxf(t, 0.f);
                if (a.res_after_act) t += r[e];
                ff[e] = t;
              }
            } else if (a.relu) {
#pragma unroll
              for (int e = 0; e < 16 / ESZ; ++e) ff[e] = fmaxf(ff[e], 0.f);
            }
            uint32_t o[4];
            if constexpr (TF32) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = __float_as_uint(p.round_tf32 ? round_tf32_rna(ff[e]) : ff[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                o[e] = pack2<F16>(ff[2 * e], ff[2 * e + 1]);
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        
        }
        fence_proxy_async();                                   // generic-proxy writes -> visible to the TMA store
        tc_fence_before();
        epi_bar();
        if (eid == 0) {
          mbar_arrive(bar_acce + 8u * buf);                    // accumulator buffer may be overwritten by tile li+2
          for (int k = 0; k < p.n_chunks; ++k)
            tma_store_2d(p.tmap_out, a.out_coff + n0 + k * (int)(cbytes / ESZ), m0, tile_stg + (uint32_t)k * chunk_sz);
          bulk_commit();
          const int next = next_unit(u);
          if (prefetch && !early && next < p.total_units) {
            bulk_wait_read1();                                 // the store of tile li-1 has finished reading the other buffer
            const int m1 = HRP_TILE_M(next) * TC_BLOCK_M, n1 = HRP_TILE_N(next) * p.block_n;
            const uint32_t stg1 = stg + (uint32_t)(sb ^ 1) * (uint32_t)(p.n_chunks * TC_BLOCK_M * p.chunk_bytes);
            const uint32_t bar1 = bar_res - 8u * (uint32_t)sb + 8u * (uint32_t)(sb ^ 1);
            mbar_arrive_expect_tx(bar1, (uint32_t)p.n_chunks * chunk_sz);
            for (int k = 0; k < p.n_chunks; ++k)
              tma_load_2d(stg1 + (uint32_t)k * chunk_sz, p.tmap_res, a.out_coff + n1 + k * (int)(cbytes / ESZ), m1, bar1);
          }
        }
        continue;
      }
      // NHWC: the 128 x BLOCK_N tile passes through a shared-memory staging tile so that residual loads and output
      // stores run as 16-byte chunks along the channel axis; the residual is fetched while the MMAs are still running.
      if (eh == 0) {
        const long long off = row_ok ? (long long)(pix * a.ld_out + a.out_coff + n0) : -1ll;
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(s_rowoff + 8u * et), "l"(off) : "memory");
      }
      epi_bar();
      if (has_res) {
        const uint8_t* res8 = static_cast<const uint8_t*>(a.res);
        for (uint32_t idx = eid; idx < (128u << cpr_log); idx += EPI_T) {
          const uint32_t row = idx >> cpr_log, ch = idx & (cpr - 1u);
          long long off;
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(off) : "r"(s_rowoff + 8u * row));
          if (off >= 0) cp_async16(stg + row * pitch + (ch << 4), res8 + (size_t)off * ESZ + (ch << 4), 16u);
        }
        cp_async_commit();
        cp_async_wait<0>();
        epi_bar();
      }
      mbar_wait(bar_accf + 8u * buf, (li >> 1) & 1);
      tc_fence_after();
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bq = lds_f4(s_bias + 4u * (uint32_t)(n0 + c0) + 16u * (uint32_t)q);
          f[q * 4 + 0] = __uint_as_float(v[q * 4 + 0]) + bq.x;
          f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) + bq.y;
          f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) + bq.z;
          f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) + bq.w;
        }
        float r[16];
        if (has_res) {
          if constexpr (TF32) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[q * 4]), "=f"(r[q * 4 + 1]), "=f"(r[q * 4 + 2]), "=f"(r[q * 4 + 3]) : "r"(my + (uint32_t)c0 * 4u + 16u * q));
          } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t w[4];
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my + (uint32_t)c0 * 2u + 16u * q));
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 ff = unpack2<F16>(w[e]);
                r[q * 8 + e * 2] = ff.x; r[q * 8 + e * 2 + 1] = ff.y;
              }
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float t = f[e];
          if (has_res && !a.res_after_act) t += r[e];
          if (a.relu) t = fmaxf(t, 0.f);
          if (has_res && a.res_after_act) t += r[e];
          f[e] = t;
        }
        if constexpr (TF32) {
          if (p.round_tf32) {
#pragma unroll
            for (int e = 0; e < 16; ++e) f[e] = round_tf32_rna(f[e]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)c0 * 4u + 16u * q), "f"(f[q * 4]), "f"(f[q * 4 + 1]), "f"(f[q * 4 + 2]), "f"(f[q * 4 + 3]) : "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              w[e] = pack2<F16>(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1]);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)c0 * 2u + 16u * q), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
        }
      }
      tc_fence_before();
      epi_bar();
      if (eid == 0) mbar_arrive(bar_acce + 8u * buf);         // accumulator buffer may be overwritten by tile li+2
      uint8_t* out8 = static_cast<uint8_t*>(a.out);
      for (uint32_t idx = eid; idx < (128u << cpr_log); idx += EPI_T) {
        const uint32_t row = idx >> cpr_log, ch = idx & (cpr - 1u);
        long long off;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(off) : "r"(s_rowoff + 8u * row));
        if (off >= 0) {
          uint4 t;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(stg + row * pitch + (ch << 4)));
          *reinterpret_cast<uint4*>(out8 + (size_t)off * ESZ + (ch << 4)) = t;
        }
      }
      epi_bar();                                           // staging tile and row offsets are reused by the next tile
    }
    // the storing thread: its TMA stores have finished READING shared memory (the writes themselves are ordered before
    // the end of the grid like any other store; waiting for their completion here would only lengthen every CTA's tail)
    if (p.epi_tma && eid == 0) bulk_wait_read0();
  }
  tc_fence_before();
  if (CS > 1) cluster_sync_all(); else __syncthreads();     // a peer's last commits still arrive on this CTA's barriers
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
#undef HRP_TILE_M
#undef HRP_TILE_N
}

template <typename T>
__global__ void cast_f32_to_half_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) out[i] = __float2half_rn(fminf(fmaxf(in[i], -65504.f), 65504.f));
  else out[i] = __float2bfloat16_rn(in[i]);
}
template <typename T>
__global__ void cast_half_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if constexpr (std::is_same<T, __half>::value) out[i] = __half2float(in[i]); else out[i] = __bfloat162float(in[i]);
}
__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = round_tf32_rna(in[i]);
}

size_t tc_stage_bytes(int block_n, int row_bytes) { return (size_t)(TC_BLOCK_M + block_n) * row_bytes; }
size_t tc_staging_bytes(int block_n, int esz) { return ((size_t)TC_BLOCK_M * (block_n * esz + 16) + 127) / 128 * 128; }
size_t tc_tail_bytes() { return 16 * TC_MAX_STAGES + 16 + 16 + 16 + 8 * TC_BLOCK_M + 8 * TC_BLOCK_M + 16 + 8 * TC_MAX_STAGES + 48; }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {          // resolved through the runtime so the library has no link-time libcuda dependency
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      cudaGetLastError();
  }
  return fn;
}

// The 128 pixels of an M tile as a box of the NHWC input: {Wo-run, rows, images}. 0 if the geometry does not tile.
bool tma_box(const ConvArgs& a, int* bw, int* bh, int* bn_img) {
  if (a.Wo >= TC_BLOCK_M) {
    if (a.Wo % TC_BLOCK_M) return false;
    *bw = TC_BLOCK_M; *bh = 1; *bn_img = 1;
    return true;
  }
  if (TC_BLOCK_M % a.Wo) return false;
  const int rows = TC_BLOCK_M / a.Wo;
  if (rows <= a.Ho) {
    if (a.Ho % rows) return false;
    *bw = a.Wo; *bh = rows; *bn_img = 1;
  } else {
    if (rows % a.Ho) return false;
    *bw = a.Wo; *bh = a.Ho; *bn_img = rows / a.Ho;
  }
  return true;
}

}  // namespace

bool conv_tc_supported(const ConvArgs& a, int tf32) {
  const int half = tf32 ? 16 : 32;
  return a.Cin % half == 0 && a.Cout % 32 == 0 && a.ld_out % 8 == 0 && a.out_coff % 8 == 0 && a.KH * a.KW <= 32;
}

// Operand-row width a layer is packed for. TMA mode (one tap per k-block) needs Cin*elem to be 64 bytes or a multiple of
// 128 and the M tile to be a box of the input; everything else goes through the cp.async gather with 128-byte rows.
int conv_tc_row_bytes(const ConvArgs& a, int tf32, int* use_tma) {
  static const int no_tma = env_int("HRP_TC_NO_TMA", 0);
  const int esz = tf32 ? 4 : 2;
  int bw, bh, bi;
  const int cb = a.Cin * esz;
  bool ok = !no_tma && (cb == 64 || cb % 128 == 0) && tma_box(a, &bw, &bh, &bi) && bw * a.stride <= 256 && bh * a.stride <= 256 &&
            a.stride <= 8 && encode_tiled() != nullptr;
  if (use_tma) *use_tma = ok ? 1 : 0;
  return ok && cb == 64 ? 64 : 128;
}

int conv_tc_launch(const ConvArgs& a, int tf32, int round_tf32, cudaStream_t st) {
  if (!conv_tc_supported(a, tf32))
    return fail(HRP_ERR_INVALID, "conv_tc: unsupported shape Cin=%d Cout=%d ld=%d coff=%d", a.Cin, a.Cout, a.ld_out, a.out_coff);
  const bool x3 = tf32 && a.x3;
  if (!a.tma_custom && !x3 && conv_slab_supported(a, tf32)) return conv_slab_launch(a, tf32, round_tf32, st);
  TcParams p{};
  p.a = a;
  p.x3 = x3 ? 1 : 0;
  const int opnd = x3 ? 2 : 1;                              // operand tiles per stage
  p.M = a.B * a.Ho * a.Wo;
  if (p.M <= 0) return HRP_OK;
  const int esz = tf32 ? 4 : 2;
  p.row_bytes = conv_tc_row_bytes(a, tf32, &p.tma);
  if (a.tma_custom && !p.tma) return fail(HRP_ERR_INVALID, "conv_tc: a custom TMA view needs a TMA-tileable geometry");
  const int kb_elems = p.row_bytes / esz;
  p.Ktot = a.KH * a.KW * a.Cin;
  p.num_kb = ceil_div(p.Ktot, kb_elems);
  if (p.tma) {
    int bw = 0, bh = 0, bi = 0;
    tma_box(a, &bw, &bh, &bi);
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.B};
    cuuint64_t gstr[3] = {(cuuint64_t)a.Cin * esz, (cuuint64_t)a.Wi * a.Cin * esz, (cuuint64_t)a.Hi * a.Wi * a.Cin * esz};
    cuuint32_t box[4] = {(cuuint32_t)kb_elems, (cuuint32_t)(bw * a.stride), (cuuint32_t)(bh * a.stride), (cuuint32_t)bi};
    cuuint32_t estr[4] = {1, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1};
    if (a.tma_custom) {
      for (int i = 0; i < 4; ++i) { gdim[i] = a.tm_gdim[i]; box[i] = a.tm_box[i]; estr[i] = 1; }
      for (int i = 0; i < 3; ++i) gstr[i] = a.tm_gstr[i];
    }
    const CUresult r = encode_tiled()(&tm, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 4, const_cast<void*>(a.in),
                                      gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for input [%d,%d,%d,%d]", (int)r, a.B, a.Hi, a.Wi, a.Cin);
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    std::memcpy(p.tmap, &tm, 128);
  }
  const int mtiles = ceil_div(p.M, TC_BLOCK_M);
  const int sms = sm_count();
  // N tile: every CTA re-reads its activation rows, so wide tiles cut that traffic and run the MMAs at their full rate; take
  // the widest that still leaves about one tile per SM THIS LAUNCH MAY HOLD (inside the multi-lane graph a launch is capped
  // at a quarter of the GPU and the SMs are fully subscribed by the other lanes and plans, so what a layer costs is its
  // duration x the SMs it holds: at that cap the 2048->256 deconv phases @8x8 take 185 us on 37 SMs with 64-wide tiles and
  // 53 us on 32 SMs with 256-wide ones; profiles/r02_tile_width_at_quarter_gpu.txt), else the narrowest.
  // (TF32: <= 128 so the fp32 staging tile fits beside the stages.)
  static const int force_bn = env_int("HRP_TC_BN", 0), force_stages = env_int("HRP_TC_STAGES", 0), force_ctas = env_int("HRP_TC_CTAS", 0);
  static const int budget_kb = env_int("HRP_TC_BUDGET_KB", 0), bn_max = env_int("HRP_TC_BN_MAX", 256);
  static const int bn_full = env_int("HRP_TC_BN_FULL_GPU", 0);        // 1: size tiles as if the launch had the whole GPU (the old rule)
  const int sm_share = (a.grid_pct > 0 && !bn_full) ? std::max(1, sms * a.grid_pct / 100) : sms;
  const int cand[4] = {256, 128, 64, 32};
  int bn = 0;
  for (int i = tf32 ? 1 : 0; i < 4; ++i) {
    if (a.Cout % cand[i] || cand[i] > bn_max) continue;
    bn = cand[i];
    if ((long long)mtiles * (a.Cout / cand[i]) >= (sm_share * 4) / 5) break;
  }
  if (bn == 0) return fail(HRP_ERR_INVALID, "conv_tc: Cout=%d must be a multiple of 32", a.Cout);
  // very short K (1x1 expansions out of 64 / 128 channels): the tile is all epilogue and the layer is bound by the bytes it
  // writes, so two co-resident CTAs with 128-wide tiles keep more stores in flight than one CTA with a 256-wide tile
  // (measured: 64->256 @64x64 with residual 81 -> 75 us, without 60 -> 47 us)
  const bool short_k = p.Ktot <= 128 && bn > 128 && a.Cout % 128 == 0 && a.sa_partial == nullptr;
  if (short_k) bn = 128;
  // Epilogue-heavy layers in general (K <= 256 and 128 columns or more per tile; 2-byte families): the drain of a tile is a
  // latency chain of in-order warps, so what counts is how many tiles an SM drains at once. Two co-resident CTAs with 128-wide
  // tiles and four epilogue warps each (eight with the second epilogue group, epi_dual) beat one CTA with eight warps on one
  // 128- or 256-wide tile wherever there are enough tiles for two CTAs per SM (quarter-GPU cap, batch 64: 32->128 @64x64 with
  // residual 129 -> 100 us, 256->128 @64x64 104 -> 80 us, 256->1024 @16x16 with residual 67 -> 52 us, 512->2048 @8x8 with
  // residual 44 -> 32 us; profiles/r02_epilogue_groups.txt). HRP_TC_EPI_HEAVY=0 restores the old choice.
  static const int epi_heavy_on = env_int("HRP_TC_EPI_HEAVY", 1);
  bool epi_heavy = false;
  if (epi_heavy_on && !tf32 && !x3 && a.sa_partial == nullptr && !a.out_nchw && a.Cout % 128 == 0 && bn >= 128 &&
      (p.Ktot <= 256 || (p.Ktot <= 512 && a.res != nullptr && a.KH == 1)) && (long long)mtiles * (a.Cout / 128) >= 2LL * sms) {
    epi_heavy = true;
    bn = 128;
  }
  if (a.sa_partial != nullptr) {
    if (a.Cout % 64 || (a.Ho * a.Wo) % TC_BLOCK_M || a.out_sy != 1 || a.out_sx != 1 || a.Ho_full != a.Ho || a.Wo_full != a.Wo || a.res != nullptr)
      return fail(HRP_ERR_INVALID, "conv_tc: fused soft-argmax needs Cout %% 64 == 0 and whole 128-pixel tiles per frame (Cout=%d, %dx%d)", a.Cout, a.Ho, a.Wo);
    bn = 64;                                                // one N tile = the 64 depth bins of one keypoint
  }
  if (force_bn && a.Cout % force_bn == 0 && (!tf32 || force_bn <= 128) && a.sa_partial == nullptr) bn = force_bn;
  p.block_n = bn;
  p.tiles_n = a.Cout / bn;
  p.total_tiles = mtiles * p.tiles_n;
  int tm = 32;
  while (tm < 2 * bn) tm <<= 1;                             // two accumulator buffers
  p.tmem_cols = tm;
  // epilogue-heavy tiles (wide N, short K) get eight epilogue warps and the SM to themselves
  static const int force_epi = env_int("HRP_TC_EPI", 0);
  int epi = (bn >= 128 && p.Ktot <= 256 && !short_k && !epi_heavy) ? 8 : 4;
  if (force_epi == 4 || force_epi == 8) epi = force_epi;
  if (a.sa_partial != nullptr) epi = 4;
  static const int ctas_share = env_int("HRP_TC_CTAS_SHARE", 0);
  int ctas = (epi == 4 && p.total_tiles >= 2 * (ctas_share ? sm_share : sms) && tm <= 256) ? 2 : 1;
  if (force_ctas) ctas = (epi == 4 && force_ctas == 2 && tm <= 256) ? 2 : 1;
  size_t budget = (size_t)TC_SMEM_LIMIT / ctas - (ctas > 1 ? 512 : 0);       // two CTAs: 113 KB each + 1 KB reserved = 228 KB
  if (budget_kb > 0 && a.grid_pct > 0) budget = std::min(budget, (size_t)budget_kb * 1024);   // experiment: leave room for a CTA of another lane's kernel
  // TMA epilogue: dense NHWC output whose tile rows are consecutive rows of the [M][Cout] matrix
  static const int no_epi_tma = env_int("HRP_TC_NO_EPI_TMA", 0);
  p.epi_tma = !no_epi_tma && !a.out_nchw && a.out_sy == 1 && a.out_sx == 1 && a.out_oy == 0 && a.out_ox == 0 && a.Ho_full == a.Ho &&
              a.Wo_full == a.Wo && a.ld_out == a.Cout && a.out_coff == 0 && (bn * esz) % 64 == 0 && encode_tiled() != nullptr;
  size_t staging = tc_staging_bytes(bn, esz);
  if (p.epi_tma) {
    p.chunk_bytes = (bn * esz) % 128 == 0 ? 128 : 64;
    p.n_chunks = bn * esz / p.chunk_bytes;
    const size_t one = (size_t)TC_BLOCK_M * bn * esz;
    // two staging buffers when at least three operand stages still fit beside them. Short-K residual layers (1x1
    // expansions: one or two k-blocks per tile, so the tile is all epilogue and the MMAs of the next tile run under it
    // anyway) trade operand stages for the second buffer, which lets the residual of tile li+1 load during tile li and
    // the store of tile li drain under tile li+1 (64->256 @64x64: 78 -> 70.5 us with residual, 46.8 -> 39.1 us without).
    static const int no_short_stg = env_int("HRP_TC_NO_SHORT_STG2", 0);
    // (epilogue-heavy layers with one or two k-blocks per tile likewise; with more k-blocks a single stage would serialise
    // every load with its MMAs, so those keep two stages and one staging buffer)
    const int min_stages = ((short_k || (epi_heavy && p.num_kb <= 2)) && !no_short_stg) ? 1 : 3;
    p.n_stg = (2048 + 2 * one + tc_tail_bytes() + (size_t)a.Cout * 4 + min_stages * opnd * tc_stage_bytes(bn, p.row_bytes) <= budget) ? 2 : 1;
    staging = (p.n_stg * one + 1023) / 1024 * 1024;
    static const int no_prefetch = env_int("HRP_TC_NO_RES_PREFETCH", 0);
    static const int early = env_int("HRP_TC_RES_EARLY", 1);
    p.res_prefetch = (p.n_stg == 2 && a.res != nullptr && !no_prefetch) ? (early ? 2 : 1) : 0;
    static const int epi_dual = env_int("HRP_TC_EPI_DUAL", 1);
    p.epi_dual = (epi_dual && epi == 4 && p.tma && !x3 && p.n_stg == 2 && a.sa_partial == nullptr && !a.out_nchw) ? 1 : 0;
    if (p.epi_dual) p.res_prefetch = 0;                       // one staging buffer per group: the residual is requested when the tile starts
    static const int epi_wide = env_int("HRP_TC_EPI_WIDE", 1);
    p.epi_wide = (epi_wide && !tf32 && (bn / (epi / 4)) % 32 == 0) ? 1 : 0;
  }
  if (a.sa_partial != nullptr) {                              // the fused heatmap head: 64 exponentials per thread and tile, no staging at all
    static const int sa_dual = env_int("HRP_TC_SA_DUAL", 1);
    p.epi_dual = (sa_dual && epi == 4 && p.tma && !x3) ? 1 : 0;
    staging = 0;
  }
  const size_t fixed = 2048 + staging + tc_tail_bytes() + (size_t)a.Cout * 4;
  int smax = (int)((budget - fixed) / (opnd * tc_stage_bytes(bn, p.row_bytes)));
  smax = std::max(1, std::min(smax, TC_MAX_STAGES));
  if (force_stages) smax = std::max(1, std::min(force_stages, smax));
  p.stages = smax;                                          // the ring runs across tiles, so depth is not tied to num_kb
  const long long kb_per_cta = (long long)p.num_kb * ceil_div(p.total_tiles, sms * ctas);
  if (kb_per_cta < p.stages) p.stages = (int)std::max(1LL, kb_per_cta);
  // Activation-resident walk for the fused heatmap head (HRP_TC_A_RES): its N tiles are the keypoints, and walking tiles N-fastest
  // across CTAs made every keypoint's tile fetch the same 128 x 256-channel activation block again -- 1.5 GB of L2 requests per
  // 64 frames, ~63 B/clk per SM, which is what the layer ran at. A CTA now owns whole M tiles: the block's k-blocks land once in
  // the A halves of stages 0..num_kb-1 and the ring streams the keypoints' weight tiles (672 -> 288 KB per M tile).
  // Measured (scripts/experiments/a_res.sh): results identical, the layer alone on the whole GPU 126 -> 142 us (the single resident
  // block leaves the MMA warp idle while the next M tile's activations load, and alone the layer is not L2-bound), the network
  // 13.85k vs 13.90k frames/s with and without (two interleaved runs each): kept as a switch, off by default.
  static const int a_res_on = env_int("HRP_TC_A_RES", 0);   // off: correct (tests pass with it) but not faster, see the comment above
  p.a_res = (a_res_on && a.sa_partial != nullptr && p.tma && !x3 && !tf32 && a.KH == 1 && a.KW == 1 && a.stride == 1 && p.tiles_n >= 2 &&
             p.stages >= p.num_kb && p.num_kb <= TC_MAX_STAGES) ? 1 : 0;
  p.lag = std::min(TC_MAX_LAG, p.stages - 1);
  p.round_tf32 = x3 ? 0 : round_tf32;                       // 3xTF32 layers exchange full fp32 activations
  p.stg_off = (int)((p.stages * opnd * tc_stage_bytes(bn, p.row_bytes) + 1023) / 1024 * 1024);
  p.bar_off = p.stg_off + (int)staging;
  int cl = 0;
  while ((16 << cl) < bn * esz) ++cl;
  p.cpr_log = cl;
  p.bias_off = p.bar_off + (int)tc_tail_bytes();
  const size_t smem = 1024 + (size_t)p.bias_off + (size_t)a.Cout * 4;
  if (p.epi_tma) {
    const cuuint64_t gdim[2] = {(cuuint64_t)a.ld_out, (cuuint64_t)p.M};
    const cuuint64_t gstr[1] = {(cuuint64_t)a.ld_out * esz};
    const cuuint32_t box[2] = {(cuuint32_t)(p.chunk_bytes / esz), (cuuint32_t)TC_BLOCK_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    const CUtensorMapSwizzle sw = p.chunk_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmo, tmr;
    CUresult r = encode_tiled()(&tmo, dt, 2, a.out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && a.res)
      r = encode_tiled()(&tmr, dt, 2, const_cast<void*>(a.res), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for output [%d,%d]", (int)r, p.M, a.ld_out);
    std::memcpy(p.tmap_out, &tmo, 128);
    if (a.res) std::memcpy(p.tmap_res, &tmr, 128);
  }
  if (smem > (size_t)TC_SMEM_LIMIT) return fail(HRP_ERR_INVALID, "conv_tc: %zu bytes of shared memory needed (block_n %d)", smem, bn);
  static bool attr_done = false;
  if (!attr_done) {
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_done = true;
  }
  // Thread-block clusters (HRP_TC_CLUSTER=2|4, off by default): `cs` CTAs take consecutive M tiles of one N tile and share
  // every weight tile through one multicast fetch. Weights re-read by every M tile are 27 % of the network's L2 request
  // bytes, but halving / quartering them left the whole-network rate where it was (13.40k / 13.41k / 13.43k frames/s for
  // cluster 1 / 2 / 4 on the same box) and made single layers 2-5 % (cluster 2) to 50 % (cluster 4, short-K layers) slower
  // because the CTAs of a cluster advance in lockstep: the multi-lane graph is bound by SM time, not by L2 requests
  // (profiles/r02_throughput_experiments.txt). Kept as a measured switch. Needs a whole number of M-tile groups.
  static const int cl_env = env_int("HRP_TC_CLUSTER", 1);
  const int cap = std::max(1, sms * ctas * (a.grid_pct > 0 ? a.grid_pct : 100) / 100);
  int cs = 1;
  if (!x3)
    for (int c = std::min(cl_env, 4); c >= 2; c >>= 1)
      if ((c & (c - 1)) == 0 && mtiles % c == 0 && cap >= c && (bn / c) % 8 == 0) { cs = c; break; }
  p.cluster = cs;
  p.total_units = (mtiles / cs) * p.tiles_n;
  if (cs == 1) p.total_units = p.total_tiles;
  if (cs > 1) p.a_res = 0;
  int grid = std::min(p.total_units * cs, cap);
  grid -= grid % cs;
  if (p.a_res) grid = std::min(mtiles, cap);                 // one unit of work = an M tile with all its N tiles
  size_t smem_req = smem;
  if (cs > 1) {
    // A cluster CTA that holds TMEM may wait for a peer that is still waiting for TMEM on another SM, so cluster CTAs must
    // never oversubscribe an SM's 512 columns among themselves: each asks for at least its TMEM share of the SM's shared
    // memory (228 KB, 1 KB reserved per CTA), which makes the co-residency limit the stricter of the two.
    const size_t share = ((size_t)228 * 1024 * tm + 511) / 512;
    smem_req = std::min((size_t)TC_SMEM_LIMIT, std::max(smem, share > 1024 ? share - 1024 : 0));
  }
  static const int dbg = env_int("HRP_TC_DEBUG", 0);
  if (dbg) fprintf(stderr, "conv_tc %dx%d %d->%d k%d s%d res%d: bn %d epi %d ctas %d grid %d stages %d n_stg %d smem %zu tiles %d dual %d a_res %d\n", a.Hi, a.Wi, a.Cin, a.Cout, a.KH,
                   a.stride, a.res != nullptr, bn, epi, ctas, grid, p.stages, p.n_stg, smem_req, p.total_tiles, p.epi_dual, p.a_res);
  static const int pdl_early = env_int("HRP_PDL_EARLY", 0);
  p.pdl_early = pdl_early;
  cudaError_t le;
  if (tf32) le = epi == 8 ? launch_pdl(conv_tc_kernel<true, 8, false>, grid, tc_threads(8), smem_req, st, p, cs) : launch_pdl(conv_tc_kernel<true, 4, false>, grid, tc_threads(4), smem_req, st, p, cs);
  else if (a.f16) le = epi == 8 ? launch_pdl(conv_tc_kernel<false, 8, true>, grid, tc_threads(8), smem_req, st, p, cs) : launch_pdl(conv_tc_kernel<false, 4, true>, grid, tc_threads(4), smem_req, st, p, cs);
  else le = epi == 8 ? launch_pdl(conv_tc_kernel<false, 8, false>, grid, tc_threads(8), smem_req, st, p, cs) : launch_pdl(conv_tc_kernel<false, 4, false>, grid, tc_threads(4), smem_req, st, p, cs);
  if (le != cudaSuccess) return fail(HRP_ERR_CUDA, "conv_tc_kernel launch: %s", cudaGetErrorString(le));
  HRP_CHECK_LAUNCH("conv_tc_kernel");
  return HRP_OK;
}

// [K][Cout] fp32 (BN folded, k = (r*KW+s)*Cin + c) -> [num_kb][Cout][row_bytes] K-major rows, 16-byte chunks XOR-swizzled
// exactly as the SWIZZLE_128B (chunk ^ (row & 7)) / SWIZZLE_64B (chunk ^ ((row >> 1) & 3)) operand layouts expect; K
// zero-padded to a whole k-block.
// tf32 == 2: the 3xTF32 image [num_kb][hi | lo][Cout][row_bytes], hi = rna_tf32(w), lo = rna_tf32(w - hi).
// tf32 == 3: 2-byte elements like 0, IEEE half (round to nearest even, saturating) instead of bf16.
size_t pack_conv_tc_bytes(int K, int Cout, int tf32, int row_bytes) {
  const bool four = tf32 == 1 || tf32 == 2;
  return (size_t)ceil_div(K, row_bytes / (four ? 4 : 2)) * Cout * row_bytes * (tf32 == 2 ? 2 : 1);
}

namespace {
uint16_t f32_to_f16_bits(float v) {          // round to nearest even, saturate to +-65504, subnormals kept
  uint32_t u;
  std::memcpy(&u, &v, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t absu = u & 0x7fffffffu;
  if (absu >= 0x7f800000u) return (uint16_t)(sign | (absu > 0x7f800000u ? 0x7e00u : 0x7bffu));   // NaN / inf (saturated)
  float av;
  std::memcpy(&av, &absu, 4);
  if (av >= 65520.f) return (uint16_t)(sign | 0x7bffu);
  if (av < 6.103515625e-5f) {                                   // subnormal half: multiples of 2^-24
    const float scaled = av * 16777216.f;                       // exact
    const uint32_t m = (uint32_t)std::nearbyint(scaled);        // round to nearest even (default rounding mode)
    return (uint16_t)(sign | m);
  }
  uint32_t r = absu + 0xfffu + ((absu >> 13) & 1u);             // round the 13 dropped bits to nearest even
  r = ((r >> 13) - (112u << 10));                               // rebias 127 -> 15
  return (uint16_t)(sign | r);
}
}  // namespace

void pack_conv_tc(const float* w_kn, int K, int Cout, int tf32, int row_bytes, void* out) {
  const bool f16 = tf32 == 3;
  if (f16) tf32 = 0;
  const int ce = tf32 ? 4 : 8, chunks = row_bytes / 16, kb_elems = chunks * ce;
  const int num_kb = ceil_div(K, kb_elems);
  uint8_t* o = static_cast<uint8_t*>(out);
  std::memset(o, 0, pack_conv_tc_bytes(K, Cout, f16 ? 3 : tf32, row_bytes));
  const bool x3 = tf32 == 2;
  auto rna = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & ~0x1fffu; float r; std::memcpy(&r, &u, 4); return r; };
  for (int kb = 0; kb < num_kb; ++kb)
    for (int n = 0; n < Cout; ++n) {
      uint8_t* row = o + ((size_t)kb * (x3 ? 2 : 1) * Cout + n) * row_bytes;
      const int sw = row_bytes == 128 ? (n & 7) : ((n >> 1) & 3);
      for (int j = 0; j < chunks; ++j) {
        uint8_t* chunk = row + ((j ^ sw) << 4);
        for (int e = 0; e < ce; ++e) {
          const int k = kb * kb_elems + j * ce + e;
          if (k >= K) continue;
          const float v = w_kn[(size_t)k * Cout + n];
          if (tf32) {
            const float hi = rna(v);                                                  // cvt.rna.tf32
            std::memcpy(chunk + e * 4, &hi, 4);
            if (x3) { const float lo = rna(v - hi); std::memcpy(chunk + (size_t)Cout * row_bytes + e * 4, &lo, 4); }
          } else if (f16) {
            const uint16_t hbits = f32_to_f16_bits(v);
            std::memcpy(chunk + e * 2, &hbits, 2);
          } else {
            uint32_t u;
            std::memcpy(&u, &v, 4);
            if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);   // round to nearest even
            const uint16_t hbits = (uint16_t)(u >> 16);
            std::memcpy(chunk + e * 2, &hbits, 2);
          }
        }
      }
    }
}

int cast_f32_to_bf16_launch(const float* in, void* out, size_t n, cudaStream_t st, int f16) {
  if (n == 0) return HRP_OK;
  if (f16) cast_f32_to_half_kernel<__half><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, static_cast<__half*>(out), n);
  else cast_f32_to_half_kernel<__nv_bfloat16><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(out), n);
  HRP_CHECK_LAUNCH("cast_f32_to_half_kernel");
  return HRP_OK;
}
int cast_bf16_to_f32_launch(const void* in, float* out, size_t n, cudaStream_t st, int f16) {
  if (n == 0) return HRP_OK;
  if (f16) cast_half_to_f32_kernel<__half><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const __half*>(in), out, n);
  else cast_half_to_f32_kernel<__nv_bfloat16><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), out, n);
  HRP_CHECK_LAUNCH("cast_half_to_f32_kernel");
  return HRP_OK;
}
int round_tf32_launch(const float* in, float* out, size_t n, cudaStream_t st) {
  if (n == 0) return HRP_OK;
  round_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
  HRP_CHECK_LAUNCH("round_tf32_kernel");
  return HRP_OK;
}

}  // namespace hrp
