// One HRNet BasicBlock -- relu(bn2(conv2(relu(bn1(conv1(x))))) + x), both convs 3x3/s1/p1 with Cin == Cout
// (HRnet.py:28-57) -- as ONE kernel: two shifted GEMMs (conv_slab.cu explains the idea) chained through shared memory.
//
// A unit of 128*n consecutive padded positions needs the intermediate y on those positions plus a (W+2)+1 halo either
// side, i.e. n1 = ceil((128n + 2(W+2) + 2) / 128) blocks starting at Y0 = Q0 - (W+2) - 1, which in turn need the input slab
// one more halo out. Per unit:
//   loader   : TMA row boxes of x -> slab X (zero halo by TMA)                      [prefetched while phase 2 of the
//   phase 1  : n1 blocks, 9 taps each, A = shifted slab X, B = W1 -> TMEM            previous unit still runs]
//   epilogue1: TMEM -> +b1, ReLU, ZERO outside the image (conv2's padding) -> bf16 -> slab Y, written in the operand
//              swizzle by the thread that owns the row, fence.proxy.async, per-block mbarrier
//   phase 2  : n blocks, A = shifted slab Y (as soon as the Y blocks a block touches are written), B = W2 -> TMEM
//   epilogue2: TMEM -> +b2, + x (cp.async from L2), ReLU -> coalesced 16-byte stores
// The intermediate never reaches HBM/L2 and one launch (prologue, weight staging, tail) disappears per block; phase 1
// recomputes the halo blocks (n1/n: 8/6 at 64x64). Results are bit-identical to the two-kernel path (same operands, same
// K order, same bf16 rounding of the intermediate). bf16, 32 channels (operand rows of 64 bytes): the layers of the
// full-resolution HRNet branch, the critical path of the multi-lane graph. Measured: +1.7 % frames/s on the full network.
// A variant that stacks the three horizontal taps along N (six N=96 MMAs per block instead of eighteen N=32 ones, slices
// recombined with warp shuffles in the epilogue) is correct but slower -- the epilogue becomes the long pole -- and is
// kept only as scripts/experiments/conv_block_nstacked.cu.txt.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_ptx.h"

namespace hrp {
namespace {
using namespace tc;

constexpr int BK_MAX_GROUPS = 3;
// MMA issuer warps: one elected thread needs ~70 clk of issue slots per tcgen05.mma (descriptor arithmetic in uniform
// registers, R2UR moves), more than the ~44 clk an M128 N32 K16 MMA takes (conv_roll.cu / conv_chain.cu measured the same
// effect). Blocks have separate accumulators, so issuer w takes blocks w, w + 2, ... of the unit's phase-1 / phase-2 sequence.
constexpr int BK_ISSUERS = 2;
constexpr int BK_FIRST_EPI = 1 + BK_ISSUERS;
constexpr int BK_MAX_N1 = 12;
constexpr int BK_SMEM_LIMIT = 227 * 1024;

struct BlockParams {
  alignas(64) unsigned char tmap[3][128];   // NHWC x as {C, W, H, B}; boxes {C, W+2, R, 1}, R = 16, 4, 1
  BlockArgs a;
  int C, Wp, Hp, HpWp;
  long long Q;
  int n, n1, kmax, units;
  int x_bytes, y_off, w_off, w_bytes, stg_off, bar_off;
  int groups, nacc, tmem_cols, cpr_log;
};

// barrier block (8-byte slots): w | x_full | x_empty | y_empty | y_full[12] | acc_full[6] | acc_empty[6] | tmem slot
template <int ROWB, bool F16>
__global__ void __launch_bounds__(32 * BK_FIRST_EPI + 128 * BK_MAX_GROUPS, 1)
conv_block_kernel(const __grid_constant__ BlockParams p) {
  constexpr int ESZ = 2;
  constexpr int KSTEPS = ROWB / 32;
  constexpr int UNITS = ROWB / 16;            // 16-byte units per operand row
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base, sY = base + (uint32_t)p.y_off, sW = base + (uint32_t)p.w_off, stg = base + (uint32_t)p.stg_off,
                 sBar = base + (uint32_t)p.bar_off;
  const uint32_t bar_w = sBar, bar_xf = sBar + 8u, bar_xe = sBar + 16u, bar_ye = sBar + 24u, bar_yf = sBar + 32u,
                 bar_af = bar_yf + 8u * BK_MAX_N1, bar_ae = bar_af + 48u, tmem_slot = bar_ae + 48u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BlockArgs& a = p.a;
  const int C = p.C, Wp = p.Wp, G = p.groups, NACC = p.nacc, n = p.n, n1 = p.n1;

  if (tid == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_xf, 1); mbar_init(bar_xe, BK_ISSUERS); mbar_init(bar_ye, BK_ISSUERS);
    for (int i = 0; i < n1; ++i) mbar_init(bar_yf + 8u * i, 4);
    for (int i = 0; i < NACC; ++i) { mbar_init(bar_af + 8u * i, 1); mbar_init(bar_ae + 8u * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane < 3) asm volatile("prefetch.tensormap [%0];" ::"l"(p.tmap[lane]) : "memory");
  // both bias vectors go to shared memory once per CTA (b1 at +0, b2 at +4C bytes): a global load per channel pair inside the
  // drain loops keeps the in-order epilogue warps on the long scoreboard
  const uint32_t s_bias = sBar + 256u;
  if (tid >= 64 && tid - 64 < 2 * C) {
    const float bvv = __ldg((tid - 64 < C ? a.b1 : a.b2 - C) + (tid - 64));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias + 4u * (uint32_t)(tid - 64)), "f"(bvv) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int unit_pos = 128 * n;
  // unit geometry: first padded row of slab X for unit u
  auto slab_row0 = [&](int u) -> int {
    const long long lo = (long long)u * unit_pos - 2LL * Wp - 2;
    return (int)(lo >= 0 ? lo / Wp : -((-lo + Wp - 1) / Wp));
  };

  if (warp == 0) {
    // ===== loader ========================================================================================================
    const bool leader = elect_one();
    if (leader) mbar_arrive_expect_tx(bar_w, 2u * (uint32_t)p.w_bytes);
    const int tap_bytes = p.w_bytes / 9;
    for (int t = 0; t < 9; ++t)
      if (leader) {
        bulk_g2s(sW + (uint32_t)(t * tap_bytes), static_cast<const uint8_t*>(a.w1) + (size_t)t * tap_bytes, (uint32_t)tap_bytes, bar_w);
        bulk_g2s(sW + (uint32_t)(p.w_bytes + t * tap_bytes), static_cast<const uint8_t*>(a.w2) + (size_t)t * tap_bytes, (uint32_t)tap_bytes, bar_w);
      }
    const int total_rows = a.B * p.Hp;
    const uint32_t row_tx = (uint32_t)(Wp * ROWB);
    int li = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++li) {
      if (li >= 1) mbar_wait(bar_xe, (li - 1) & 1);            // phase 1 of the previous unit has read slab X
      const int ra = slab_row0(u);
      const long long hi = (long long)u * unit_pos - Wp - 1 + 128LL * n1 + Wp + 1;
      const int rb = (int)((hi + Wp - 1) / Wp);
      const int r0 = ra < 0 ? 0 : ra, r1 = rb > total_rows ? total_rows : rb;
      if (leader) mbar_arrive_expect_tx(bar_xf, (uint32_t)(r1 - r0) * row_tx);
      int R = r0;
      while (R < r1) {
        const int b = R / p.Hp, yy = R - b * p.Hp;
        const int seg = min(r1 - R, p.Hp - yy);
        int done = 0;
        while (done < seg) {
          const int left = seg - done;
          const int m = left >= 16 ? 0 : (left >= 4 ? 1 : 2), rows = left >= 16 ? 16 : (left >= 4 ? 4 : 1);
          if (leader) tma_load_4d(sX + (uint32_t)(R + done - ra) * row_tx, p.tmap[m], 0, -1, yy + done - 1, b, bar_xf);
          done += rows;
        }
        R += seg;
      }
    }
  } else if (warp < BK_FIRST_EPI) {
    // ===== MMA issuers (all lanes walk the loops, one elected lane issues; tc_ptx.h); issuer iw owns every BK_ISSUERS-th block ==
    const bool leader = elect_one();
    const int iw = warp - 1;
    int seq = 0;                                                  // running block number (phase 1 and phase 2 blocks of all units)
    constexpr uint32_t FMT = F16 ? 0u : 1u;                     // operand format: IEEE half / bf16
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dhi = umma_desc_hi(ROWB);
    const uint32_t wt16 = (uint32_t)(C * ROWB) >> 4, w1_16 = (sW >> 4) | (1u << 16), w2_16 = ((sW + (uint32_t)p.w_bytes) >> 4) | (1u << 16);
    uint32_t tap16[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tap16[t] = (uint32_t)(((t / 3) * Wp + (t % 3)) * (ROWB / 16));
    mbar_wait(bar_w, 0);
    int li = 0, ab = 0, use = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++li) {
      mbar_wait(bar_xf, li & 1);
      tc_fence_after();
      // phase 1: Y block i, tap (0,0) reads slab X row (Y0 + 128 i - Wp - 1) - ra*Wp
      const long long y0 = (long long)u * unit_pos - Wp - 1;
      uint32_t a16 = ((sX >> 4) + (uint32_t)(y0 - Wp - 1 - (long long)slab_row0(u) * Wp) * (ROWB / 16)) | (1u << 16);
      for (int i = 0; i < n1; ++i, a16 += 128u * (ROWB / 16), ++seq) {
        const bool mine = seq % BK_ISSUERS == iw;
        if (!mine) {                                              // the other issuer's block: only its X-slab release is shared
          if (i == n1 - 1 && leader) umma_commit(bar_xe);         // (this issuer's phase-1 MMAs of the unit have been issued)
          __syncwarp();
          if (++ab == NACC) { ab = 0; ++use; }
          continue;
        }
        if (use >= 1) mbar_wait(bar_ae + 8u * ab, (use - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ab * C);
        if (leader) {
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk)
              umma_lo<false>(tmem_d, a16 + tap16[t] + 2u * kk, w1_16 + (uint32_t)t * wt16 + 2u * kk, dhi, idesc, (t | kk) != 0 ? 1u : 0u);
          umma_commit(bar_af + 8u * ab);
          if (i == n1 - 1) umma_commit(bar_xe);                 // slab X may be refilled (the residual comes from L2)
        }
        __syncwarp();
        if (++ab == NACC) { ab = 0; ++use; }
      }
      // phase 2: output block j, tap (r,s) reads slab Y row 128 j + r*Wp + s; needs Y blocks j .. j + kmax
      uint32_t y16 = (sY >> 4) | (1u << 16);
      int ready = -1;                                           // Y blocks 0..ready are known to be written
      for (int j = 0; j < n; ++j, y16 += 128u * (ROWB / 16), ++seq) {
        if (seq % BK_ISSUERS != iw) {
          if (j == n - 1 && leader) umma_commit(bar_ye);
          __syncwarp();
          if (++ab == NACC) { ab = 0; ++use; }
          continue;
        }
        const int need = min(j + p.kmax, n1 - 1);
        for (; ready < need; ++ready) mbar_wait(bar_yf + 8u * (ready + 1), li & 1);
        if (use >= 1) mbar_wait(bar_ae + 8u * ab, (use - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ab * C);
        if (leader) {
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk)
              umma_lo<false>(tmem_d, y16 + tap16[t] + 2u * kk, w2_16 + (uint32_t)t * wt16 + 2u * kk, dhi, idesc, (t | kk) != 0 ? 1u : 0u);
          umma_commit(bar_af + 8u * ab);
          if (j == n - 1) umma_commit(bar_ye);                  // slab Y may be overwritten by the next unit's phase 1
        }
        __syncwarp();
        if (++ab == NACC) { ab = 0; ++use; }
      }
    }
  } else if (warp < BK_FIRST_EPI + 4 * G) {
    // ===== epilogue warps: each its own pipeline over the 32 rows of its TMEM lane quarter; group g takes every G-th block =
    const int e = warp - BK_FIRST_EPI, g = e >> 2, quarter = warp & 3;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t pitch = (uint32_t)C * ESZ + 16u;
    const uint32_t cpr_log = (uint32_t)p.cpr_log, cpr = 1u << cpr_log;
    const uint32_t wst = stg + (uint32_t)e * 32u * pitch;
    const uint32_t my = wst + (uint32_t)lane * pitch;
    const uint8_t* x8 = static_cast<const uint8_t*>(a.x);
    uint8_t* out8 = static_cast<uint8_t*>(a.out);
    const int Qi = (int)p.Q, HpWp = p.HpWp;
    const size_t row_b = (size_t)C * ESZ;
    int ab = 0, use = 0, gsel = 0, li = 0;
    auto interior = [&](int q, int* pix) -> bool {
      if (q < 0 || q >= Qi) { *pix = -1; return false; }
      const int b = q / HpWp, rem = q - b * HpWp, yy = rem / Wp, xx = rem - yy * Wp;
      const bool ok = yy >= 1 && yy <= a.H && xx >= 1 && xx <= a.W;
      *pix = ok ? (b * a.H + (yy - 1)) * a.W + (xx - 1) : -1;
      return ok;
    };
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++li) {
      const int q0 = u * unit_pos, y0 = q0 - Wp - 1;
      for (int blk = 0; blk < n1 + n; ++blk) {
        if (gsel == g) {
          const uint32_t t_row = t_lane + (uint32_t)(ab * C);
          if (blk < n1) {
            // ---- epilogue 1: y = relu(acc + b1) inside the image, 0 outside, into slab Y row R (operand swizzle) -------------
            const int R = 128 * blk + quarter * 32 + lane;
            int pix;
            const bool ok = interior(y0 + R, &pix);
            if (li >= 1) mbar_wait(bar_ye, (li - 1) & 1);       // phase 2 of the previous unit has read slab Y
            mbar_wait(bar_af + 8u * ab, use & 1);
            tc_fence_after();
            const uint32_t yrow = sY + (uint32_t)R * ROWB;
            const uint32_t swz = ROWB == 128 ? ((uint32_t)R & 7u) : (((uint32_t)R >> 1) & 3u);
            for (int c0 = 0; c0 < C; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(t_row + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 bq;
                  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bq.x), "=f"(bq.y) : "r"(s_bias + 4u * (uint32_t)(c0 + q2 * 8 + k * 2)));
                  const float f0 = ok ? fmaxf(__uint_as_float(v[q2 * 8 + k * 2]) + bq.x, 0.f) : 0.f;
                  const float f1 = ok ? fmaxf(__uint_as_float(v[q2 * 8 + k * 2 + 1]) + bq.y, 0.f) : 0.f;
                  w[k] = pack2<F16>(f0, f1);
                }
                const uint32_t un = (uint32_t)(c0 * ESZ) / 16u + (uint32_t)q2;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(yrow + ((un ^ swz) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
            fence_proxy_async();                                // generic-proxy writes -> visible to tcgen05.mma
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_ae + 8u * ab); mbar_arrive(bar_yf + 8u * blk); }
          } else {
            // ---- epilogue 2: out = relu(acc + b2 + x) ---------------------------------------------------------------------------
            const int j = blk - n1;
            int pix;
            interior(q0 + 128 * j + quarter * 32 + lane, &pix);
            for (uint32_t idx = lane; idx < (32u << cpr_log); idx += 32) {
              const uint32_t row = idx >> cpr_log, ch = idx & (cpr - 1u);
              const int pr = __shfl_sync(0xffffffffu, pix, (int)row);
              if (pr >= 0) cp_async16(wst + row * pitch + (ch << 4), x8 + (size_t)pr * row_b + (ch << 4), 16u);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            mbar_wait(bar_af + 8u * ab, use & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < C; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(t_row + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                uint32_t w[4];
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my + (uint32_t)c0 * 2u + 16u * q2));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 bq;
                  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bq.x), "=f"(bq.y) : "r"(s_bias + 4u * (uint32_t)(C + c0 + q2 * 8 + k * 2)));
                  const float2 rr = unpack2<F16>(w[k]);
                  const float f0 = fmaxf(__uint_as_float(v[q2 * 8 + k * 2]) + bq.x + rr.x, 0.f);
                  const float f1 = fmaxf(__uint_as_float(v[q2 * 8 + k * 2 + 1]) + bq.y + rr.y, 0.f);
                  w[k] = pack2<F16>(f0, f1);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)c0 * 2u + 16u * q2), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ae + 8u * ab);
            for (uint32_t idx = lane; idx < (32u << cpr_log); idx += 32) {
              const uint32_t row = idx >> cpr_log, ch = idx & (cpr - 1u);
              const int pr = __shfl_sync(0xffffffffu, pix, (int)row);
              if (pr >= 0) {
                uint4 t;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(wst + row * pitch + (ch << 4)));
                *reinterpret_cast<uint4*>(out8 + (size_t)pr * row_b + (ch << 4)) = t;
              }
            }
            __syncwarp();
          }
        }
        if (++gsel == G) gsel = 0;
        if (++ab == NACC) { ab = 0; ++use; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn block_encode_tiled() {
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

int block_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool block_plan(const BlockArgs& a, BlockParams* p, size_t* smem) {
  static const int off = block_env("HRP_NO_BLOCK_FUSION", 0), force_n = block_env("HRP_BLOCK_N", 0), force_g = block_env("HRP_BLOCK_GROUPS", 0);
  if (off || a.C != 32 || a.W + 2 > 256 || a.W < 8 || a.H < 1 || a.B < 1) return false;     // bf16, 64-byte operand rows
  if (block_encode_tiled() == nullptr) return false;
  constexpr int ROWB = 64, ESZ = 2;
  {
    ConvArgs g{};                                            // the weight images are the ones conv_tc / conv_slab would use
    g.B = 1; g.Hi = g.Ho = a.H; g.Wi = g.Wo = a.W; g.Cin = g.Cout = a.C; g.KH = g.KW = 3; g.stride = 1; g.pad_h = g.pad_w = 1; g.ld_out = a.C;
    if (conv_tc_row_bytes(g, 0, nullptr) != ROWB) return false;
  }
  p->a = a;
  p->C = a.C;
  p->Wp = a.W + 2; p->Hp = a.H + 2; p->HpWp = p->Wp * p->Hp;
  p->Q = (long long)a.B * p->HpWp;
  if (p->Q > 0x7fff0000LL) return false;
  p->w_bytes = 9 * a.C * ROWB;
  int cl = 0;
  while ((16 << cl) < a.C * ESZ) ++cl;
  p->cpr_log = cl;
  p->kmax = (127 + 2 * p->Wp + 2) / 128;
  const size_t warp_stg = (size_t)32 * (a.C * ESZ + 16), tail = 512;
  const int slots = std::max(1, sm_count() * (a.grid_pct > 0 ? a.grid_pct : 100) / 100);
  auto geom = [&](int n, int* n1, size_t* xb, size_t* yb) {
    *n1 = (128 * n + 2 * p->Wp + 2 + 127) / 128;
    const int xrows = (128 * *n1 + 2 * p->Wp + 2 + p->Wp - 1) / p->Wp + 1;
    *xb = ((size_t)xrows * p->Wp * ROWB + 1023) / 1024 * 1024;
    *yb = ((size_t)(128 * *n1) * ROWB + 1023) / 1024 * 1024;
  };
  int best_n = 0, best_g = 0;
  double best_cost = 0.0;
  for (int G = BK_MAX_GROUPS; G >= 1 && best_n == 0; --G) {
    if (force_g && G != force_g) continue;
    for (int n = 2; n <= 10; ++n) {
      if (force_n && n != force_n) continue;
      int n1; size_t xb, yb;
      geom(n, &n1, &xb, &yb);
      if (n1 > BK_MAX_N1) continue;
      const size_t need = 1024 + xb + yb + 2 * (size_t)p->w_bytes + 4 * G * warp_stg + tail;
      if (need > (size_t)BK_SMEM_LIMIT) continue;
      const long long units = (p->Q + 128LL * n - 1) / (128LL * n);
      const long long rounds = (units + slots - 1) / slots;
      const double cost = (double)rounds * (n1 + n + 2.0);      // blocks of work per unit + a fixed per-unit allowance
      if (best_n == 0 || cost < best_cost) { best_n = n; best_g = G; best_cost = cost; }
    }
  }
  if (best_n == 0) return false;
  p->n = best_n;
  size_t xb, yb;
  geom(best_n, &p->n1, &xb, &yb);
  p->groups = best_g;
  p->nacc = 2 * best_g;
  int tm = 32;
  while (tm < p->nacc * a.C) tm <<= 1;
  p->tmem_cols = tm;
  p->x_bytes = (int)xb;
  p->y_off = p->x_bytes;
  p->w_off = p->y_off + (int)yb;
  p->stg_off = p->w_off + 2 * p->w_bytes;
  p->bar_off = (p->stg_off + (int)(4 * best_g * warp_stg) + 127) / 128 * 128;
  p->units = (int)((p->Q + 128LL * best_n - 1) / (128LL * best_n));
  *smem = 1024 + (size_t)p->bar_off + tail;
  return *smem <= (size_t)BK_SMEM_LIMIT;
}

}  // namespace

bool conv_block_supported(const BlockArgs& a) {
  BlockParams p{};
  size_t smem = 0;
  return block_plan(a, &p, &smem);
}

int conv_block_launch(const BlockArgs& a, cudaStream_t st) {
  BlockParams p{};
  size_t smem = 0;
  if (!block_plan(a, &p, &smem)) return fail(HRP_ERR_INVALID, "conv_block: unsupported block (C=%d, %dx%d)", a.C, a.H, a.W);
  constexpr int ROWB = 64, ESZ = 2;
  const cuuint64_t gdim[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.C * ESZ, (cuuint64_t)a.W * a.C * ESZ, (cuuint64_t)a.H * a.W * a.C * ESZ};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const int box_rows[3] = {16, 4, 1};
  for (int m = 0; m < 3; ++m) {
    CUtensorMap tm;
    const cuuint32_t box[4] = {(cuuint32_t)(ROWB / ESZ), (cuuint32_t)p.Wp, (cuuint32_t)box_rows[m], 1};
    const CUresult r = block_encode_tiled()(&tm, a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.x), gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "conv_block: cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::memcpy(p.tmap[m], &tm, 128);
  }
  static bool attr_done = false;
  if (!attr_done) {
    HRP_CUDA(cudaFuncSetAttribute(conv_block_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_block_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_LIMIT));
    attr_done = true;
  }
  const int grid = std::min(p.units, std::max(1, sm_count() * (a.grid_pct > 0 ? a.grid_pct : 100) / 100));
  if (a.f16) conv_block_kernel<64, true><<<grid, 32 * BK_FIRST_EPI + 128 * p.groups, smem, st>>>(p);
  else conv_block_kernel<64, false><<<grid, 32 * BK_FIRST_EPI + 128 * p.groups, smem, st>>>(p);
  HRP_CHECK_LAUNCH("conv_block_kernel");
  return HRP_OK;
}

}  // namespace hrp
