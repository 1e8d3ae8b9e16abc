// 3x3 / stride 1 / pad 1 convolution with Cin == Cout (every BasicBlock conv of the HRNet branches, HRnet.py:28-57) as a
// "shifted GEMM" on tcgen05: the im2col matrix is never materialised, not even in shared memory.
//
// Why: these layers have 32..128 channels, so a 128-pixel tile needs only a few hundred MMA cycles, but the implicit-GEMM
// kernel (conv_tc.cu) re-reads every input pixel nine times from L2 (one TMA box per tap) and is bound by that traffic
// (profiles/r01_ncu_conv_tc_32ch_3x3_tma.txt). Here a CTA loads a slab of zero-padded input rows ONCE and issues the nine
// taps as nine MMAs whose A descriptors start (r-1)*(W+2) + (s-1) rows further down the same slab:
//
//   padded position q = b*(H+2)*(W+2) + yy*(W+2) + xx   <->  input/output pixel (yy-1, xx-1) of frame b
//   out(q) = sum_{r,s} in_padded(q + (r-1)*(W+2) + (s-1)) . w[r][s]            (valid for 1 <= yy <= H, 1 <= xx <= W)
//
// M rows of the GEMM are consecutive padded positions (border positions compute garbage that is never stored: 6 % of the
// MMA work at 64x64, irrelevant next to the 9x cut in operand traffic). The slab rows are K-major operand rows
// (SWIZZLE_64B when Cin*elem == 64 bytes, else SWIZZLE_128B with one plane per 128 bytes of channels) written by TMA,
// one {channels, W+2, 1, 1} box per padded row with the halo columns / rows zero-filled by the TMA unit; tcgen05.mma
// reads a swizzled K-major operand correctly from any row offset because the swizzle is a function of the absolute
// shared-memory address (scripts/experiments/umma_shift_test.cu). Weights (all nine taps, the pack_conv_tc image) stay
// resident in shared memory for the life of the persistent CTA.
//
// Warp roles: warp 0 loader (TMA), warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (TMEM -> +bias, +residual, ReLU ->
// staging tile -> 16-byte coalesced stores). Slabs and TMEM accumulators are double-buffered.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_ptx.h"

namespace hrp {
namespace {
using namespace tc;

constexpr int SL_MAX_GROUPS = 3;
constexpr int SL_SMEM_LIMIT = 227 * 1024;

struct SlabParams {
  alignas(64) unsigned char tmap[3][128];   // NHWC input as {C, W, H, B}; boxes {row_bytes/esz, W+2, R, 1} for R = 16, 4, 1
  ConvArgs a;
  int C, row_bytes, planes, kb_elems;
  int Wp, Hp, HpWp;
  long long Q;            // padded positions in the batch
  int nblk, units;
  int slab_rows;          // padded rows one slab buffer holds
  int plane_bytes;        // slab_rows * Wp * row_bytes, 1024-aligned
  int slab_bytes;         // planes * plane_bytes
  int w_bytes, w_off, stg_off, bar_off;
  int groups;             // epilogue groups of four warps; blocks go round-robin over groups
  int nacc;               // TMEM accumulator buffers (2 per group)
  int pdl_early;     // HRP_PDL_EARLY: let the next kernel start launching right after this one's prologue
  int tmem_cols, cpr_log, round_tf32;
};

// barrier block layout (8-byte slots): w | slab_full[2] | slab_empty[2] | acc_full[6] | acc_empty[6] | tmem slot
template <bool TF32, int ROWB, int PLANES, bool F16>
__global__ void __launch_bounds__(64 + 128 * SL_MAX_GROUPS, 1)
conv_slab_kernel(const __grid_constant__ SlabParams p) {
  constexpr int ESZ = TF32 ? 4 : 2;
  constexpr int KSTEPS = ROWB / 32;          // tcgen05.mma per operand row (K = 32 bytes each)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sSlab = base, sW = base + (uint32_t)p.w_off, stg = base + (uint32_t)p.stg_off, sBar = base + (uint32_t)p.bar_off;
  const uint32_t bar_w = sBar, bar_sf = sBar + 8u, bar_se = sBar + 24u, bar_af = sBar + 40u, bar_ae = sBar + 88u, tmem_slot = sBar + 136u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ConvArgs& a = p.a;
  const int C = p.C, Wp = p.Wp, G = p.groups, NACC = p.nacc;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_sf + 8u * i, 1); mbar_init(bar_se + 8u * i, 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(bar_af + 8u * i, 1); mbar_init(bar_ae + 8u * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane < 3) asm volatile("prefetch.tensormap [%0];" ::"l"(p.tmap[lane]) : "memory");
  // the bias vector goes to shared memory once per CTA: a global load per four columns inside the drain loop keeps the
  // in-order epilogue warps on the long scoreboard
  const uint32_t s_bias = sBar + 256u;
  if (tid >= 64 && tid - 64 < C / 4) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias) + (tid - 64));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(s_bias + 16u * (uint32_t)(tid - 64)), "f"(b4.x), "f"(b4.y), "f"(b4.z), "f"(b4.w) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (p.pdl_early) pdl_trigger();

  const int unit_pos = 128 * p.nblk;
  if (warp == 0) {
    // ===== loader: all lanes walk the loops, one elected lane issues (uniform TMA operands, see tc_ptx.h: elect_one) =====
    const bool leader = elect_one();
    if (leader) mbar_arrive_expect_tx(bar_w, (uint32_t)p.w_bytes);
    const int tap_bytes = p.w_bytes / 9;
    for (int t = 0; t < 9; ++t)          // weights are constants: staged before the previous layer has finished
      if (leader) bulk_g2s(sW + (uint32_t)(t * tap_bytes), static_cast<const uint8_t*>(a.w) + (size_t)t * tap_bytes, (uint32_t)tap_bytes, bar_w);
    pdl_wait();
    const int total_rows = a.B * p.Hp;
    const uint32_t row_tx = (uint32_t)(Wp * p.row_bytes);
    int li = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++li) {
      const int buf = li & 1;
      if (li >= 2) mbar_wait(bar_se + 8u * buf, ((li >> 1) & 1) ^ 1);
      const long long q0 = (long long)u * unit_pos;
      const long long lo = q0 - Wp - 1, hi = q0 + unit_pos + Wp + 1;
      const int ra = (int)(lo >= 0 ? lo / Wp : -((-lo + Wp - 1) / Wp));      // floor
      const int rb = (int)((hi + Wp - 1) / Wp);
      const int r0 = ra < 0 ? 0 : ra, r1 = rb > total_rows ? total_rows : rb;
      const uint32_t dst0 = sSlab + (uint32_t)(buf * p.slab_bytes);
      const uint32_t bar = bar_sf + 8u * buf;
      if (leader) mbar_arrive_expect_tx(bar, (uint32_t)((r1 - r0) * p.planes) * row_tx);
      // the slab's padded rows, frame by frame, as few TMA boxes as possible (16-, 4- and 1-row boxes): the TMA unit works
      // through the boxes of one SM nearly serially, so 19 one-row boxes cost far more than the bytes they move
      int R = r0;
      while (R < r1) {
        const int b = R / p.Hp, yy = R - b * p.Hp;
        const int seg = min(r1 - R, p.Hp - yy);              // rows left in this frame
        int done = 0;
        while (done < seg) {
          const int left = seg - done;
          const int m = left >= 16 ? 0 : (left >= 4 ? 1 : 2), rows = left >= 16 ? 16 : (left >= 4 ? 4 : 1);
          const uint32_t d = dst0 + (uint32_t)(R + done - ra) * row_tx;
          for (int h = 0; h < p.planes; ++h)
            if (leader) tma_load_4d(d + (uint32_t)(h * p.plane_bytes), p.tmap[m], h * p.kb_elems, -1, yy + done - 1, b, bar);
          done += rows;
        }
        R += seg;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: all lanes walk the loops (uniform operands), one elected lane issues. The per-block body is
    // 9 x PLANES x KSTEPS MMAs with compile-time trip counts; every operand is a 32-bit descriptor low word = base +
    // precomputed offset, so the issue stream stays a few instructions per MMA (a small-N MMA is only ~44 clk). ==========
    const bool leader = elect_one();
    constexpr uint32_t FMT = TF32 ? 2u : (F16 ? 0u : 1u);       // operand format: TF32 / IEEE half / bf16
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dhi = umma_desc_hi(ROWB);
    const uint32_t wt16 = (uint32_t)(C * ROWB) >> 4, plane16 = (uint32_t)p.plane_bytes >> 4, w16 = (sW >> 4) | (1u << 16);
    uint32_t tap16[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tap16[t] = (uint32_t)(((t / 3) * Wp + (t % 3)) * (ROWB / 16));
    mbar_wait(bar_w, 0);
    int li = 0, ab = 0, use = 0;                                  // ab = block % NACC, use = block / NACC
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++li) {
      const int buf = li & 1;
      mbar_wait(bar_sf + 8u * buf, (li >> 1) & 1);
      tc_fence_after();
      const long long q0 = (long long)u * unit_pos;
      const long long lo = q0 - Wp - 1;
      const long long ra = lo >= 0 ? lo / Wp : -((-lo + Wp - 1) / Wp);
      // low descriptor word of tap (0,0) of block 0: slab row (q0 - ra*Wp) - Wp - 1
      uint32_t a16 = (((sSlab + (uint32_t)(buf * p.slab_bytes)) >> 4) + (uint32_t)(q0 - ra * Wp - Wp - 1) * (ROWB / 16)) | (1u << 16);
      for (int j = 0; j < p.nblk; ++j, a16 += 128u * (ROWB / 16)) {
        if (use >= 1) mbar_wait(bar_ae + 8u * ab, (use - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ab * C);
        if (leader) {
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int h = 0; h < PLANES; ++h)
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk)
                umma_lo<TF32>(tmem_d, a16 + tap16[t] + (uint32_t)h * plane16 + 2u * kk, w16 + (uint32_t)(t * PLANES + h) * wt16 + 2u * kk, dhi, idesc,
                              (t | h | kk) != 0 ? 1u : 0u);
          umma_commit(bar_af + 8u * ab);
          if (j == p.nblk - 1) umma_commit(bar_se + 8u * buf);                 // slab free once these MMAs have read it
        }
        __syncwarp();
        if (++ab == NACC) { ab = 0; ++use; }
      }
    }
  } else if (warp < 2 + 4 * G) {
    // ===== epilogue: every warp is its own pipeline over the 32 tile rows of its TMEM lane quarter =========================
    // group g takes blocks gb % G == g; no barrier wider than a warp: residual rows arrive by cp.async into the warp's
    // staging tile, are combined in place with the accumulator, and leave as 16-byte chunks, 512 contiguous bytes per request
    pdl_wait();
    const int e = warp - 2, g = e >> 2, quarter = warp & 3;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t pitch = (uint32_t)C * ESZ + 16u;
    const uint32_t cpr_log = (uint32_t)p.cpr_log, cpr = 1u << cpr_log;
    const uint32_t wst = stg + (uint32_t)e * 32u * pitch;            // this warp's 32-row staging tile
    const uint32_t my = wst + (uint32_t)lane * pitch;
    const bool has_res = a.res != nullptr;
    const uint8_t* res8 = static_cast<const uint8_t*>(a.res);
    uint8_t* out8 = static_cast<uint8_t*>(a.out);
    const uint32_t Qu = (uint32_t)p.Q, HpWp = (uint32_t)p.HpWp, Wpu = (uint32_t)Wp;
    const size_t row_b = (size_t)C * ESZ;
    int gb = 0, ab = 0, use = 0, gsel = 0;                        // gsel = gb % G
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const uint32_t q0 = (uint32_t)u * (uint32_t)unit_pos;
      for (int j = 0; j < p.nblk; ++j, ++gb) {
        if (gsel == g) {
          const uint32_t q = q0 + 128u * (uint32_t)j + (uint32_t)(quarter * 32 + lane);
          const uint32_t b = q / HpWp, rem = q - b * HpWp, yy = rem / Wpu, xx = rem - yy * Wpu;
          const bool ok = q < Qu && yy >= 1u && yy <= (uint32_t)a.Hi && xx >= 1u && xx <= (uint32_t)a.Wi;
          const int pix = ok ? (int)((b * (uint32_t)a.Hi + (yy - 1u)) * (uint32_t)a.Wi + (xx - 1u)) : -1;
          if (has_res) {
            for (uint32_t idx = lane; idx < (32u << cpr_log); idx += 32) {
              const uint32_t row = idx >> cpr_log, ch = idx & (cpr - 1u);
              const int pr = __shfl_sync(0xffffffffu, pix, (int)row);
              if (pr >= 0) cp_async16(wst + row * pitch + (ch << 4), res8 + (size_t)pr * row_b + (ch << 4), 16u);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
          }
          mbar_wait(bar_af + 8u * ab, use & 1);
          tc_fence_after();
          const uint32_t t_row = t_lane + (uint32_t)(ab * C);
          for (int c0 = 0; c0 < C; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_row + (uint32_t)c0, v);
            tmem_ld_wait();
            float f[16], r[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float4 bq;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bq.x), "=f"(bq.y), "=f"(bq.z), "=f"(bq.w) : "r"(s_bias + 4u * (uint32_t)c0 + 16u * (uint32_t)q4));
              f[q4 * 4 + 0] = __uint_as_float(v[q4 * 4 + 0]) + bq.x;
              f[q4 * 4 + 1] = __uint_as_float(v[q4 * 4 + 1]) + bq.y;
              f[q4 * 4 + 2] = __uint_as_float(v[q4 * 4 + 2]) + bq.z;
              f[q4 * 4 + 3] = __uint_as_float(v[q4 * 4 + 3]) + bq.w;
            }
            if (has_res) {
              if constexpr (TF32) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[q4 * 4]), "=f"(r[q4 * 4 + 1]), "=f"(r[q4 * 4 + 2]), "=f"(r[q4 * 4 + 3]) : "r"(my + (uint32_t)c0 * 4u + 16u * q4));
              } else {
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                  uint32_t w[4];
                  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my + (uint32_t)c0 * 2u + 16u * q2));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 ff = unpack2<F16>(w[k]);
                    r[q2 * 8 + k * 2] = ff.x; r[q2 * 8 + k * 2 + 1] = ff.y;
                  }
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float t = f[k];
              if (has_res && !a.res_after_act) t += r[k];
              if (a.relu) t = fmaxf(t, 0.f);
              if (has_res && a.res_after_act) t += r[k];
              f[k] = t;
            }
            if constexpr (TF32) {
              if (p.round_tf32) {
#pragma unroll
                for (int k = 0; k < 16; ++k) f[k] = round_tf32_rna(f[k]);
              }
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)c0 * 4u + 16u * q4), "f"(f[q4 * 4]), "f"(f[q4 * 4 + 1]), "f"(f[q4 * 4 + 2]), "f"(f[q4 * 4 + 3]) : "memory");
            } else {
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  w[k] = pack2<F16>(f[q2 * 8 + k * 2], f[q2 * 8 + k * 2 + 1]);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + (uint32_t)c0 * 2u + 16u * q2), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_ae + 8u * ab);             // this warp's quarter of the accumulator is drained
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
EncodeTiledFn slab_encode_tiled() {
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

int slab_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// geometry + shared-memory plan; returns false when the layer does not fit this kernel
bool slab_plan(const ConvArgs& a, int tf32, SlabParams* p, size_t* smem) {
  static const int off = slab_env("HRP_NO_SLAB", 0), force_nblk = slab_env("HRP_SLAB_NBLK", 0);
  if (off) return false;
  const int esz = tf32 ? 4 : 2;
  if (a.KH != 3 || a.KW != 3 || a.stride != 1 || a.pad_h != 1 || a.pad_w != 1 || a.Cin != a.Cout || a.out_nchw) return false;
  if (a.Ho != a.Hi || a.Wo != a.Wi || a.out_sy != 1 || a.out_sx != 1 || a.Ho_full != a.Ho || a.Wo_full != a.Wo) return false;
  if (a.ld_out != a.Cout || a.out_coff != 0) return false;
  const int cb = a.Cin * esz;
  if (!(cb == 64 || cb == 128 || cb == 256) || (tf32 && cb == 64)) return false;
  if (a.Cout % 16 || a.Cout > 256 || a.Wi + 2 > 256 || a.Wi < 8) return false;
  if (slab_encode_tiled() == nullptr) return false;
  if (conv_tc_row_bytes(a, tf32, nullptr) != (cb == 64 ? 64 : 128)) return false;   // the weight image is shared with conv_tc
  p->a = a;
  p->C = a.Cin;
  p->row_bytes = cb == 64 ? 64 : 128;
  p->planes = cb / p->row_bytes;
  p->kb_elems = p->row_bytes / esz;
  p->Wp = a.Wi + 2; p->Hp = a.Hi + 2; p->HpWp = p->Wp * p->Hp;
  p->Q = (long long)a.B * p->HpWp;
  p->w_bytes = 9 * p->planes * a.Cout * p->row_bytes;
  if (p->w_bytes > 96 * 1024) return false;                 // resident weights only (streamed variant: not yet)
  int cl = 0;
  while ((16 << cl) < a.Cout * esz) ++cl;
  p->cpr_log = cl;
  const int sms = sm_count();
  static const int force_groups = slab_env("HRP_SLAB_GROUPS", 0);
  const size_t warp_stg = (size_t)32 * (a.Cout * esz + 16), tail = 1024;   // barriers (256 B) + the bias vector (<= 128 floats)
  const size_t wres = (size_t)((p->w_bytes + 1023) / 1024 * 1024);
  // Plan: epilogue groups G (each four warps with a private staging tile and two TMEM accumulators) and unit size n
  // (128-position blocks per slab). The epilogue is the long pole (residual fetch + store latency per block), so take
  // the most groups that leave room for a useful slab; then the unit size by a small cost model: rounds of units per
  // CTA x bytes one unit moves through L2 (slab in, tile out, residual in, plus a fixed per-unit latency allowance).
  int best = 0, best_g = 0;
  double best_cost = 0.0;
  for (int G = SL_MAX_GROUPS; G >= 1 && best == 0; --G) {
    if (force_groups && G != force_groups) continue;
    if (2 * G * a.Cout > 512) continue;
    for (int n = 1; n <= 8; ++n) {
      const int rows = (128 * n + 2 * p->Wp + 2 + p->Wp - 1) / p->Wp + 1;
      const size_t plane = ((size_t)rows * p->Wp * p->row_bytes + 1023) / 1024 * 1024;
      const size_t need = 1024 + 2 * plane * p->planes + wres + 4 * G * warp_stg + tail;
      if (need > (size_t)SL_SMEM_LIMIT) continue;
      if (n < 2 && G > 1) continue;                           // a slab this small re-reads more halo than it saves
      const long long units = (p->Q + 128LL * n - 1) / (128LL * n);
      const long long rounds = (units + sms - 1) / sms;
      const double unit_bytes = (double)rows * p->Wp * p->row_bytes * p->planes + 128.0 * n * a.Cout * esz * (a.res ? 2.0 : 1.0) + 16384.0;
      const double cost = (double)rounds * unit_bytes;
      if (best == 0 || cost < best_cost) { best = n; best_cost = cost; best_g = G; }
    }
  }
  if (best == 0) return false;
  p->groups = best_g;
  p->nacc = 2 * best_g;
  int tm = 32;
  while (tm < p->nacc * a.Cout) tm <<= 1;
  p->tmem_cols = tm;
  if (force_nblk >= 1 && force_nblk <= 8) best = force_nblk;
  p->nblk = best;
  p->slab_rows = (128 * best + 2 * p->Wp + 2 + p->Wp - 1) / p->Wp + 1;
  p->plane_bytes = (int)(((size_t)p->slab_rows * p->Wp * p->row_bytes + 1023) / 1024 * 1024);
  p->slab_bytes = p->plane_bytes * p->planes;
  p->units = (int)((p->Q + 128LL * best - 1) / (128LL * best));
  p->w_off = 2 * p->slab_bytes;
  p->stg_off = p->w_off + (int)wres;
  p->bar_off = p->stg_off + (int)(4 * best_g * warp_stg + 127) / 128 * 128;
  *smem = 1024 + (size_t)p->bar_off + tail;
  return *smem <= (size_t)SL_SMEM_LIMIT;
}

}  // namespace

bool conv_slab_supported(const ConvArgs& a, int tf32) {
  SlabParams p{};
  size_t smem = 0;
  return slab_plan(a, tf32, &p, &smem);
}

int conv_slab_launch(const ConvArgs& a, int tf32, int round_tf32, cudaStream_t st) {
  SlabParams p{};
  size_t smem = 0;
  if (!slab_plan(a, tf32, &p, &smem)) return fail(HRP_ERR_INVALID, "conv_slab: unsupported layer");
  if (p.Q <= 0) return HRP_OK;
  p.round_tf32 = round_tf32;
  const int esz = tf32 ? 4 : 2;
  const cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.Cin * esz, (cuuint64_t)a.Wi * a.Cin * esz, (cuuint64_t)a.Hi * a.Wi * a.Cin * esz};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const int box_rows[3] = {16, 4, 1};
  for (int m = 0; m < 3; ++m) {
    CUtensorMap tm;
    const cuuint32_t box[4] = {(cuuint32_t)p.kb_elems, (cuuint32_t)p.Wp, (cuuint32_t)box_rows[m], 1};
    const CUresult r = slab_encode_tiled()(&tm, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 4, const_cast<void*>(a.in),
                                           gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "conv_slab: cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::memcpy(p.tmap[m], &tm, 128);
  }
  static bool attr_done = false;
  if (!attr_done) {
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<true, 128, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<true, 128, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 64, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 128, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 128, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 64, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, 128, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_LIMIT));
    attr_done = true;
  }
  const int grid = std::min(p.units, std::max(1, sm_count() * (a.grid_pct > 0 ? a.grid_pct : 100) / 100));
  const int threads = 64 + 128 * p.groups;
  static const int pdl_early = slab_env("HRP_PDL_EARLY", 0);
  p.pdl_early = pdl_early;
  cudaError_t le;
  if (tf32 && p.planes == 1) le = launch_pdl(conv_slab_kernel<true, 128, 1, false>, grid, threads, smem, st, p);
  else if (tf32) le = launch_pdl(conv_slab_kernel<true, 128, 2, false>, grid, threads, smem, st, p);
  else if (a.f16) {
    if (p.row_bytes == 64) le = launch_pdl(conv_slab_kernel<false, 64, 1, true>, grid, threads, smem, st, p);
    else if (p.planes == 1) le = launch_pdl(conv_slab_kernel<false, 128, 1, true>, grid, threads, smem, st, p);
    else le = launch_pdl(conv_slab_kernel<false, 128, 2, true>, grid, threads, smem, st, p);
  }
  else if (p.row_bytes == 64) le = launch_pdl(conv_slab_kernel<false, 64, 1, false>, grid, threads, smem, st, p);
  else if (p.planes == 1) le = launch_pdl(conv_slab_kernel<false, 128, 1, false>, grid, threads, smem, st, p);
  else le = launch_pdl(conv_slab_kernel<false, 128, 2, false>, grid, threads, smem, st, p);
  if (le != cudaSuccess) return fail(HRP_ERR_CUDA, "conv_slab_kernel launch: %s", cudaGetErrorString(le));
  HRP_CHECK_LAUNCH("conv_slab_kernel");
  return HRP_OK;
}

}  // namespace hrp
