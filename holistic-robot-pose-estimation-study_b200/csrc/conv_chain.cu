// A whole branch of an HRNet module -- up to four BasicBlocks, relu(bn2(conv2(relu(bn1(conv1(x))))) + x) each, all 3x3/s1/p1
// with Cin == Cout (HRnet.py:28-57, 137-149) -- as ONE kernel, for the low-resolution branches (128 ch @ 16x16, 256 ch @ 8x8).
//
// Those layers are tiny (64 images x 256 or 64 pixels) and as separate launches each of them pays launch, pipeline fill,
// a round trip of the activations through L2 and a tail for ~2-5 us of tensor-core work. Here one CTA owns one IMAGE for
// the whole chain: its zero-padded activation ((H+2) x (W+2) positions x C channels, bf16) lives in shared memory in the
// UMMA operand layout (one [positions][128 B] SWIZZLE_128B matrix per 64-channel plane), every conv is a shifted GEMM
// over padded-position space (conv_slab.cu: the tap (r, s) is the same matrix read r*(W+2)+s rows further down), and
// only the weights stream -- from L2, where all CTAs read the same bytes -- through a ring of bulk copies:
//   loader   : TMA box {64 ch, W+2, H+2} per plane of image b -> buffer X (halo zero-filled by the TMA unit), then the
//              weight k-blocks (one tap x 64 input channels x Cout rows) of conv 0, 1, ... into the ring
//   MMA      : conv j reads X (j even) or Y (j odd): per k-block, MT tiles x 4 MMAs (M128, N = C, K16) into MT accumulators
//   epilogue : 8 warps (4 TMEM lane quarters x 2 column halves): +bias, ReLU -> bf16 -> the OTHER buffer in operand
//              swizzle (conv1: Y; conv2: + residual read from X, result in place into X); interior positions only, so
//              the halo stays zero for the next conv
//   copy-out : after the last conv, the interior of buffer X -> out, 16-byte units, by the epilogue warps (a 4-D TMA tensor
//              store of the padded box with clipped halo traps with `illegal instruction` on this driver; not pursued)
// MMA and epilogue of one image alternate (conv j+1 needs all of conv j), the weight ring keeps running underneath.
// Results are bit-identical to the layer-by-layer path (same operands, same K order, same bf16 roundings).
// Measured (B200, batch 64, alone): 105 us per 8-conv chain for both shapes, against 8 x 12.8 us (128 ch) and 8 x 17.3 us
// (256 ch) layer by layer -- on 64 SMs instead of all of them, which is what the multi-lane graph needs: +1.8 % frames/s
// and -9 % latency at batch 16 on the full network. ncu (profiles/r01_ncu_conv_chain_128ch.txt): the tensor pipe is
// active 58 % of the time; the N = 128 MMAs read 8 KB of shared memory per 64-clock slot, i.e. run at the 128 B/clk
// shared-memory limit and lose ~30 % to the weight ring's writes and the unaligned row shifts; the serial epilogue
// costs ~15 %. The ring depth does not matter (2 stages = 3 stages), nor does rotating the tap order per CTA to
// spread the L2 reads. Below 8 images the executor runs the same convs layer by layer (network.cu, OP_CHAIN).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_ptx.h"

namespace hrp {
namespace {
using namespace tc;

constexpr int CH_EPI_WARPS = 8;
// MMA issuer warps: one elected thread needs ~100 clk of issue slots per tcgen05.mma (descriptor arithmetic, R2UR moves),
// more than the 64 clk an M128 N128 K16 MMA occupies the tensor pipe (conv_roll.cu measured this). The MT tiles of an
// image have separate accumulators, so issuer w takes tiles w, w + 3, ... and the instruction streams interleave.
constexpr int CH_ISSUERS = 3;
constexpr int CH_THREADS = 32 * (1 + CH_ISSUERS + CH_EPI_WARPS);
constexpr int CH_MAX_STAGES = 6;
constexpr int CH_SMEM_LIMIT = 227 * 1024;

struct ChainParams {
  alignas(64) unsigned char tmap_in[128];    // NHWC x   as {C, W, H, B}, box {64, W+2, H+2, 1}, SWIZZLE_128B
  ChainArgs a;
  int Wp, Hp, P, MT, margin;
  int plane_bytes, y_off, w_off, bar_off, stages, kb_bytes, tmem_cols;
};

// barrier block (8-byte slots): x_full | x_free | acc_full | epi_done | w_full[6] | w_empty[6] | tmem slot
template <int PLANES, bool F16>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainParams p) {
  constexpr int C = 64 * PLANES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base, sY = base + (uint32_t)p.y_off, sW = base + (uint32_t)p.w_off, sBar = base + (uint32_t)p.bar_off;
  const uint32_t bar_xfull = sBar, bar_xfree = sBar + 8u, bar_acc = sBar + 16u, bar_epi = sBar + 24u, bar_wf = sBar + 32u,
                 bar_we = bar_wf + 8u * CH_MAX_STAGES, tmem_slot = bar_we + 8u * CH_MAX_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainArgs& a = p.a;
  const int Wp = p.Wp, MT = p.MT, S = p.stages, nconv = a.nconv;
  const uint32_t plane_bytes = (uint32_t)p.plane_bytes, kb_bytes = (uint32_t)p.kb_bytes, margin_b = (uint32_t)p.margin * 128u;

  if (tid == 0) {
    mbar_init(bar_xfull, 1); mbar_init(bar_xfree, 1); mbar_init(bar_acc, CH_ISSUERS); mbar_init(bar_epi, CH_EPI_WARPS);
    for (int i = 0; i < S; ++i) { mbar_init(bar_wf + 8u * i, 1); mbar_init(bar_we + 8u * i, CH_ISSUERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(p.tmap_in) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // zero the leading margin of every X plane (the taps of the first image row reach into it) and all of Y (its halo and
  // margin are never written afterwards: the epilogues store interior positions only)
  for (uint32_t i = (uint32_t)tid * 16u; i < (uint32_t)PLANES * margin_b; i += blockDim.x * 16u) {
    const uint32_t pl = i / margin_b, o = i - pl * margin_b;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sX + pl * plane_bytes + o), "r"(0u) : "memory");
  }
  for (uint32_t i = (uint32_t)tid * 16u; i < (uint32_t)PLANES * plane_bytes; i += blockDim.x * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sY + i), "r"(0u) : "memory");
  // every conv's bias vector goes to shared memory once per CTA: a global load per channel pair inside the drain loop keeps
  // the in-order epilogue warps on the long scoreboard (measured on conv_roll.cu)
  const uint32_t s_bias = sBar + 256u;
  for (int i = tid; i < nconv * C; i += blockDim.x) {
    const float bvv = __ldg(a.b[i / C] + (i % C));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias + 4u * (uint32_t)i), "f"(bvv) : "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== loader ========================================================================================================
    const bool leader = elect_one();
    const uint32_t x_tx = (uint32_t)PLANES * (uint32_t)p.P * 128u;
    int s = 0, use = 0, li = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++li) {
      if (li >= 1) mbar_wait(bar_xfree, (li - 1) & 1);            // the previous image has left buffer X
      if (leader) {
        mbar_arrive_expect_tx(bar_xfull, x_tx);
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) tma_load_4d(sX + (uint32_t)pl * plane_bytes + margin_b, p.tmap_in, pl * 64, -1, -1, b, bar_xfull);
      }
      for (int j = 0; j < nconv; ++j) {
        const uint8_t* wj = static_cast<const uint8_t*>(a.w[j]);
        for (int kb = 0; kb < 9 * PLANES; ++kb) {
          if (use >= 1) mbar_wait(bar_we + 8u * s, (use - 1) & 1);
          if (leader) {
            mbar_arrive_expect_tx(bar_wf + 8u * s, kb_bytes);
            bulk_g2s(sW + (uint32_t)s * kb_bytes, wj + (size_t)kb * kb_bytes, kb_bytes, bar_wf + 8u * s);
          }
          __syncwarp();
          if (++s == S) { s = 0; ++use; }
        }
      }
    }
  } else if (warp <= CH_ISSUERS) {
    // ===== MMA issuers (all lanes walk the loops, one elected lane issues; tc_ptx.h): issuer iw owns tiles iw, iw + 3, ... =====
    const bool leader = elect_one();
    const int iw = warp - 1;
    constexpr uint32_t FMT = F16 ? 0u : 1u;                     // operand format: IEEE half / bf16
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dhi = umma_desc_hi(128);
    int s = 0, use = 0, li = 0, nc = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++li) {
      mbar_wait(bar_xfull, li & 1);
      for (int j = 0; j < nconv; ++j, ++nc) {
        if (j >= 1) mbar_wait(bar_epi, (nc - 1) & 1);              // the previous conv's output is in shared memory, TMEM is free
        tc_fence_after();
        const uint32_t a_base = (((j & 1) ? sY : sX) >> 4) | (1u << 16);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const uint32_t tap16 = (uint32_t)((t / 3) * Wp + (t % 3) + p.margin - Wp - 1) * 8u;
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            mbar_wait(bar_wf + 8u * s, use & 1);
            tc_fence_after();
            const uint32_t a16 = a_base + (uint32_t)pl * (plane_bytes >> 4) + tap16;
            const uint32_t b16 = ((sW + (uint32_t)s * kb_bytes) >> 4) | (1u << 16);
            for (int m = iw; m < MT; m += CH_ISSUERS) {
              const uint32_t tm = tmem_base + (uint32_t)(m * C), am = a16 + (uint32_t)m * 1024u;
              if (leader) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_lo<false>(tm, am + 2u * kk, b16 + 2u * kk, dhi, idesc, (t | pl | kk) != 0 ? 1u : 0u);
              }
              __syncwarp();
            }
            if (leader) umma_commit(bar_we + 8u * s);
            __syncwarp();
            if (++s == S) { s = 0; ++use; }
          }
        }
        if (leader) umma_commit(bar_acc);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warp e owns TMEM lane quarter (warp & 3) and column half (e >> 2) of every tile =======================
    const int e = warp - 1 - CH_ISSUERS, quarter = warp & 3, half = e >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int c_lo = half * (C / 2), c_hi = c_lo + C / 2;
    int nc = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
      for (int j = 0; j < nconv; ++j, ++nc) {
        const bool second = (j & 1) != 0;
        const uint32_t bias_j = s_bias + (uint32_t)(j * C) * 4u;     // conv j's bias vector (shared memory, filled at kernel start)
        const uint32_t dst = second ? sX : sY;
        mbar_wait(bar_acc, nc & 1);
        tc_fence_after();
        for (int m = 0; m < MT; ++m) {
          const int q = m * 128 + quarter * 32 + lane;
          const int yy = q / Wp, xx = q - yy * Wp;
          const bool ok = q < p.P && yy >= 1 && yy <= a.H && xx >= 1 && xx <= a.W;
          if (!__any_sync(0xffffffffu, ok)) continue;
          const uint32_t R = (uint32_t)(q + p.margin);
          const uint32_t row = dst + R * 128u, swz = R & 7u;
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            uint32_t v[2][16];
            tmem_ld16(t_lane + (uint32_t)(m * C + c0), v[0]);
            tmem_ld16(t_lane + (uint32_t)(m * C + c0 + 16), v[1]);
            tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int cc = c0 + 16 * hh;
                const uint32_t prow = row + (uint32_t)(cc >> 6) * plane_bytes;
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                  const uint32_t un = (uint32_t)((cc & 63) >> 3) + (uint32_t)q2;
                  const uint32_t addr = prow + ((un ^ swz) << 4);
                  uint32_t w[4];
                  if (second) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    float2 bq;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bq.x), "=f"(bq.y) : "r"(bias_j + 4u * (uint32_t)(cc + q2 * 8 + k * 2)));
                    float f0 = __uint_as_float(v[hh][q2 * 8 + k * 2]) + bq.x, f1 = __uint_as_float(v[hh][q2 * 8 + k * 2 + 1]) + bq.y;
                    if (second) {
                      const float2 rr = unpack2<F16>(w[k]);
                      f0 += rr.x; f1 += rr.y;
                    }
                    w[k] = pack2<F16>(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                }
              }
            }
          }
        }
        fence_proxy_async();                                      // generic-proxy writes -> visible to tcgen05.mma
        tc_fence_before();
        __syncwarp();
        if (j + 1 < nconv) {
          if (lane == 0) mbar_arrive(bar_epi);
        } else {
          // last conv of the image: every epilogue warp is done with X -> one thread stores it and frees the buffer
          asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
          if (lane == 0) mbar_arrive(bar_epi);
          {
            // copy-out: 16-byte units of the interior pixels, un-swizzled, coalesced along each pixel's channels
            const int et = e * 32 + lane, upp = C / 8;
            uint8_t* o8 = static_cast<uint8_t*>(a.out) + (size_t)b * a.H * a.W * C * 2;
            for (int idx = et; idx < a.H * a.W * upp; idx += 32 * CH_EPI_WARPS) {
              const int pix = idx / upp, u = idx - pix * upp, y = pix / a.W, x = pix - y * a.W;
              const uint32_t R = (uint32_t)((y + 1) * Wp + (x + 1) + p.margin);
              uint4 t;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                           : "r"(sX + (uint32_t)(u >> 3) * plane_bytes + R * 128u + ((((uint32_t)u & 7u) ^ (R & 7u)) << 4)));
              *reinterpret_cast<uint4*>(o8 + (size_t)idx * 16) = t;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
            if (e == 0 && lane == 0) mbar_arrive(bar_xfree);
          }
        }
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
EncodeTiledFn chain_encode_tiled() {
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

int chain_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool chain_plan(const ChainArgs& a, ChainParams* p, size_t* smem) {
  static const int off = chain_env("HRP_NO_CHAIN_FUSION", 0), force_s = chain_env("HRP_CHAIN_STAGES", 0);
  if (off || (a.C != 128 && a.C != 256) || a.H < 1 || a.W < 1 || a.B < 1 || a.W + 2 > 256 || a.H + 2 > 256) return false;
  if (a.nconv < 2 || a.nconv > 8 || (a.nconv & 1)) return false;
  if (chain_encode_tiled() == nullptr) return false;
  {
    ConvArgs g{};                                            // the weight images are the ones conv_tc would use
    g.B = 1; g.Hi = g.Ho = a.H; g.Wi = g.Wo = a.W; g.Cin = g.Cout = a.C; g.KH = g.KW = 3; g.stride = 1; g.pad_h = g.pad_w = 1; g.ld_out = a.C;
    if (conv_tc_row_bytes(g, 0, nullptr) != 128) return false;
  }
  p->a = a;
  p->Wp = a.W + 2; p->Hp = a.H + 2; p->P = p->Wp * p->Hp;
  p->MT = (p->P + 127) / 128;
  if (p->MT * a.C > 512) return false;                       // MT accumulators of C columns in TMEM
  p->margin = (p->Wp + 1 + 7) / 8 * 8;                      // whole 8-row swizzle groups: the TMA boxes start 1024-byte aligned
  const int planes = a.C / 64;
  p->plane_bytes = ((p->P + p->margin) * 128 + 1023) / 1024 * 1024;
  p->y_off = planes * p->plane_bytes;
  p->w_off = 2 * p->y_off;
  p->kb_bytes = a.C * 128;
  // the taps of the last tile read up to 128*MT + 2*Wp + 2 rows into a plane; what lies behind the last Y plane must still
  // be inside the allocation (it only feeds accumulator rows that are never stored)
  const int overrun = std::max(0, (128 * p->MT + p->margin + p->Wp + 1) * 128 - p->plane_bytes);
  const size_t tail = 256 + (size_t)8 * a.C * 4;             // barriers + the bias vectors of up to eight convs
  int S = force_s ? force_s : CH_MAX_STAGES;
  S = std::min(S, CH_MAX_STAGES);
  while (S >= 2 && 1024 + (size_t)p->w_off + (size_t)S * p->kb_bytes + tail > (size_t)CH_SMEM_LIMIT) --S;
  if (S < 2 || (size_t)S * p->kb_bytes + tail < (size_t)overrun) return false;
  p->stages = S;
  p->bar_off = p->w_off + S * p->kb_bytes;
  int tm = 32;
  while (tm < p->MT * a.C) tm <<= 1;
  p->tmem_cols = tm;
  *smem = 1024 + (size_t)p->bar_off + tail;
  return *smem <= (size_t)CH_SMEM_LIMIT;
}

template <int PLANES, bool F16>
int chain_launch_t(const ChainParams& p, int grid, size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    HRP_CUDA(cudaFuncSetAttribute(conv_chain_kernel<PLANES, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_LIMIT));
    attr_done = true;
  }
  conv_chain_kernel<PLANES, F16><<<grid, CH_THREADS, smem, st>>>(p);
  HRP_CHECK_LAUNCH("conv_chain_kernel");
  return HRP_OK;
}

}  // namespace

bool conv_chain_supported(const ChainArgs& a) {
  ChainParams p{};
  size_t smem = 0;
  return chain_plan(a, &p, &smem);
}

int conv_chain_launch(const ChainArgs& a, cudaStream_t st) {
  ChainParams p{};
  size_t smem = 0;
  if (!chain_plan(a, &p, &smem)) return fail(HRP_ERR_INVALID, "conv_chain: unsupported chain (C=%d, %dx%d, %d convs)", a.C, a.H, a.W, a.nconv);
  const cuuint64_t gdim[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W * a.C * 2, (cuuint64_t)a.H * a.W * a.C * 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {64, (cuuint32_t)p.Wp, (cuuint32_t)p.Hp, 1};
  {
    CUtensorMap tm;
    const CUresult r = chain_encode_tiled()(&tm, a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.x), gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "conv_chain: cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::memcpy(p.tmap_in, &tm, 128);
  }
  // one image per CTA and one CTA per SM (the buffers fill the shared memory): no lane cap, the images ARE the grid
  const int grid = std::min(a.B, sm_count());
  if (a.f16) return a.C == 128 ? chain_launch_t<2, true>(p, grid, smem, st) : chain_launch_t<4, true>(p, grid, smem, st);
  return a.C == 128 ? chain_launch_t<2, false>(p, grid, smem, st) : chain_launch_t<4, false>(p, grid, smem, st);
}

}  // namespace hrp
