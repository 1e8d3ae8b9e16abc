// The 64-channel branch of an HRNet module -- four BasicBlocks, relu(bn2(conv2(relu(bn1(conv1(x))))) + x) each, all
// 3x3/s1/p1, 64 -> 64 channels on a 32x32 map (HRnet.py:28-57, 137-149) -- as ONE kernel, one CTA per image.
//
// conv_chain.cu does this for the 128 / 256-channel branches with two ping-pong activation buffers in shared memory; at
// 64 channels x 34x34 padded positions one buffer is 148 KB, so two do not fit. Here ONE buffer holds the activation and
// every conv writes its output IN PLACE, d = W+3 (>= W+2+1) rows further up: the shifted GEMM of conv_slab.cu reads, for
// output position q, input positions q - (W+3) .. q + (W+3), so by the time the tile that ends at q is stored d rows up,
// no later tile of the same conv needs what it overwrites. The image therefore drifts d rows per conv through a buffer
// that is nconv*d rows longer than the image (8 convs: 189 KB), and all convs run top-down in one uninterrupted MMA
// stream: conv j+1 starts on the first tiles of conv j's output while conv j's last tiles are still in flight.
//   loader   : TMA box {64 ch, W+2, H+2} of image b -> buffer rows [nconv*d, ...) (halo zero-filled by the TMA unit), then
//              the weight k-blocks (one tap: 64 Cout rows x 128 B) of conv 0, 1, ... -- once per PASS -- into a ring
//   MMA      : conv j in passes of TP tiles (128 positions each): per tap, TP x 4 MMAs (M128, N64, K16) into one of two
//              TMEM accumulator sets, issued by THREE warps (one tile each: a single issuing thread cannot feed N64 MMAs); pass g may start when the epilogues of pass g-2 (its accumulators) and of the
//              passes of conv j-1 that produced its input rows are done
//   epilogue : 16 warps (4 TMEM lane quarters x 4 column slices), each thread one position x 16 channels: +bias, ReLU
//              (conv2: + the block input, which this very thread wrote to global memory two convs earlier -- the block
//              outputs go to global memory anyway, the last one being the result), bf16, written in the operand swizzle
//              at the drifted position; halo positions are written as zeros (the next conv's padding)
// Bit-identical to the layer-by-layer path (same operands, same K order, same bf16 roundings).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_ptx.h"

namespace hrp {
namespace {
using namespace tc;

constexpr int RL_EPI_WARPS = 16;                // 4 TMEM lane quarters x 4 column quarters: 16 channels of one position per thread
constexpr int RL_CPT = 64 / (RL_EPI_WARPS / 4);  // channels per epilogue thread
// MMA issuer warps. One elected thread spends ~9 issue slots (descriptor arithmetic in uniform registers, R2UR moves,
// the MMA itself) of ~10 clk each per tcgen05.mma, i.e. ~100 clk per instruction, while an M128 N64 K16 MMA occupies the
// tensor pipe for 48 clk: a single issuer leaves the pipe half idle (measured: 135 us per 8-conv chain). The tiles of a
// pass have separate accumulators, so issuer w takes tiles w, w + 3, ... and the three instruction streams interleave.
constexpr int RL_ISSUERS = 3;
constexpr int RL_THREADS = 32 * (1 + RL_ISSUERS + RL_EPI_WARPS);
constexpr int RL_MAX_STAGES = 6;
constexpr int RL_SMEM_LIMIT = 227 * 1024;
constexpr int RL_C = 64;
constexpr int RL_TP_MAX = 3;                   // tiles per pass (two accumulator sets of 3 x 64 TMEM columns)

struct RollParams {
  alignas(64) unsigned char tmap_in[128];    // NHWC x as {C, W, H, B}, box {64, W+2, H+2, 1}, SWIZZLE_128B
  ChainArgs a;
  void* scratch[2];                          // block outputs 0, 1, 2 (ping-pong); the last block writes a.out
  int Wp, Hp, P, T, TP, NP, d;               // padded grid, tiles with interior positions, tiles per pass, passes per conv, drift
  int w_off, bar_off, stages, tmem_cols;
};

// barrier block (8-byte slots): x_full | x_free | acc_full[2] | acc_empty[2] | w_full[6] | w_empty[6] | tmem slot
template <bool F16>
__global__ void __launch_bounds__(RL_THREADS, 1)
conv_roll_kernel(const __grid_constant__ RollParams p) {
  constexpr int C = RL_C;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base, sW = base + (uint32_t)p.w_off, sBar = base + (uint32_t)p.bar_off;
  const uint32_t bar_xfull = sBar, bar_xfree = sBar + 8u, bar_af = sBar + 16u, bar_ae = sBar + 32u, bar_wf = sBar + 48u,
                 bar_we = bar_wf + 8u * RL_MAX_STAGES, tmem_slot = bar_we + 8u * RL_MAX_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainArgs& a = p.a;
  const int Wp = p.Wp, T = p.T, TP = p.TP, NP = p.NP, S = p.stages, nconv = a.nconv, d = p.d;
  constexpr uint32_t KB_BYTES = C * 128;     // one tap: 64 Cout rows x 64 Cin x 2 B

  if (tid == 0) {
    mbar_init(bar_xfull, 1); mbar_init(bar_xfree, RL_EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_af + 8u * i, RL_ISSUERS); mbar_init(bar_ae + 8u * i, RL_EPI_WARPS); }
    for (int i = 0; i < S; ++i) { mbar_init(bar_wf + 8u * i, 1); mbar_init(bar_we + 8u * i, RL_ISSUERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(p.tmap_in) : "memory");
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== loader ========================================================================================================
    const bool leader = elect_one();
    const uint32_t x_tx = (uint32_t)p.P * 128u;
    int s = 0, use = 0, li = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++li) {
      if (li >= 1) mbar_wait(bar_xfree, (li - 1) & 1);            // the previous image's last epilogue has left the buffer
      if (leader) {
        mbar_arrive_expect_tx(bar_xfull, x_tx);
        tma_load_4d(sX + (uint32_t)(nconv * d) * 128u, p.tmap_in, 0, -1, -1, b, bar_xfull);
      }
      for (int j = 0; j < nconv; ++j) {
        const uint8_t* wj = static_cast<const uint8_t*>(a.w[j]);
        for (int pass = 0; pass < NP; ++pass)
          for (int t = 0; t < 9; ++t) {
            if (use >= 1) mbar_wait(bar_we + 8u * s, (use - 1) & 1);
            if (leader) {
              mbar_arrive_expect_tx(bar_wf + 8u * s, KB_BYTES);
              bulk_g2s(sW + (uint32_t)s * KB_BYTES, wj + (size_t)t * KB_BYTES, KB_BYTES, bar_wf + 8u * s);
            }
            __syncwarp();
            if (++s == S) { s = 0; ++use; }
          }
      }
    }
  } else if (warp <= RL_ISSUERS) {
    // ===== MMA issuers (all lanes walk the loops, one elected lane issues; tc_ptx.h): issuer iw owns tiles m0 + iw, + 3, ... ===
    const bool leader = elect_one();
    const int iw = warp - 1;
    constexpr uint32_t FMT = F16 ? 0u : 1u;                     // operand format: IEEE half / bf16
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dhi = umma_desc_hi(128);
    int s = 0, use = 0, li = 0;
    int g = 0, waited = -1;                                        // global pass counter; highest pass whose epilogue has been awaited
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++li) {
      mbar_wait(bar_xfull, li & 1);
      for (int j = 0; j < nconv; ++j) {
        const uint32_t in_row = (uint32_t)((nconv - j) * d);       // first buffer row of conv j's input image
        for (int pass = 0; pass < NP; ++pass, ++g) {
          // accumulators of pass g-2 drained; input rows written: the passes of conv j-1 up to the one that covers this
          // pass's last tile + halo (the first conv of an image reads what the TMA load delivered)
          int target = g - 2;
          if (j >= 1) target = max(target, g - pass - NP + min(pass + 1, NP - 1));
          for (; waited < target; ) { ++waited; mbar_wait(bar_ae + 8u * (uint32_t)(waited & 1), (uint32_t)((waited >> 1) & 1)); }
          tc_fence_after();
          const int m0 = pass * TP, m1 = min(T, m0 + TP);
          const uint32_t acc0 = tmem_base + (uint32_t)((g & 1) * TP * C);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            mbar_wait(bar_wf + 8u * s, use & 1);
            tc_fence_after();
            // output position q reads input position q + (r-1) Wp + (s-1)
            const uint32_t a16 = ((sX >> 4) + (uint32_t)((int)in_row + m0 * 128 + (t / 3 - 1) * Wp + (t % 3 - 1)) * 8u) | (1u << 16);
            const uint32_t b16 = ((sW + (uint32_t)s * KB_BYTES) >> 4) | (1u << 16);
            for (int m = m0 + iw; m < m1; m += RL_ISSUERS) {
              const uint32_t tm = acc0 + (uint32_t)((m - m0) * C), am = a16 + (uint32_t)(m - m0) * 1024u;
              if (leader) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_lo<false>(tm, am + 2u * kk, b16 + 2u * kk, dhi, idesc, (t | kk) != 0 ? 1u : 0u);
              }
              __syncwarp();
            }
            if (leader) umma_commit(bar_we + 8u * s);
            __syncwarp();
            if (++s == S) { s = 0; ++use; }
          }
          if (leader) umma_commit(bar_af + 8u * (uint32_t)(g & 1));
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: warp e owns TMEM lane quarter (warp & 3) and column slice (e >> 2) of every tile ========================
    constexpr int CPT = RL_CPT, UPT = CPT / 8;                     // channels / 16-byte units per thread
    const int e = warp - 1 - RL_ISSUERS, quarter = warp & 3, slice = e >> 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int c_lo = slice * CPT;
    int g = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
      const size_t img = (size_t)b * a.H * a.W * C * 2;
      for (int j = 0; j < nconv; ++j) {
        const bool second = (j & 1) != 0;
        const int blk = j >> 1;
        const uint8_t* res8 = second ? (blk == 0 ? static_cast<const uint8_t*>(a.x) : static_cast<const uint8_t*>(p.scratch[(blk - 1) & 1])) + img : nullptr;
        uint8_t* out8 = second ? (j == nconv - 1 ? static_cast<uint8_t*>(a.out) : static_cast<uint8_t*>(p.scratch[blk & 1])) + img : nullptr;
        const uint32_t out_row = (uint32_t)((nconv - j - 1) * d);
        // this thread's bias values stay in registers for the whole conv (a global load per element inside the drain loop
        // kept the in-order epilogue warps on the long scoreboard: they, not the MMAs, bounded the kernel)
        float bv[CPT];
#pragma unroll
        for (int i = 0; i < CPT / 4; ++i) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.b[j] + c_lo) + i);
          bv[4 * i] = t4.x; bv[4 * i + 1] = t4.y; bv[4 * i + 2] = t4.z; bv[4 * i + 3] = t4.w;
        }
        // conv2: the block input (residual) of tile slot mi is requested one PASS ahead -- as soon as the slot's previous
        // contents are consumed -- so the L2 round trip hides behind the rest of the pass
        uint4 r4[RL_TP_MAX][UPT] = {};
        auto fetch_res = [&](int mi, int m) {
          const int q = m * 128 + quarter * 32 + lane;
          const int yy = q / Wp, xx = q - yy * Wp;
          if (m < T && q < p.P && yy >= 1 && yy <= a.H && xx >= 1 && xx <= a.W) {
            const uint8_t* src = res8 + ((size_t)(yy - 1) * a.W + (xx - 1)) * (C * 2) + (size_t)c_lo * 2;
#pragma unroll
            for (int k = 0; k < UPT; ++k) r4[mi][k] = *reinterpret_cast<const uint4*>(src + 16 * k);
          }
        };
        if (second) {
#pragma unroll
          for (int mi = 0; mi < RL_TP_MAX; ++mi) if (mi < TP) fetch_res(mi, mi);
        }
        for (int pass = 0; pass < NP; ++pass, ++g) {
          const int m0 = pass * TP, m1 = min(T, m0 + TP);
          mbar_wait(bar_af + 8u * (uint32_t)(g & 1), (uint32_t)((g >> 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int mi = 0; mi < RL_TP_MAX; ++mi) {
            const int m = m0 + mi;
            if (m >= m1) break;
            const int q = m * 128 + quarter * 32 + lane;
            const int yy = q / Wp, xx = q - yy * Wp;
            const bool in_grid = q < p.P;
            const bool ok = in_grid && yy >= 1 && yy <= a.H && xx >= 1 && xx <= a.W;
            const size_t pix = ok ? ((size_t)(yy - 1) * a.W + (xx - 1)) * (C * 2) + (size_t)c_lo * 2 : 0;
            uint32_t v[CPT];
            static_assert(CPT == 16, "one tcgen05.ld.x16 per tile and thread");
            tmem_ld16(t_lane + (uint32_t)((g & 1) * TP * C + mi * C + c_lo), v);
            tmem_ld_wait();
            if (in_grid) {
              const uint32_t R = out_row + (uint32_t)q;
              const uint32_t row = sX + R * 128u, swz = R & 7u;
#pragma unroll
              for (int k = 0; k < UPT; ++k) {                          // 16-byte unit k of this thread's slice: channels c_lo + 8k .. +7
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (ok) {
                  const uint32_t rr[4] = {r4[mi][k].x, r4[mi][k].y, r4[mi][k].z, r4[mi][k].w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float f0 = __uint_as_float(v[k * 8 + i * 2]) + bv[k * 8 + i * 2];
                    float f1 = __uint_as_float(v[k * 8 + i * 2 + 1]) + bv[k * 8 + i * 2 + 1];
                    if (second) {
                      const float2 x2 = unpack2<F16>(rr[i]);
                      f0 += x2.x; f1 += x2.y;
                    }
                    w[i] = pack2<F16>(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                  }
                  if (second) *reinterpret_cast<uint4*>(out8 + pix + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                const uint32_t un = (uint32_t)(slice * UPT + k);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((un ^ swz) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
            if (second && pass + 1 < NP) fetch_res(mi, m + TP);        // the slot is free: next pass's residual for it
          }
          if (pass == NP - 1 && e == 0) {
            // grid positions behind the last tile that holds an interior position are halo: zero them (the tiles never cover them)
            for (int q = T * 128 + lane; q < p.P; q += 32) {
              const uint32_t R = out_row + (uint32_t)q;
#pragma unroll
              for (uint32_t un = 0; un < 8; ++un)
                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sX + R * 128u + (un << 4)), "r"(0u) : "memory");
            }
          }
          fence_proxy_async();                                      // generic-proxy writes -> visible to tcgen05.mma
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_ae + 8u * (uint32_t)(g & 1));
            if (j == nconv - 1 && pass == NP - 1) mbar_arrive(bar_xfree);
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
EncodeTiledFn roll_encode_tiled() {
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

int roll_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool roll_plan(const ChainArgs& a, RollParams* p, size_t* smem) {
  static const int off = roll_env("HRP_NO_ROLL_FUSION", 0), force_s = roll_env("HRP_ROLL_STAGES", 0), force_tp = roll_env("HRP_ROLL_TP", 0);
  if (off || a.C != RL_C || a.H < 1 || a.W < 1 || a.B < 1 || a.W + 2 > 256 || a.H + 2 > 256) return false;
  if (a.nconv < 2 || a.nconv > 8 || (a.nconv & 1)) return false;
  if (roll_encode_tiled() == nullptr) return false;
  {
    ConvArgs g{};                                            // the weight images are the ones conv_tc / conv_slab would use
    g.B = 1; g.Hi = g.Ho = a.H; g.Wi = g.Wo = a.W; g.Cin = g.Cout = a.C; g.KH = g.KW = 3; g.stride = 1; g.pad_h = g.pad_w = 1; g.ld_out = a.C;
    if (conv_tc_row_bytes(g, 0, nullptr) != 128) return false;
  }
  p->a = a;
  p->Wp = a.W + 2; p->Hp = a.H + 2; p->P = p->Wp * p->Hp;
  p->T = (a.H * p->Wp + a.W) / 128 + 1;                     // tiles up to the last interior position
  p->d = (p->Wp + 1 + 3) / 4 * 4;                           // >= Wp + 1; nconv is even, so nconv*d is a whole number of 8-row swizzle groups
  int tp = std::min(RL_TP_MAX, p->T);                        // two accumulator sets of TP x 64 columns in 512 TMEM columns
  tp = (p->T + (p->T + tp - 1) / tp - 1) / ((p->T + tp - 1) / tp);   // even out the passes
  if (force_tp >= 1 && force_tp <= RL_TP_MAX) tp = std::min(force_tp, p->T);
  p->TP = tp;
  p->NP = (p->T + tp - 1) / tp;
  // rows: the drifting image, whole tiles of the first conv's input, and the lower halo the last tile's taps reach into
  const int rows = a.nconv * p->d + std::max(p->P, 128 * p->T) + p->Wp + 1;
  const size_t xbytes = ((size_t)rows * 128 + 1023) / 1024 * 1024;
  const size_t tail = 256;
  int S = force_s ? force_s : RL_MAX_STAGES;
  S = std::min(S, RL_MAX_STAGES);
  while (S >= 2 && 1024 + xbytes + (size_t)S * RL_C * 128 + tail > (size_t)RL_SMEM_LIMIT) --S;
  if (S < 2) return false;
  p->stages = S;
  p->w_off = (int)xbytes;
  p->bar_off = p->w_off + S * RL_C * 128;
  int tm = 32;
  while (tm < 2 * tp * RL_C) tm <<= 1;
  if (tm > 512) return false;
  p->tmem_cols = tm;
  *smem = 1024 + (size_t)p->bar_off + tail;
  return *smem <= (size_t)RL_SMEM_LIMIT;
}

}  // namespace

bool conv_roll_supported(const ChainArgs& a) {
  RollParams p{};
  size_t smem = 0;
  return roll_plan(a, &p, &smem);
}

// scratch0 / scratch1: two [B,H,W,64] bf16 buffers for the outputs of the inner blocks (unused for a single block)
int conv_roll_launch(const ChainArgs& a, void* scratch0, void* scratch1, cudaStream_t st) {
  RollParams p{};
  size_t smem = 0;
  if (!roll_plan(a, &p, &smem)) return fail(HRP_ERR_INVALID, "conv_roll: unsupported chain (C=%d, %dx%d, %d convs)", a.C, a.H, a.W, a.nconv);
  if (a.nconv > 2 && (!scratch0 || (a.nconv > 4 && !scratch1))) return fail(HRP_ERR_INVALID, "conv_roll: scratch buffers missing");
  p.scratch[0] = scratch0; p.scratch[1] = scratch1;
  const cuuint64_t gdim[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W * a.C * 2, (cuuint64_t)a.H * a.W * a.C * 2};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {64, (cuuint32_t)p.Wp, (cuuint32_t)p.Hp, 1};
  {
    CUtensorMap tm;
    const CUresult r = roll_encode_tiled()(&tm, a.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.x), gdim, gstr, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HRP_ERR_CUDA, "conv_roll: cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::memcpy(p.tmap_in, &tm, 128);
  }
  static bool attr_done = false;
  if (!attr_done) {
    HRP_CUDA(cudaFuncSetAttribute(conv_roll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
    HRP_CUDA(cudaFuncSetAttribute(conv_roll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_LIMIT));
    attr_done = true;
  }
  // one image per CTA and one CTA per SM (the buffer fills the shared memory): the images ARE the grid
  const int grid = std::min(a.B, sm_count());
  if (a.f16) conv_roll_kernel<true><<<grid, RL_THREADS, smem, st>>>(p);
  else conv_roll_kernel<false><<<grid, RL_THREADS, smem, st>>>(p);
  HRP_CHECK_LAUNCH("conv_roll_kernel");
  return HRP_OK;
}

}  // namespace hrp
