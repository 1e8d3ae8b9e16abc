// fp32 implicit-GEMM convolution (FFMA) + the small non-GEMM layers of the backbones.
//
// This is the HRP_PREC_FP32 parity family: the same arithmetic type the reference uses for nn.Conv2d / nn.Linear /
// nn.ConvTranspose2d (lib/models/backbones/HRnet.py, Resnet.py, full_net.py:214-238), with BatchNorm folded into the
// weights and the bias / residual-add / ReLU of Bottleneck.forward and BasicBlock.forward (HRnet.py:41-98) applied in
// the epilogue. The tensor-core families (conv_tc.cu) share the ConvArgs descriptor.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "kernels.h"

namespace hrp {

constexpr int CV_THREADS = 256;
constexpr int CV_BK = 16;

// BM x BN output tile per CTA, 16x16 threads, (BM/16)x(BN/16) outputs per thread.
// MAJOR_M: adjacent threads own adjacent pixels (coalesced NCHW stores); otherwise adjacent channels (NHWC stores).
template <int BM, int BN, bool MAJOR_M>
__global__ void __launch_bounds__(CV_THREADS)
conv_igemm_f32_kernel(const ConvArgs p) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_LD = BM + 4, B_LD = BN + 4;
  constexpr int A_F4 = BM * CV_BK / 4 / CV_THREADS;            // float4 A loads per thread per k-tile
  constexpr int B_F4_TOTAL = CV_BK * BN / 4;                   // float4 B loads per CTA per k-tile
  __shared__ __align__(16) float As[2][CV_BK][A_LD];
  __shared__ __align__(16) float Bs[2][CV_BK][B_LD];

  const float* __restrict__ in = static_cast<const float*>(p.in);
  const float* __restrict__ wt = static_cast<const float*>(p.w);
  const int tid = threadIdx.x;
  const int tm = MAJOR_M ? (tid & 15) : (tid >> 4);
  const int tn = MAJOR_M ? (tid >> 4) : (tid & 15);
  const int M = p.B * p.Ho * p.Wo;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ktiles = p.KH * p.KW * p.Cin / CV_BK;

  // per-thread A-load coordinates (fixed over the K loop)
  int a_row[A_F4], a_kq[A_F4], a_iy0[A_F4], a_ix0[A_F4];
  const float* a_base[A_F4];
  bool a_ok[A_F4];
#pragma unroll
  for (int l = 0; l < A_F4; ++l) {
    const int idx = tid + l * CV_THREADS;
    a_row[l] = idx >> 2;
    a_kq[l] = idx & 3;
    const int m = m0 + a_row[l];
    a_ok[l] = m < M;
    const int mm = a_ok[l] ? m : 0;
    const int ox = mm % p.Wo, r = mm / p.Wo, oy = r % p.Ho, b = r / p.Ho;
    a_iy0[l] = oy * p.stride - p.pad_h;
    a_ix0[l] = ox * p.stride - p.pad_w;
    a_base[l] = in + (size_t)b * p.Hi * p.Wi * p.Cin;
  }
  const int b_row = tid / (BN / 4), b_col = (tid % (BN / 4)) * 4;
  const bool b_active = tid < B_F4_TOTAL;
  const bool b_ok = b_active && (n0 + b_col < p.Cout);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int c0 = 0, fr = 0, fs = 0;  // k-tile -> (channel offset, filter row, filter col)
  float4 a_reg[A_F4], b_reg;

  auto load_tile = [&](int kt) {
#pragma unroll
    for (int l = 0; l < A_F4; ++l) {
      const int iy = a_iy0[l] + fr, ix = a_ix0[l] + fs;
      const bool ok = a_ok[l] && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi;
      a_reg[l] = ok ? __ldg(reinterpret_cast<const float4*>(a_base[l] + ((size_t)iy * p.Wi + ix) * p.Cin + c0 + a_kq[l] * 4))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    b_reg = b_ok ? __ldg(reinterpret_cast<const float4*>(wt + (size_t)(kt * CV_BK + b_row) * p.Cout + n0 + b_col))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    c0 += CV_BK;
    if (c0 == p.Cin) { c0 = 0; if (++fs == p.KW) { fs = 0; ++fr; } }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int l = 0; l < A_F4; ++l) {
      As[buf][a_kq[l] * 4 + 0][a_row[l]] = a_reg[l].x;
      As[buf][a_kq[l] * 4 + 1][a_row[l]] = a_reg[l].y;
      As[buf][a_kq[l] * 4 + 2][a_row[l]] = a_reg[l].z;
      As[buf][a_kq[l] * 4 + 3][a_row[l]] = a_reg[l].w;
    }
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_row][b_col]) = b_reg;
  };

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < CV_BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][tm * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
      if (TN >= 4) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * TN + j]);
          b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
        }
      } else {
        const float2 v = *reinterpret_cast<const float2*>(&Bs[buf][k][tn * TN]);
        b[0] = v.x; b[1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: + bias (BN folded) [+ residual] [ReLU]
  const float* __restrict__ res = static_cast<const float*>(p.res);
  float* __restrict__ out = static_cast<float*>(p.out);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm * TM + i;
    if (m >= M) continue;
    const int ox = m % p.Wo, r = m / p.Wo, oy = r % p.Ho, b = r / p.Ho;
    const int y = oy * p.out_sy + p.out_oy, x = ox * p.out_sx + p.out_ox;
    const size_t pix = ((size_t)b * p.Ho_full + y) * p.Wo_full + x;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j] + __ldg(p.bias + n);
      if (p.out_nchw) {
        const size_t o = (((size_t)b * p.Cout + n) * p.Ho_full + y) * p.Wo_full + x;
        if (res && !p.res_after_act) v += __ldg(res + o);
        if (p.relu) v = fmaxf(v, 0.f);
        if (res && p.res_after_act) v += __ldg(res + o);
        out[o] = v;
      } else {
        const size_t o = pix * p.ld_out + p.out_coff + n;
        if (res && !p.res_after_act) v += __ldg(res + o);
        if (p.relu) v = fmaxf(v, 0.f);
        if (res && p.res_after_act) v += __ldg(res + o);
        out[o] = v;
      }
    }
  }
}

// ---- nn.Linear on a few rows: y[M][N] = x[M][K] . W[K][N] + b (full_net.py:214-238, the regression heads) ---------------
// M is the batch (64 frames): one 64x64 tile per 64 output columns leaves 16-32 CTAs walking K = 1024-2048 alone, 85-300 us
// for 0.1-0.5 GFLOP. Here a CTA owns 64 rows x 16 columns, so N/16 x ceil(M/64) CTAs share the weight matrix (each reads
// its 16 columns once, 4-8 MB in total) and the K loop is a six-deep cp.async ring of 32-wide chunks (x: 64x32, W: 32x16),
// enough bytes in flight to hide the L2 latency. Plain fp32 FMAs in ascending k for every output, independent of M: the
// frames of a batch stay bit-identical to the same frames run alone. Measured (ncu, alone): 27.8 us for 1024 -> 1024 and
// 52 us for 2048 -> 2048 at 64 rows (the 64x64-tile kernel: 67 us on average), i.e. 0.85 us per chunk -- bound by the
// shared-memory pipe (40 LDS.128 per thread per chunk at 4 cycles each), not by the copies; a 4x4 register tile with
// the chunk's k split over four thread groups would cut that 2.5x. The heads are off the critical path of the graph
// (they overlap the deconv head), so neither frames/s nor the single-call latency moves.
constexpr int LS_TM = 64, LS_TN = 16, LS_KC = 32, LS_STAGES = 6, LS_XLD = LS_KC + 4;
constexpr int LS_STAGE_FLOATS = LS_TM * LS_XLD + LS_KC * LS_TN;

__global__ void __launch_bounds__(256)
linear_skinny_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                         int M, int K, int N, int ld_out, int relu) {
  extern __shared__ __align__(16) float ls_smem[];
  const int tid = threadIdx.x, tm = tid >> 2, tc = (tid & 3) * 4;
  const int n0 = blockIdx.x * LS_TN, m0 = blockIdx.y * LS_TM;
  const int chunks = K / LS_KC;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ls_smem);
  auto issue = [&](int c) {
    if (c < chunks) {
      const int st = c % LS_STAGES, k0 = c * LS_KC;
      const uint32_t xs = sbase + (uint32_t)(st * LS_STAGE_FLOATS) * 4u, ws = xs + (uint32_t)(LS_TM * LS_XLD) * 4u;
#pragma unroll
      for (int l = 0; l < LS_TM * LS_KC / 4 / 256; ++l) {          // 64 rows x 8 float4
        const int idx = tid + l * 256, r = idx >> 3, q = idx & 7;
        const bool ok = m0 + r < M;
        const float* src = x + (size_t)(ok ? m0 + r : 0) * K + k0 + q * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs + (uint32_t)(r * LS_XLD + q * 4) * 4u), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      }
      if (tid < LS_KC * LS_TN / 4) {                               // 32 rows x 4 float4
        const int r = tid >> 2, q = tid & 3;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(ws + (uint32_t)(r * LS_TN + q * 4) * 4u), "l"(w + (size_t)(k0 + r) * N + n0 + q * 4) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int c = 0; c < LS_STAGES - 1; ++c) issue(c);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < chunks; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(LS_STAGES - 2) : "memory");
    __syncthreads();                                               // chunk c has landed; everyone is done with chunk c-1's slot
    issue(c + LS_STAGES - 1);
    const float* xs = ls_smem + (c % LS_STAGES) * LS_STAGE_FLOATS + tm * LS_XLD;
    const float* ws = ls_smem + (c % LS_STAGES) * LS_STAGE_FLOATS + LS_TM * LS_XLD + tc;
#pragma unroll
    for (int kk = 0; kk < LS_KC; kk += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + kk);
      const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + (kk + j) * LS_TN);
        acc[0] = fmaf(xe[j], wv.x, acc[0]); acc[1] = fmaf(xe[j], wv.y, acc[1]);
        acc[2] = fmaf(xe[j], wv.z, acc[2]); acc[3] = fmaf(xe[j], wv.w, acc[3]);
      }
    }
  }
  const int m = m0 + tm;
  if (m < M) {
    const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + tc)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o = make_float4(acc[0] + b4.x, acc[1] + b4.y, acc[2] + b4.z, acc[3] + b4.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(y + (size_t)m * ld_out + n0 + tc) = o;
  }
}

static bool linear_skinny_ok(const ConvArgs& a) {
  static const bool off = [] { const char* v = getenv("HRP_NO_SKINNY_LINEAR"); return v && atoi(v) != 0; }();
  return !off && a.KH == 1 && a.KW == 1 && a.Hi == 1 && a.Wi == 1 && a.Ho == 1 && a.Wo == 1 && a.stride == 1 && a.pad_h == 0 && a.pad_w == 0 &&
         !a.out_nchw && a.res == nullptr && a.out_sy == 1 && a.out_sx == 1 && a.out_oy == 0 && a.out_ox == 0 && a.out_coff == 0 &&
         a.Cin % LS_KC == 0 && a.Cout % LS_TN == 0 && a.ld_out % 4 == 0 && a.Cin / LS_KC >= LS_STAGES;
}

int conv_f32_launch(const ConvArgs& a, cudaStream_t s) {
  if (linear_skinny_ok(a) && a.B > 0) {
    constexpr size_t smem = (size_t)LS_STAGES * LS_STAGE_FLOATS * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
      HRP_CUDA(cudaFuncSetAttribute(linear_skinny_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    dim3 grid(a.Cout / LS_TN, ceil_div(a.B, LS_TM));
    linear_skinny_f32_kernel<<<grid, 256, smem, s>>>(static_cast<const float*>(a.in), static_cast<const float*>(a.w), a.bias,
                                                     static_cast<float*>(a.out), a.B, a.Cin, a.Cout, a.ld_out, a.relu);
    HRP_CHECK_LAUNCH("linear_skinny_f32_kernel");
    return HRP_OK;
  }
  if (a.Cin % CV_BK != 0) return fail(HRP_ERR_INVALID, "conv_f32: Cin=%d must be a multiple of %d", a.Cin, CV_BK);
  if (a.Cout % 4 != 0 && a.Cout > 4) return fail(HRP_ERR_INVALID, "conv_f32: Cout=%d must be a multiple of 4", a.Cout);
  const int M = a.B * a.Ho * a.Wo;
  if (M <= 0) return HRP_OK;
  if (a.out_nchw) {
    dim3 grid(ceil_div(M, 64), ceil_div(a.Cout, 64));
    conv_igemm_f32_kernel<64, 64, true><<<grid, CV_THREADS, 0, s>>>(a);
  } else if (a.Cout <= 32) {
    dim3 grid(ceil_div(M, 128), ceil_div(a.Cout, 32));
    conv_igemm_f32_kernel<128, 32, false><<<grid, CV_THREADS, 0, s>>>(a);
  } else if (M >= 128 * 2 * sm_count()) {
    dim3 grid(ceil_div(M, 128), ceil_div(a.Cout, 64));
    conv_igemm_f32_kernel<128, 64, false><<<grid, CV_THREADS, 0, s>>>(a);
  } else {
    dim3 grid(ceil_div(M, 64), ceil_div(a.Cout, 64));
    conv_igemm_f32_kernel<64, 64, false><<<grid, CV_THREADS, 0, s>>>(a);
  }
  HRP_CHECK_LAUNCH("conv_igemm_f32_kernel");
  return HRP_OK;
}

// ---- stem: Cin = 3 straight from the NCHW input image --------------------------------------------------------------------
// HRnet.py:500-502 (3x3 s2 p1) and Resnet.py:58-60 (7x7 s2 p3). K = 27 / 147 is too thin for a tensor-core tile; the
// layer is 0.06 / 0.31 GFLOP per frame. 64 pixels x 4 channel groups per CTA, weights broadcast from shared memory.
template <typename OutT, bool ROUND_TF32>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 OutT* __restrict__ out, int B, int Hi, int Wi, int Ho, int Wo, int KH, int KW, int pad) {
  extern __shared__ __align__(16) float sw[];  // [3*KH*KW][64]
  const int K = 3 * KH * KW;
  for (int i = threadIdx.x; i < K * 16; i += 256)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  const int g = threadIdx.x >> 6;
  const long long pix = (long long)blockIdx.x * 64 + (threadIdx.x & 63);
  const long long total = (long long)B * Ho * Wo;
  if (pix >= total) return;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int b = (int)(pix / ((long long)Wo * Ho));
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = __ldg(bias + g * 16 + j);
  const int iy0 = oy * 2 - pad, ix0 = ox * 2 - pad;
  int k = 0;
  for (int c = 0; c < 3; ++c) {
    const float* ip = in + ((size_t)b * 3 + c) * Hi * Wi;
    for (int r = 0; r < KH; ++r) {
      const int iy = iy0 + r;
      for (int s = 0; s < KW; ++s, ++k) {
        const int ix = ix0 + s;
        const float v = (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) ? __ldg(ip + (size_t)iy * Wi + ix) : 0.f;
        const float4* wk = reinterpret_cast<const float4*>(sw + k * 64 + g * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ww = wk[j];
          acc[j * 4 + 0] = fmaf(v, ww.x, acc[j * 4 + 0]);
          acc[j * 4 + 1] = fmaf(v, ww.y, acc[j * 4 + 1]);
          acc[j * 4 + 2] = fmaf(v, ww.z, acc[j * 4 + 2]);
          acc[j * 4 + 3] = fmaf(v, ww.w, acc[j * 4 + 3]);
        }
      }
    }
  }
  OutT* op = out + (size_t)pix * 64 + g * 16;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float v = fmaxf(acc[j], 0.f);
    if constexpr (ROUND_TF32) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
    if constexpr (sizeof(OutT) == 4) op[j] = v; else op[j] = __float2bfloat16_rn(v);
  }
}

int stem_conv_launch(const float* in_nchw, const float* w, const float* bias, void* out, int B, int Hi, int Wi,
                     int Ho, int Wo, int KH, int KW, int pad, int out_mode, cudaStream_t s) {
  const long long total = (long long)B * Ho * Wo;
  if (total <= 0) return HRP_OK;
  const size_t smem = (size_t)3 * KH * KW * 64 * sizeof(float);
  const unsigned blocks = (unsigned)ceil_div64(total, 64);
  if (out_mode == 1) {
    stem_conv_kernel<__nv_bfloat16, false><<<blocks, 256, smem, s>>>(in_nchw, w, bias, static_cast<__nv_bfloat16*>(out), B, Hi, Wi, Ho, Wo, KH, KW, pad);
  } else if (out_mode == 2) {
    stem_conv_kernel<float, true><<<blocks, 256, smem, s>>>(in_nchw, w, bias, static_cast<float*>(out), B, Hi, Wi, Ho, Wo, KH, KW, pad);
  } else {
    stem_conv_kernel<float, false><<<blocks, 256, smem, s>>>(in_nchw, w, bias, static_cast<float*>(out), B, Hi, Wi, Ho, Wo, KH, KW, pad);
  }
  HRP_CHECK_LAUNCH("stem_conv_kernel");
  return HRP_OK;
}

// NCHW fp32 image -> zero-padded NHWC4 operand image of the tensor-core stems (kernels.h). One thread per padded pixel;
// the three plane reads are coalesced along x, the write is one 8- or 16-byte store. Borders are rewritten every
// forward because the arena recycles this memory.
template <int MODE>      // 0 bf16, 1 fp32 rounded to TF32, 2 fp32 as is, 3 IEEE half
__global__ void stem_pack_kernel(const float* __restrict__ in, void* __restrict__ out, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * STEM_HP * STEM_WP;
  if (i >= total) return;
  const int xp = (int)(i % STEM_WP);
  const int yp = (int)((i / STEM_WP) % STEM_HP);
  const int b = (int)(i / ((long long)STEM_WP * STEM_HP));
  const int x = xp - STEM_PAD, y = yp - STEM_PAD;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (x >= 0 && x < 256 && y >= 0 && y < 256) {
    const float* ip = in + ((size_t)b * 3 * 256 + y) * 256 + x;
    v0 = __ldg(ip); v1 = __ldg(ip + 256 * 256); v2 = __ldg(ip + 2 * 256 * 256);
  }
  if constexpr (MODE == 2) {
    reinterpret_cast<uint4*>(out)[i] = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), 0u);
  } else if constexpr (MODE == 1) {
    uint32_t r0, r1, r2;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(v0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(v1));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(v2));
    reinterpret_cast<uint4*>(out)[i] = make_uint4(r0, r1, r2, 0u);
  } else if constexpr (MODE == 3) {
    const __half2 a = __floats2half2_rn(v0, v1), c = __floats2half2_rn(v2, 0.f);       // inputs are in [0, 1]
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&c);
    reinterpret_cast<uint2*>(out)[i] = r;
  } else {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), c = __floats2bfloat162_rn(v2, 0.f);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&c);
    reinterpret_cast<uint2*>(out)[i] = r;
  }
}

int stem_pack_launch(const float* in_nchw, void* out, int B, int tf32, cudaStream_t s) {
  const long long total = (long long)B * STEM_HP * STEM_WP;
  if (total <= 0) return HRP_OK;
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (tf32 == 3) stem_pack_kernel<3><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else if (tf32 == 2) stem_pack_kernel<2><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else if (tf32) stem_pack_kernel<1><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else stem_pack_kernel<0><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  HRP_CHECK_LAUNCH("stem_pack_kernel");
  return HRP_OK;
}

// ---- element-wise layers: 4 channels per thread, NHWC ---------------------------------------------------------------------
// nn.MaxPool2d(3, 2, 1), Resnet.py:22,61. 16 bytes of channels per thread (8 bf16 / 4 fp32), 32-bit index arithmetic, the
// nine taps loaded before the first max (edge taps re-read the centre pixel: max is idempotent), bf16 maxima taken in
// bf16 (exact).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Hi, int Wi, int Ho, int Wo, int C) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  const uint32_t cv = (uint32_t)C / V;
  const uint32_t total = (uint32_t)B * Ho * Wo * cv;
  if (i >= total) return;
  const uint32_t c = (i % cv) * V;
  uint32_t r = i / cv;
  const int ox = (int)(r % (uint32_t)Wo); r /= (uint32_t)Wo;
  const int oy = (int)(r % (uint32_t)Ho);
  const uint32_t b = r / (uint32_t)Ho;
  const T* img = in + (size_t)b * Hi * Wi * C + c;
  uint4 v[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= Hi) iy = oy * 2;                    // the window's centre row is always inside
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= Wi) ix = ox * 2;
      v[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)iy * Wi + ix) * C));
    }
  }
  uint4 m = v[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) {
    if constexpr (sizeof(T) == 4) {
      m.x = __float_as_uint(fmaxf(__uint_as_float(m.x), __uint_as_float(v[k].x))); m.y = __float_as_uint(fmaxf(__uint_as_float(m.y), __uint_as_float(v[k].y)));
      m.z = __float_as_uint(fmaxf(__uint_as_float(m.z), __uint_as_float(v[k].z))); m.w = __float_as_uint(fmaxf(__uint_as_float(m.w), __uint_as_float(v[k].w)));
    } else {
      auto mx = [](uint32_t p, uint32_t q) {
        if constexpr (std::is_same<T, __half>::value) {
          const __half2 h = __hmax2(*reinterpret_cast<const __half2*>(&p), *reinterpret_cast<const __half2*>(&q));
          return *reinterpret_cast<const uint32_t*>(&h);
        } else {
          const __nv_bfloat162 h = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&p), *reinterpret_cast<const __nv_bfloat162*>(&q));
          return *reinterpret_cast<const uint32_t*>(&h);
        }
      };
      m.x = mx(m.x, v[k].x); m.y = mx(m.y, v[k].y); m.z = mx(m.z, v[k].z); m.w = mx(m.w, v[k].w);
    }
  }
  *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + oy) * Wo + ox) * C + c) = m;
}

int maxpool3x3s2_launch(const void* in, void* out, int B, int Hi, int Wi, int C, int bf16, cudaStream_t s) {
  const int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
  const int V = bf16 ? 8 : 4;
  const long long total = (long long)B * Ho * Wo * (C / V);
  if (total <= 0) return HRP_OK;
  if (C % V || total > 0x7fffffffLL) return fail(HRP_ERR_INVALID, "maxpool: unsupported shape (C=%d, %lld vectors)", C, total);
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (bf16 == 2) maxpool3x3s2_kernel<__half><<<blocks, 256, 0, s>>>(static_cast<const __half*>(in), static_cast<__half*>(out), B, Hi, Wi, Ho, Wo, C);
  else if (bf16) maxpool3x3s2_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), B, Hi, Wi, Ho, Wo, C);
  else maxpool3x3s2_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(in), static_cast<float*>(out), B, Hi, Wi, Ho, Wo, C);
  HRP_CHECK_LAUNCH("maxpool3x3s2_kernel");
  return HRP_OK;
}

// HighResolutionModule.forward fusion sum (HRnet.py:256-263): same-resolution terms plus nearest-upsampled
// low-resolution terms (nn.Upsample(scale_factor=2^(j-i), mode='nearest'), HRnet.py:206), then ReLU.
// 16 bytes of channels per thread (8 bf16 / 4 fp32), 32-bit index arithmetic, all terms' loads issued before the adds.
template <typename T>
__global__ void __launch_bounds__(256) fuse_sum_kernel(const FuseArgs a) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  const uint32_t cv = (uint32_t)a.C / V;
  const uint32_t total = (uint32_t)a.B * a.H * a.W * cv;
  if (i >= total) return;
  const uint32_t c = (i % cv) * V;
  uint32_t r = i / cv;
  const uint32_t x = r % (uint32_t)a.W; r /= (uint32_t)a.W;
  const uint32_t y = r % (uint32_t)a.H;
  const uint32_t b = r / (uint32_t)a.H;
  const size_t o = (size_t)(i / cv) * a.C + c;
  // every term's load is issued before the first add; the loops are unrolled over the maximum term counts with
  // compile-time indices so that the vectors (and the argument struct) stay in registers / constant space -- run-time
  // indexing put both into a 224-byte local-memory frame
  uint4 rs[4], rl[3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < a.n_same) rs[k] = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(a.same[k]) + o));
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < a.n_low) {
      const int sh = a.shift[k];
      const uint32_t h = (uint32_t)a.H >> sh, w = (uint32_t)a.W >> sh;
      rl[k] = __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(a.low[k]) + ((size_t)(b * h + (y >> sh)) * w + (x >> sh)) * a.C + c));
    }
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  auto add = [&](const uint4& v) {
    if constexpr (sizeof(T) == 4) {
      acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y); acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
    } else {
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f;
        if constexpr (std::is_same<T, __half>::value) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
        else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
        acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
      }
    }
  };
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < a.n_same) add(rs[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < a.n_low) add(rl[k]);
  if (a.relu) {
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = fmaxf(acc[e], 0.f);
  }
  uint4 outv;
  if constexpr (sizeof(T) == 4) {
    if (a.round_tf32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { uint32_t t; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(acc[e])); acc[e] = __uint_as_float(t); }
    }
    outv = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
  } else {
    uint32_t w4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if constexpr (std::is_same<T, __half>::value) {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w4[e]) : "f"(acc[2 * e + 1]), "f"(acc[2 * e]));
      } else {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
        w4[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    outv = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
  *reinterpret_cast<uint4*>(static_cast<T*>(a.out) + o) = outv;
}

int fuse_sum_launch(const FuseArgs& a, int bf16, cudaStream_t s) {
  const int V = bf16 ? 8 : 4;
  const long long total = (long long)a.B * a.H * a.W * (a.C / V);
  if (total <= 0) return HRP_OK;
  if (a.C % V || total > 0x7fffffffLL || a.n_same > 4 || a.n_low > 3) return fail(HRP_ERR_INVALID, "fuse_sum: unsupported shape (C=%d, %lld vectors)", a.C, total);
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (bf16 == 2) fuse_sum_kernel<__half><<<blocks, 256, 0, s>>>(a);
  else if (bf16) fuse_sum_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(a);
  else fuse_sum_kernel<float><<<blocks, 256, 0, s>>>(a);
  HRP_CHECK_LAUNCH("fuse_sum_kernel");
  return HRP_OK;
}

// global average pool: nn.AvgPool2d(8) on [B,2048,8,8] (full_net.py:82,350) and F.avg_pool2d over the whole map
// (HRnet.py:567). One thread per (frame, channel), channels adjacent across the warp.
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  const T* p = in + (size_t)b * HW * C + c;
  float s = 0.f;
  for (int k = 0; k < HW; ++k) {
    if constexpr (sizeof(T) == 4) s += __ldg(p + (size_t)k * C);
    else if constexpr (std::is_same<T, __half>::value) s += __half2float(p[(size_t)k * C]);
    else s += __bfloat162float(p[(size_t)k * C]);
  }
  out[i] = s / (float)HW;
}

int avgpool_launch(const void* in, float* out, int B, int HW, int C, int bf16, cudaStream_t s) {
  if (B * C <= 0) return HRP_OK;
  if (bf16 == 2) avgpool_kernel<__half><<<ceil_div(B * C, 256), 256, 0, s>>>(static_cast<const __half*>(in), out, B, HW, C);
  else if (bf16) avgpool_kernel<__nv_bfloat16><<<ceil_div(B * C, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in), out, B, HW, C);
  else avgpool_kernel<float><<<ceil_div(B * C, 256), 256, 0, s>>>(static_cast<const float*>(in), out, B, HW, C);
  HRP_CHECK_LAUNCH("avgpool_kernel");
  return HRP_OK;
}

// ---- heads --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
depth_head_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ k_value, float* __restrict__ depth, int C) {
  const int b = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) s = fmaf(__ldg(feat + (size_t)b * C + c), __ldg(w + c), s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    const float gamma = t + bias[0];                       // depth_layer, full_net.py:315
    depth[b] = gamma * k_value[b] / 1000.0f;               // full_net.py:334-336
  }
}

int depth_head_launch(const float* feat, const float* w, const float* bias, const float* k_value, float* depth, int B,
                      int C, cudaStream_t s) {
  if (B <= 0) return HRP_OK;
  depth_head_kernel<<<B, 256, 0, s>>>(feat, w, bias, k_value, depth, C);
  HRP_CHECK_LAUNCH("depth_head_kernel");
  return HRP_OK;
}

// reg_joint_map head (full_net.py:92-101, 376-379): joint_final_layer (1x1 conv, d2 -> dof, bias) on the [B,HW,d2] NHWC map
// that joint_conv_layers left, then HeatmapIntegralJoint (lib/utils/integral.py:211-251): softmax over the HW positions of
// each joint's map, expected position index / HW in [0,1), scaled into the joint's bounds. One CTA per frame; the logits
// (dof x HW floats) never leave shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
joint_map_head_kernel(const T* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ bounds, float* __restrict__ pose, int HW, int C, int dof) {
  extern __shared__ float jm[];                       // logits [dof][HW]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* yb = y + (size_t)b * HW * C;
  for (int o = warp; o < dof * HW; o += 8) {          // one warp per (joint, position): a C-long dot product
    const int j = o / HW, p = o - j * HW;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      float v;
      if constexpr (sizeof(T) == 4) v = __ldg(reinterpret_cast<const float*>(yb) + (size_t)p * C + c);
      else if constexpr (std::is_same<T, __half>::value) v = __half2float(yb[(size_t)p * C + c]);
      else v = __bfloat162float(yb[(size_t)p * C + c]);
      acc = fmaf(__ldg(w + (size_t)j * C + c), v, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) jm[o] = acc + bias[j];
  }
  __syncthreads();
  for (int j = warp; j < dof; j += 8) {
    const float* l = jm + j * HW;
    float mx = -INFINITY;
    for (int p = lane; p < HW; p += 32) mx = fmaxf(mx, l[p]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float se = 0.f, sp = 0.f;
    for (int p = lane; p < HW; p += 32) { const float e = expf(l[p] - mx); se += e; sp += e * (float)p; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, off); sp += __shfl_xor_sync(0xffffffffu, sp, off); }
    if (lane == 0) {
      const float coord = (sp / se) / (float)HW;                       // integral.py:239-243
      const float lo = bounds[2 * j], hi = bounds[2 * j + 1];
      pose[(size_t)b * dof + j] = coord * (hi - lo) + lo;               // integral.py:245-249
    }
  }
}

int joint_map_head_launch(const void* y, const float* w, const float* bias, const float* bounds, float* pose, int B, int HW, int C,
                          int dof, int bf16, cudaStream_t s) {
  if (B <= 0) return HRP_OK;
  const size_t sm = (size_t)dof * HW * sizeof(float);
  if (bf16 == 2) joint_map_head_kernel<__half><<<B, 256, sm, s>>>(static_cast<const __half*>(y), w, bias, bounds, pose, HW, C, dof);
  else if (bf16) joint_map_head_kernel<__nv_bfloat16><<<B, 256, sm, s>>>(static_cast<const __nv_bfloat16*>(y), w, bias, bounds, pose, HW, C, dof);
  else joint_map_head_kernel<float><<<B, 256, sm, s>>>(static_cast<const float*>(y), w, bias, bounds, pose, HW, C, dof);
  HRP_CHECK_LAUNCH("joint_map_head_kernel");
  return HRP_OK;
}

// DepthNet head of the constructor variants (full_net.py:293-330): add_fc -- the bottleneck MLP 2048 -> 1024 -> 512 ->
// BatchNorm1d -> LeakyReLU -> 1024 -> 2048 with two averaged skips -- and multi_kp -- `dn` depth outputs per frame. Everything
// around the LeakyReLU is linear, so (composed in fp64 at hrp_finalize_weights, network.cu)
//   z = lrelu(Wz f + bz)  [Z = 512],   gamma_d = Af_d . f + Bz_d . z + c_d,   depth_d = gamma_d * k / 1000.
// Z == 0: no MLP (multi_kp alone), gamma_d = Af_d . f + c_d. One CTA per frame.
__global__ void __launch_bounds__(256)
depth_head_ex_kernel(const float* __restrict__ feat, const float* __restrict__ Af, const float* __restrict__ Bz,
                     const float* __restrict__ Wz, const float* __restrict__ bz, const float* __restrict__ c,
                     const float* __restrict__ k_value, float* __restrict__ depth, float* __restrict__ depths, int C, int Z, int dn,
                     int root_index) {
  extern __shared__ float dsm[];                      // f [C], z [Z]
  float* f = dsm;
  float* z = dsm + C;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 4; i < C; i += 256 * 4)
    *reinterpret_cast<float4*>(f + i) = __ldg(reinterpret_cast<const float4*>(feat + (size_t)b * C + i));
  __syncthreads();
  for (int r = warp; r < Z; r += 8) {
    const float4* wr = reinterpret_cast<const float4*>(Wz + (size_t)r * C);
    float acc = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
      const float4 w = __ldg(wr + i);
      const float4 x = *reinterpret_cast<const float4*>(f + 4 * i);
      acc = fmaf(w.x, x.x, acc); acc = fmaf(w.y, x.y, acc); acc = fmaf(w.z, x.z, acc); acc = fmaf(w.w, x.w, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const float v = acc + bz[r];
      z[r] = v > 0.f ? v : 0.01f * v;                 // nn.LeakyReLU() default slope
    }
  }
  __syncthreads();
  for (int d = warp; d < dn; d += 8) {
    float acc = 0.f;
    for (int i = lane; i < C; i += 32) acc = fmaf(__ldg(Af + (size_t)d * C + i), f[i], acc);
    for (int i = lane; i < Z; i += 32) acc = fmaf(__ldg(Bz + (size_t)d * Z + i), z[i], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const float v = (acc + c[d]) * k_value[b] / 1000.0f;       // full_net.py:326-327 / 334-336
      if (depths != nullptr) depths[(size_t)b * dn + d] = v;
      if (d == root_index) depth[b] = v;
    }
  }
}

int depth_head_ex_launch(const float* feat, const float* Af, const float* Bz, const float* Wz, const float* bz, const float* c,
                         const float* k_value, float* depth, float* depths, int B, int C, int Z, int dn, int root_index, cudaStream_t s) {
  if (B <= 0) return HRP_OK;
  if (C % 4 || dn < 1 || root_index < 0 || root_index >= dn) return fail(HRP_ERR_INVALID, "depth_head_ex: bad sizes (C=%d dn=%d root=%d)", C, dn, root_index);
  depth_head_ex_kernel<<<B, 256, (size_t)(C + Z) * sizeof(float), s>>>(feat, Af, Bz, Wz, bz, c, k_value, depth, depths, C, Z, dn, root_index);
  HRP_CHECK_LAUNCH("depth_head_ex_kernel");
  return HRP_OK;
}

// The pose / rotation refinement loops (full_net.py:376-394, 430-444) have no nonlinearity (dropout is the identity in eval
// mode), so n iterations of  s <- s + dec(fc2(fc1([xf, s])))  are the affine map  s_n = G_n xf + P_n s_0 + g_n  with
// G_n, P_n, g_n composed once in fp64 at hrp_finalize_weights (network.cu, compose_heads). One launch evaluates every
// iterate of both heads: row r = n * (dof + 6) + j of G is the j-th state component after n + 1 iterations.
// s0_default [dof + 6]: the module's init_pose / init_rot buffers; ovr_pose / ovr_rot [B, dof] / [B, 6]: per-call
// overrides (full_net.py:268-272), used when non-null and (flags == null or flags[k] != 0).
__global__ void __launch_bounds__(256)
heads_affine_kernel(const float* __restrict__ xf, const float* __restrict__ G, const float* __restrict__ P,
                    const float* __restrict__ g, const float* __restrict__ s0_default, const float* ovr_pose,
                    const float* ovr_rot, const int* flags, float* __restrict__ iters, float* __restrict__ pose,
                    float* __restrict__ rot, int F, int dof, int n_iter, int rot_matmul) {
  extern __shared__ float hx[];                       // xf row, then the two initial states, then u (rot_matmul)
  const int b = blockIdx.x, R1 = dof + 6, R = n_iter * R1;
  for (int c = threadIdx.x * 4; c < F; c += 256 * 4)
    *reinterpret_cast<float4*>(hx + c) = __ldg(reinterpret_cast<const float4*>(xf + (size_t)b * F + c));
  float* s0 = hx + F;
  if (threadIdx.x < R1) {
    const int j = threadIdx.x;
    const bool is_rot = j >= dof;
    const float* o = is_rot ? ovr_rot : ovr_pose;
    const bool use = o != nullptr && (flags == nullptr || flags[is_rot ? 1 : 0] != 0);
    s0[j] = use ? o[(size_t)b * (is_rot ? 6 : dof) + (is_rot ? j - dof : j)] : s0_default[j];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pblk = dof * dof + 36;
  float* u = s0 + R1;                                 // rot_matmul: A xf + c, the state-independent part of decrot's output
  for (int r = warp; r < R; r += 8) {
    if (rot_matmul && r >= R1 && (r % R1) >= dof) continue;      // composed rows of later rotation iterates do not exist
    const float4* gr = reinterpret_cast<const float4*>(G + (size_t)r * F);
    float acc = 0.f;
    for (int c = lane; c < F / 4; c += 32) {
      const float4 w = __ldg(gr + c);
      const float4 x = *reinterpret_cast<const float4*>(hx + 4 * c);
      acc = fmaf(w.x, x.x, acc); acc = fmaf(w.y, x.y, acc); acc = fmaf(w.z, x.z, acc); acc = fmaf(w.w, x.w, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const int n = r / R1, j = r - n * R1;
      const bool is_rot = j >= dof;
      const int w = is_rot ? 6 : dof, jj = is_rot ? j - dof : j;
      const float* pr = P + (size_t)n * pblk + (is_rot ? dof * dof : 0) + jj * w;
      const float* st = s0 + (is_rot ? dof : 0);
      float v = acc + g[r];
      if (rot_matmul && is_rot) { u[jj] = v; continue; }
      for (int q = 0; q < w; ++q) v = fmaf(pr[q], st[q], v);
      iters[(size_t)b * R + r] = v;
      if (n == n_iter - 1) {
        if (is_rot) rot[(size_t)b * 6 + jj] = v; else pose[(size_t)b * dof + jj] = v;
      }
    }
  }
  if (rot_matmul) {
    // rot_iterative_matmul (full_net.py:413-429): d = decrot(...) = u + M s is itself a 6-D rotation, COMPOSED with the
    // current one: s <- first two rows of R(d) R(s) (geometries.py:100-132). M = rotation block of P's first iterate.
    __syncthreads();
    if (threadIdx.x == 0) {
      const float* M = P + dof * dof;
      float s[6];
      for (int q = 0; q < 6; ++q) s[q] = s0[dof + q];
      auto rotmat = [](const float* v, float* Rm) {       // rows x, y, z: x = a1/|a1|, z = (x X a2)/|.|, y = z X x
        const float nx = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float x0 = v[0] / nx, x1 = v[1] / nx, x2 = v[2] / nx;
        float z0 = x1 * v[5] - x2 * v[4], z1 = x2 * v[3] - x0 * v[5], z2 = x0 * v[4] - x1 * v[3];
        const float nz = sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
        z0 /= nz; z1 /= nz; z2 /= nz;
        Rm[0] = x0; Rm[1] = x1; Rm[2] = x2;
        Rm[3] = z1 * x2 - z2 * x1; Rm[4] = z2 * x0 - z0 * x2; Rm[5] = z0 * x1 - z1 * x0;
        Rm[6] = z0; Rm[7] = z1; Rm[8] = z2;
      };
      for (int n = 0; n < n_iter; ++n) {
        float d[6], Rd[9], Rs[9];
        for (int j = 0; j < 6; ++j) {
          float v = u[j];
          for (int q = 0; q < 6; ++q) v = fmaf(M[j * 6 + q], s[q], v);
          d[j] = v;
        }
        rotmat(d, Rd); rotmat(s, Rs);
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 3; ++j) s[i * 3 + j] = Rd[i * 3] * Rs[j] + Rd[i * 3 + 1] * Rs[3 + j] + Rd[i * 3 + 2] * Rs[6 + j];
        for (int j = 0; j < 6; ++j) iters[(size_t)b * R + (size_t)n * R1 + dof + j] = s[j];
      }
      for (int j = 0; j < 6; ++j) rot[(size_t)b * 6 + j] = s[j];
    }
  }
}

int heads_affine_launch(const float* xf, const float* G, const float* P, const float* g, const float* s0_default,
                        const float* ovr_pose, const float* ovr_rot, const int* flags, float* iters, float* pose, float* rot,
                        int B, int F, int dof, int n_iter, int rot_matmul, cudaStream_t s) {
  if (B <= 0) return HRP_OK;
  if (F % 4) return fail(HRP_ERR_INVALID, "heads_affine: feature width %d is not a multiple of 4", F);
  heads_affine_kernel<<<B, 256, (size_t)(F + dof + 6 + 6) * sizeof(float), s>>>(xf, G, P, g, s0_default, ovr_pose, ovr_rot, flags, iters, pose,
                                                                                rot, F, dof, n_iter, rot_matmul);
  HRP_CHECK_LAUNCH("heads_affine_kernel");
  return HRP_OK;
}

// ---- host-side packing ------------------------------------------------------------------------------------------------------
void pack_conv_f32(const float* w, const float* conv_bias, const float* bn_w, const float* bn_b, const float* bn_mean,
                   const float* bn_var, int Cout, int Cin, int KH, int KW, float* w_out, float* bias_out) {
  std::vector<double> scale(Cout, 1.0);
  for (int o = 0; o < Cout; ++o) {
    double b = conv_bias ? (double)conv_bias[o] : 0.0;
    if (bn_w) {
      scale[o] = (double)bn_w[o] / std::sqrt((double)bn_var[o] + 1e-5);   // eps: nn.BatchNorm2d default
      b = (double)bn_b[o] + (b - (double)bn_mean[o]) * scale[o];
    }
    bias_out[o] = (float)b;
  }
  for (int r = 0; r < KH; ++r)
    for (int s = 0; s < KW; ++s)
      for (int c = 0; c < Cin; ++c)
        for (int o = 0; o < Cout; ++o)
          w_out[((size_t)(r * KW + s) * Cin + c) * Cout + o] =
              (float)((double)w[(((size_t)o * Cin + c) * KH + r) * KW + s] * scale[o]);
}

}  // namespace hrp
