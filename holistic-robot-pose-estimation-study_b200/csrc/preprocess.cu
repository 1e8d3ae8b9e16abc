// Input side of the boundary: what the reference does on the CPU before `model(...)` is called.
//
//  * crop_resize_u8: raw camera frame (uint8 HWC) + integer crop box -> the 256x256 network crop, as the reference's
//    DataLoader builds it: `resize_image` pastes the box into a zero square (lib/dataset/roboutils.py:142-171),
//    `CropResizeToAspectAugmentation` scales it with F.interpolate(bilinear, align_corners=False) on /255 floats and
//    truncates back to uint8 (lib/dataset/augmentations.py:189-262), the camera follows through `get_K_crop_resize`
//    (lib/utils/geometries.py:360-402), the strict box through `bbox_transform` + clipping (lib/dataset/dream.py:445-449,
//    roboutils.py:248-263) and k_value = sqrt(fx fy 1000^2 / max(|dx|,|dy|)^2) (lib/core/function.py:98-110,
//    scripts/test.py:143-153). Output: uint8 NCHW crops (what the DataLoader hands to the device), K', k_value.
//  * u8 NCHW -> fp32 NCHW with the `/ 255.` of scripts/test.py:93-96 (fp32 family), and the tensor-core families' packed
//    stem image straight from uint8 (conv_f32.cu stem_pack_kernel's layout), so a forward fed with uint8 crops uploads a
//    quarter of the bytes and never materialises fp32 images.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace hrp {
namespace {

// one thread per frame; fp32 op-for-op as the torch expressions of get_K_crop_resize, float64 for bbox_transform
__global__ void prep_camera_kernel(const int* __restrict__ crop_box, const float* __restrict__ k_box, const float* __restrict__ K_in,
                                   float* __restrict__ K_out, float* __restrict__ k_value, int B, int out_size) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int wmin = crop_box[4 * b], hmin = crop_box[4 * b + 1], wmax = crop_box[4 * b + 2], hmax = crop_box[4 * b + 3];
  const int S = max(max(wmax - wmin, hmax - hmin), 1);               // roboutils.py:145 (a degenerate box yields a zero crop)
  const int x_off = (S - (wmax - wmin)) / 2, y_off = (S - (hmax - hmin)) / 2;
  const float* Ki = K_in + 9 * b;
  // roboutils.py:167-169 (numpy float64, then .float() in get_K_crop_resize)
  const float k02 = (float)((double)Ki[2] - (double)(wmin - x_off)), k12 = (float)((double)Ki[5] - (double)(hmin - y_off));
  const float final_w = (float)out_size, final_h = (float)out_size, cw = (float)S, ch = (float)S;
  const float cj = __fdiv_rn(__fadd_rn(0.f, (float)S), 2.f), ci = cj;                                  // box = [0, 0, S, S]
  float cx = __fsub_rn(__fadd_rn(k02, __fdiv_rn(__fsub_rn(cw, 1.f), 2.f)), cj);
  float cy = __fsub_rn(__fadd_rn(k12, __fdiv_rn(__fsub_rn(ch, 1.f), 2.f)), ci);
  const float center_x = __fdiv_rn(__fsub_rn(cw, 1.f), 2.f), center_y = __fdiv_rn(__fsub_rn(ch, 1.f), 2.f);
  const float dcx = __fsub_rn(cx, center_x), dcy = __fsub_rn(cy, center_y);
  const float sx = __fdiv_rn(final_w, cw), sy = __fdiv_rn(final_h, ch);
  const float scx = __fdiv_rn(__fsub_rn(final_w, 1.f), 2.f), scy = __fdiv_rn(__fsub_rn(final_h, 1.f), 2.f);
  const float fx = __fmul_rn(sx, Ki[0]), fy = __fmul_rn(sy, Ki[4]);
  cx = __fadd_rn(scx, __fmul_rn(sx, dcx));
  cy = __fadd_rn(scy, __fmul_rn(sy, dcy));
  float* Ko = K_out + 9 * b;
  for (int i = 0; i < 9; ++i) Ko[i] = Ki[i];
  if (S == out_size) {             // the crop already has the target size: the reference returns before any resize (augmentations.py:193-195)
    Ko[2] = k02; Ko[5] = k12;
    cx = k02; cy = k12;
  } else {
    Ko[0] = fx; Ko[4] = fy; Ko[2] = cx; Ko[5] = cy;
  }
  const float fxn = Ko[0], fyn = Ko[4];
  if (k_value == nullptr) return;
  // bbox_transform (roboutils.py:248-263): corners -> K_original^-1 -> K' (float64), clipped to the crop
  const double ifx = 1.0 / (double)Ki[0], ify = 1.0 / (double)Ki[4];
  auto tx = [&](double x, double y) {          // first row of K' K^-1 [x, y, 1]; K has no skew in this pipeline (geometries.py:363)
    const double X = (x - (double)Ki[1] * ify * (y - (double)Ki[5]) - (double)Ki[2]) * ifx, Y = (y - (double)Ki[5]) * ify;
    return (double)fxn * X + (double)Ko[1] * Y + (double)cx;
  };
  auto ty = [&](double y) { return (double)fyn * ((y - (double)Ki[5]) * ify) + (double)cy; };
  const double bx0 = k_box[4 * b], by0 = k_box[4 * b + 1], bx1 = k_box[4 * b + 2], by1 = k_box[4 * b + 3];
  const double lim = (double)out_size;
  auto clip = [&](double v) { return fmin(fmax(v, 0.0), lim); };
  const float t0 = (float)clip(tx(bx0, by0)), t1 = (float)clip(ty(by0)), t2 = (float)clip(tx(bx1, by0)), t3 = (float)clip(ty(by1));
  const float side = fmaxf(fabsf(__fsub_rn(t2, t0)), fabsf(__fsub_rn(t3, t1)));
  const float area = __fmul_rn(side, side);
  k_value[b] = __fsqrt_rn(__fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(fxn, fyn), 1000.f), 1000.f), area));   // function.py:107-110
}

// one thread per output pixel: the three channels of crop pixel (oy, ox) of frame b
__global__ void __launch_bounds__(256)
crop_resize_u8_kernel(const uint8_t* __restrict__ frames, int Hf, int Wf, const int* __restrict__ crop_box, uint8_t* __restrict__ out,
                      int B, int out_size) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * out_size * out_size;
  if (i >= total) return;
  const int ox = (int)(i % out_size), oy = (int)((i / out_size) % out_size), b = (int)(i / ((long long)out_size * out_size));
  const int wmin = crop_box[4 * b], hmin = crop_box[4 * b + 1], wmax = crop_box[4 * b + 2], hmax = crop_box[4 * b + 3];
  const int S = max(wmax - wmin, hmax - hmin);
  if (S <= 0 || wmax < wmin || hmax < hmin) {                       // degenerate box: an all-zero crop, never a fault
    for (int c = 0; c < 3; ++c) out[(((size_t)b * 3 + c) * out_size + oy) * out_size + ox] = 0;
    return;
  }
  const int x_off = (S - (wmax - wmin)) / 2, y_off = (S - (hmax - hmin)) / 2;
  const uint8_t* f = frames + (size_t)b * Hf * Wf * 3;
  if (S == out_size) {                                              // target size already: pixels are copied, not resampled (augmentations.py:193-195)
    const int yy = oy - y_off, xx = ox - x_off, fy = hmin + yy, fx = wmin + xx;
    const bool in = yy >= 0 && yy < hmax - hmin && xx >= 0 && xx < wmax - wmin && fy >= 0 && fy < Hf && fx >= 0 && fx < Wf;
    for (int c = 0; c < 3; ++c) out[(((size_t)b * 3 + c) * out_size + oy) * out_size + ox] = in ? f[((size_t)fy * Wf + fx) * 3 + c] : 0;
    return;
  }
  // ATen area_pixel_compute_scale / _source_index, align_corners = False, no scale factor
  const float scale = __fdiv_rn((float)S, (float)out_size);
  auto src = [&](int d, int* i0, int* i1, float* l0, float* l1) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)d, 0.5f)), 0.5f);
    if (s < 0.f) s = 0.f;
    *i0 = (int)s;
    *i1 = *i0 + (*i0 < S - 1 ? 1 : 0);
    *l1 = fminf(fmaxf(__fsub_rn(s, (float)*i0), 0.f), 1.f);
    *l0 = __fsub_rn(1.f, *l1);
  };
  int x0, x1, y0, y1;
  float wx0, wx1, wy0, wy1;
  src(ox, &x0, &x1, &wx0, &wx1);
  src(oy, &y0, &y1, &wy0, &wy1);
  auto px = [&](int sy, int sx, int c) -> float {          // the zero square with the box pasted in (roboutils.py:146-153)
    const int yy = sy - y_off, xx = sx - x_off;
    if (yy < 0 || yy >= hmax - hmin || xx < 0 || xx >= wmax - wmin) return 0.f;
    const int fy = hmin + yy, fx = wmin + xx;
    if (fy < 0 || fy >= Hf || fx < 0 || fx >= Wf) return 0.f;
    return __fdiv_rn((float)f[((size_t)fy * Wf + fx) * 3 + c], 255.f);                   // augmentations.py:198
  };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float r0 = __fadd_rn(__fmul_rn(px(y0, x0, c), wx0), __fmul_rn(px(y0, x1, c), wx1));
    const float r1 = __fadd_rn(__fmul_rn(px(y1, x0, c), wx0), __fmul_rn(px(y1, x1, c), wx1));
    const float v = __fadd_rn(__fmul_rn(r0, wy0), __fmul_rn(r1, wy1));
    const float q = __fmul_rn(v, 255.f);                                                    // augmentations.py:258: (* 255).to(uint8) truncates
    out[(((size_t)b * 3 + c) * out_size + oy) * out_size + ox] = (uint8_t)(int)q;
  }
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uchar4 v = reinterpret_cast<const uchar4*>(in)[i];
  reinterpret_cast<float4*>(out)[i] = make_float4(__fdiv_rn((float)v.x, 255.f), __fdiv_rn((float)v.y, 255.f), __fdiv_rn((float)v.z, 255.f),
                                                  __fdiv_rn((float)v.w, 255.f));           // scripts/test.py:93-96
}

// uint8 NCHW image -> zero-padded NHWC4 operand image of the tensor-core stems (the layout of conv_f32.cu's
// stem_pack_kernel, kernels.h), with the `/ 255.` folded in. MODE: 0 bf16, 1 fp32 rounded to TF32, 2 fp32 as is, 3 IEEE half.
template <int MODE>
__global__ void stem_pack_u8_kernel(const uint8_t* __restrict__ in, void* __restrict__ out, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * STEM_HP * STEM_WP;
  if (i >= total) return;
  const int xp = (int)(i % STEM_WP);
  const int yp = (int)((i / STEM_WP) % STEM_HP);
  const int b = (int)(i / ((long long)STEM_WP * STEM_HP));
  const int x = xp - STEM_PAD, y = yp - STEM_PAD;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (x >= 0 && x < 256 && y >= 0 && y < 256) {
    const uint8_t* ip = in + ((size_t)b * 3 * 256 + y) * 256 + x;
    v0 = __fdiv_rn((float)__ldg(ip), 255.f); v1 = __fdiv_rn((float)__ldg(ip + 256 * 256), 255.f); v2 = __fdiv_rn((float)__ldg(ip + 2 * 256 * 256), 255.f);
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
    const __half2 a = __floats2half2_rn(v0, v1), c = __floats2half2_rn(v2, 0.f);
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

}  // namespace

int stem_pack_u8_launch(const uint8_t* in_nchw, void* out, int B, int mode, cudaStream_t s) {
  const long long total = (long long)B * STEM_HP * STEM_WP;
  if (total <= 0) return HRP_OK;
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (mode == 3) stem_pack_u8_kernel<3><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else if (mode == 2) stem_pack_u8_kernel<2><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else if (mode == 1) stem_pack_u8_kernel<1><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  else stem_pack_u8_kernel<0><<<blocks, 256, 0, s>>>(in_nchw, out, B);
  HRP_CHECK_LAUNCH("stem_pack_u8_kernel");
  return HRP_OK;
}

int u8_to_f32_launch(const uint8_t* in, float* out, size_t n, cudaStream_t s) {
  if (n == 0) return HRP_OK;
  if (n % 4) return fail(HRP_ERR_INVALID, "u8_to_f32: element count %zu is not a multiple of 4", n);
  u8_to_f32_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(in, out, n / 4);
  HRP_CHECK_LAUNCH("u8_to_f32_kernel");
  return HRP_OK;
}

}  // namespace hrp

using namespace hrp;

extern "C" int hrp_crop_resize_u8(const uint8_t* frames, int B, int Hf, int Wf, const int32_t* crop_box, const float* k_box,
                                  const float* K_in, uint8_t* crops, float* K_out, float* k_value, void* stream) {
  if (B < 0 || Hf <= 0 || Wf <= 0) return fail(HRP_ERR_INVALID, "hrp_crop_resize_u8: bad sizes B=%d frame %dx%d", B, Hf, Wf);
  if (B == 0) return HRP_OK;
  if (!frames || !crop_box || !K_in || !crops || !K_out) return fail(HRP_ERR_INVALID, "hrp_crop_resize_u8: null pointer");
  if (k_value && !k_box) return fail(HRP_ERR_INVALID, "hrp_crop_resize_u8: k_value needs the strict box");
  cudaStream_t st = (cudaStream_t)stream;
  prep_camera_kernel<<<ceil_div(B, 128), 128, 0, st>>>(crop_box, k_box, K_in, K_out, k_value, B, 256);
  HRP_CHECK_LAUNCH("prep_camera_kernel");
  const long long total = (long long)B * 256 * 256;
  crop_resize_u8_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(frames, Hf, Wf, crop_box, crops, B, 256);
  HRP_CHECK_LAUNCH("crop_resize_u8_kernel");
  return HRP_OK;
}
