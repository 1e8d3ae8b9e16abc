// Internal launch interface between the network executor (network.cu) and the kernels.
#pragma once
#include "common.h"

namespace hrp {

// ---- convolution as implicit GEMM over NHWC activations ------------------------------------------------------------
// M = B*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin. The same descriptor serves plain convs, the four sub-pixel
// phases of a stride-2 transposed conv (out_s* / out_o* place the phase's pixels on the full-resolution grid) and
// nn.Linear (H = W = 1).
struct ConvArgs {
  const void* in;        // NHWC [B,Hi,Wi,Cin]
  const void* w;         // packed weights (layout depends on the kernel family)
  const float* bias;     // [Cout] fp32, BN folded (never null)
  const void* res;       // residual, same layout/shape as out, or null
  void* out;             // NHWC [B,Ho_full,Wo_full,Cout]  (or NCHW fp32 when out_nchw)
  int B, Hi, Wi, Cin, Ho, Wo, Cout;
  int KH, KW, stride, pad_h, pad_w;
  int out_sy, out_sx, out_oy, out_ox, Ho_full, Wo_full;
  int relu;
  int res_after_act;   // 0: act(acc + bias + res)   1: act(acc + bias) + res   (HRnet.py:539-540)
  int out_nchw;
  int ld_out;            // channel stride of one output pixel (>= Cout; lets two GEMMs share one row)
  int out_coff;          // first output channel inside that row
  // Tensor-core family only: explicit TMA view of `in` (element units / byte strides) replacing the default NHWC one.
  // Used by the stems, whose K axis is a window of a padded 4-channel image (see stem_pack_launch).
  int tma_custom;
  unsigned long long tm_gdim[4], tm_gstr[3];
  unsigned tm_box[4];
  // Tensor-core family only: percentage of the GPU's CTA slots this launch's persistent grid may take (0 = all). The
  // multi-lane graph runs several convs at once; most are latency-bound, so sharing the SMs beats time-slicing them.
  int grid_pct;
  // Tensor-core family only: the conv IS the heatmap head (Cout = keypoints x 64 depth bins over a 64x64 map). Instead of
  // storing logits, each 128-pixel x 64-bin tile reduces its own online-softmax state and writes one 5-float partial to
  // sa_partial[(frame*keypoints + keypoint) * (Ho*Wo/128) + tile] for softargmax_finalize_launch. `out` is not written.
  float* sa_partial;
  // TF32 family only: 3xTF32 (conv_tc.cu). `w` is the pack_conv_tc(tf32 = 2) image, `in` / `res` / `out` are full fp32.
  int x3;
  // 2-byte families only: the operands are IEEE half (f16 family) instead of bf16; `w` is the pack_conv_tc(tf32 = 3) image.
  int f16;
};

int conv_f32_launch(const ConvArgs& a, cudaStream_t s);

// Tensor-core family (conv_tc.cu): tcgen05.mma with bf16 (tf32 = 0) or TF32 (tf32 = 1) operands, fp32 accumulation in
// TMEM. Activations are NHWC bf16 / fp32; `a.w` is the pack_conv_tc() image. round_tf32: round fp32 NHWC outputs to
// TF32 (nearest) for the next layer's operands.
bool conv_tc_supported(const ConvArgs& a, int tf32);
int conv_tc_launch(const ConvArgs& a, int tf32, int round_tf32, cudaStream_t s);
// Shifted-GEMM kernel for 3x3 / stride 1 / pad 1 convs with Cin == Cout (conv_slab.cu); same operands and weight image
// as conv_tc. conv_tc_launch dispatches to it when it applies.
bool conv_slab_supported(const ConvArgs& a, int tf32);
int conv_slab_launch(const ConvArgs& a, int tf32, int round_tf32, cudaStream_t s);
// Operand-row width (64 / 128 bytes) the layer described by `a` (shape fields only) must be packed for; *use_tma tells
// whether its activations will arrive by TMA tensor copies or by the cp.async gather.
int conv_tc_row_bytes(const ConvArgs& a, int tf32, int* use_tma);
size_t pack_conv_tc_bytes(int K, int Cout, int tf32, int row_bytes);
void pack_conv_tc(const float* w_kn, int K, int Cout, int tf32, int row_bytes, void* out);   // from pack_conv_f32's [K][Cout]; tf32 = 2: 3xTF32 image, 3: IEEE half
// One whole BasicBlock (two 3x3/s1/p1 convs, Cin == Cout == C, BN folded, ReLU, identity residual) in one launch
// (conv_block.cu): bf16 NHWC in/out, w1/w2 = pack_conv_tc images (64-byte operand rows), b1/b2 folded biases.
struct BlockArgs {
  const void* x;
  const void *w1, *w2;
  const float *b1, *b2;
  void* out;
  int B, H, W, C;
  int grid_pct;          // as ConvArgs::grid_pct
  int f16;               // operands are IEEE half instead of bf16
};
bool conv_block_supported(const BlockArgs& a);
int conv_block_launch(const BlockArgs& a, cudaStream_t s);
// A chain of BasicBlocks (nconv = 2 x blocks <= 8 convs) on one low-resolution branch in one launch, one image per CTA,
// activations resident in shared memory, weights streamed (conv_chain.cu): bf16 NHWC in/out, C = 128 or 256,
// w[j] = pack_conv_tc images (128-byte operand rows), b[j] folded biases; conv 2k, 2k+1 form block k.
struct ChainArgs {
  const void* x;
  void* out;
  const void* w[8];
  const float* b[8];
  int nconv;
  int B, H, W, C;
  int f16;               // operands are IEEE half instead of bf16
};
bool conv_chain_supported(const ChainArgs& a);
int conv_chain_launch(const ChainArgs& a, cudaStream_t s);
// The same for the 64-channel branch (32x32 map), whose two ping-pong buffers would not fit: one shared-memory buffer, every
// conv writes its output in place a few rows further up (conv_roll.cu). w[j] = pack_conv_tc images with 128-byte rows;
// scratch0 / scratch1: [B,H,W,64] bf16 buffers for the outputs of the inner blocks.
bool conv_roll_supported(const ChainArgs& a);
int conv_roll_launch(const ChainArgs& a, void* scratch0, void* scratch1, cudaStream_t s);

int cast_f32_to_bf16_launch(const float* in, void* out, size_t n, cudaStream_t s, int f16 = 0);   // f16: IEEE half instead of bf16
int cast_bf16_to_f32_launch(const void* in, float* out, size_t n, cudaStream_t s, int f16 = 0);
int round_tf32_launch(const float* in, float* out, size_t n, cudaStream_t s);

// Stem: NCHW fp32 [B,3,Hi,Wi] -> NHWC [B,Ho,Wo,64], KxK stride 2, BN folded, ReLU. w packed [(c*KH+r)*KW+s][64].
// out_mode: 0 fp32, 1 bf16, 2 fp32 rounded to TF32.
int stem_conv_launch(const float* in_nchw, const float* w, const float* bias, void* out, int B, int Hi, int Wi,
                     int Ho, int Wo, int KH, int KW, int pad, int out_mode, cudaStream_t s);

// Stem, tensor-core families: NCHW fp32 [B,3,256,256] -> zero-padded NHWC4 [B, STEM_HP, STEM_WP, 4] (bf16, or fp32
// rounded to TF32), image pixel (y,x) at padded (y+3, x+3), channel 3 = 0. A KxK/2 stem conv then is an implicit GEMM
// whose K axis per kernel row is the contiguous 8-pixel x 4-channel window starting at padded x = 2*ox + 3 - pad:
// 32 elements = one 64-byte (bf16) / 128-byte (TF32) operand row, fetched by TMA through a view whose "ox" dimension
// has a 2-pixel byte stride (overlapping windows). Taps s >= K and channel 3 carry zero weights.
constexpr int STEM_PAD = 3, STEM_HP = 256 + 2 * STEM_PAD, STEM_WP = 264;
int stem_pack_launch(const float* in_nchw, void* out, int B, int tf32 /*0 bf16, 1 fp32 rounded to TF32, 2 fp32 as is (3xTF32), 3 IEEE half*/, cudaStream_t s);

// Same from uint8 NCHW images with the `/ 255.` of scripts/test.py:93-96 folded in (preprocess.cu); u8 -> fp32 NCHW for the
// fp32 family's direct stem.
int stem_pack_u8_launch(const uint8_t* in_nchw, void* out, int B, int mode, cudaStream_t s);
int u8_to_f32_launch(const uint8_t* in, float* out, size_t n, cudaStream_t s);

// `bf16` below: element type of the activations, 0 fp32, 1 bf16, 2 IEEE half
int maxpool3x3s2_launch(const void* in, void* out, int B, int Hi, int Wi, int C, int bf16, cudaStream_t s);

// out = relu( sum_i same[i] + sum_j nearest_up(low[j], 2^shift[j]) ), NHWC, up to 4 + 3 terms (HRNet fusion).
struct FuseArgs {
  const void* same[4];
  const void* low[3];
  int shift[3];
  int n_same, n_low;
  void* out;
  int B, H, W, C;
  int relu;
  int round_tf32;
};
int fuse_sum_launch(const FuseArgs& a, int bf16, cudaStream_t s);

// NHWC [B,HW,C] -> fp32 [B,C] mean over HW.
int avgpool_launch(const void* in, float* out, int B, int HW, int C, int bf16, cudaStream_t s);

// ---- heads ---------------------------------------------------------------------------------------------------------------
// depth[b] = (dot(feat[b], w) + bias) * k_value[b] / 1000          full_net.py:312-336
int depth_head_launch(const float* feat, const float* w, const float* bias, const float* k_value, float* depth, int B,
                      int C, cudaStream_t s);
// Every iterate of both refinement heads as one affine map per iterate (conv_f32.cu, heads_affine_kernel):
// iters [B, n_iter, dof+6]; pose [B,dof] / rot [B,6] receive the last iterate. G [n_iter*(dof+6), F], P per iterate
// {dof x dof, 6 x 6}, g [n_iter*(dof+6)], s0_default [dof+6]; ovr_* optional per-frame initial states.
int joint_map_head_launch(const void* y, const float* w, const float* bias, const float* bounds, float* pose, int B, int HW, int C,
                          int dof, int bf16, cudaStream_t s);
int depth_head_ex_launch(const float* feat, const float* Af, const float* Bz, const float* Wz, const float* bz, const float* c,
                         const float* k_value, float* depth, float* depths, int B, int C, int Z, int dn, int root_index, cudaStream_t s);
int heads_affine_launch(const float* xf, const float* G, const float* P, const float* g, const float* s0_default,
                        const float* ovr_pose, const float* ovr_rot, const int* flags, float* iters, float* pose, float* rot,
                        int B, int F, int dof, int n_iter, int rot_matmul, cudaStream_t s);

// ---- integral layer / kinematics (softargmax.cu, fk_project.cu) ----------------------------------------------------------
size_t softargmax_workspace(int B, int K, int D, int H, int W);
int softargmax_launch(const float* hm, int B, int K, int D, int H, int W, const float* Kmat, const float* root_z,
                      float depth_factor, float image_size, int rootid, int fixroot, float* uvd, float* xyz, void* ws,
                      size_t ws_bytes, float* root_uv, float* trans, float* kp2d, cudaStream_t stream, int* launches);
int softargmax_finalize_launch(const float* partial, int B, int K, int chunks, int D, int H, int W, const float* Kmat,
                               const float* root_z, float depth_factor, float image_size, int rootid, int fixroot,
                               float* uvd, float* xyz, float* root_uv, float* trans, float* kp2d, cudaStream_t stream);
int fk_launch(const hrp_fk* fk, const float* q, const float* rot6d, const float* trans, const float* Kmat, int64_t N,
              float* xyz, float* uv, cudaStream_t stream);

// ---- host-side weight packing (fp32 family) --------------------------------------------------------------------------------
// OIHW (+conv bias, +BN) -> [(r*KW+s)*Cin + c][Cout] with the BN scale folded in; bias_out[Cout].
void pack_conv_f32(const float* w_oihw, const float* conv_bias, const float* bn_w, const float* bn_b,
                   const float* bn_mean, const float* bn_var, int Cout, int Cin, int KH, int KW, float* w_out,
                   float* bias_out);

}  // namespace hrp
