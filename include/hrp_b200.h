/*
 * hrp_b200.h -- C ABI of libhrp_b200.so: HoliRobPose full-network inference forward on NVIDIA B200 (sm_100a).
 *
 * The reference (Grz684/Holistic-Robot-Pose-Estimation-Study) has no FFI layer: its boundary for this path is the
 * Python call  model(x_reg, x_root, k_value, K)  on an nn.Module (lib/models/full_net.py:262-466, call sites
 * lib/core/function.py:133-141, scripts/test.py:161, scripts/real_test.py:295) followed by
 * point_projection_from_3d_tensor(K, xyz) (lib/utils/transforms.py:17-21). This header is the FFI that boundary binds
 * to when the path is replaced; INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions: plain C, no torch types. Every function returns HRP_OK (0) or a negative hrp_status; the message is in
 * hrp_last_error() (thread-local). Device pointers are caller-owned and must live on the handle's device; work is
 * enqueued asynchronously on the caller's stream and no call synchronises the host (the reference's device->host copy
 * at full_net.py:342 does not exist here). A handle is not re-entrant; use one handle per (device, precision).
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef HRP_B200_H_
#define HRP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HRP_OK = 0,
  HRP_ERR_INVALID = -1,     /* bad argument / unsupported configuration (never a silent fallback) */
  HRP_ERR_CUDA = -2,        /* CUDA runtime/driver error */
  HRP_ERR_STATE = -3,       /* call order violated (e.g. forward before finalize) */
  HRP_ERR_WEIGHT = -4,      /* unknown / missing / mis-shaped tensor */
  HRP_ERR_NOMEM = -5
} hrp_status;

typedef enum { HRP_F32 = 0, HRP_I64 = 1 } hrp_dtype;

/* Arithmetic of the conv / linear contractions. Everything else (softmax, kinematics, heads' epilogues) is fp32. */
typedef enum {
  HRP_PREC_FP32 = 0,        /* fp32 FFMA implicit GEMM: parity mode for every configuration */
  HRP_PREC_TF32 = 1,        /* tcgen05 kind::tf32, operands rounded to nearest TF32, fp32 accumulate in TMEM: parity
                               mode; single-pass on the ResNet-50 keypoint backbone and the DepthNet, 3xTF32 (below) on
                               the layers of an HRNet-W32 keypoint backbone */
  HRP_PREC_BF16 = 2,        /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM */
  HRP_PREC_F16 = 4,         /* tcgen05 kind::f16 with IEEE-half operands (11-bit significand, the same as TF32; conversions
                               saturate at +-65504), fp32 accumulate in TMEM, fp16 NHWC activations: TF32-grade accuracy at
                               the bf16 family's speed and memory traffic */
  HRP_PREC_TF32X3 = 3       /* 3xTF32 on every conv layer: operands split into hi + lo TF32 halves, three tcgen05 products
                               per k-step (Ahi Whi + Alo Whi + Ahi Wlo), fp32-grade results at a third of the TF32 rate.
                               HRP_PREC_TF32 itself uses it on the layers of an HRNet-W32 KEYPOINT backbone */
} hrp_precision;

typedef enum { HRP_BACKBONE_RESNET50 = 0, HRP_BACKBONE_HRNET32 = 1 } hrp_backbone;

/* ---- kinematic program: compiled form of a URDF (see holistic-robot-pose-estimation-study_b200/urdf.py) ----------
 * Replaces URDFRobot.get_keypoints / get_keypoints_root / get_TWL (lib/utils/urdf_robot.py:95-135,193-223) and
 * URDF.link_fk_batch (lib/utils/urdfpytorch/urdf.py:3064-3167). */
#define HRP_FK_MAX_STEPS 32
#define HRP_FK_MAX_KP 32
#define HRP_FK_MAX_SLOTS 8
#define HRP_FK_PARENT_BASE (-1)
#define HRP_FK_PARENT_PREV (-2)

typedef struct {
  int32_t dof;               /* columns of the joint-configuration matrix */
  int32_t nkpt;              /* keypoints per pose */
  int32_t n_steps;           /* movable joints on some keypoint path, chain-contiguous order */
  int32_t n_slots;           /* saved frames needed by branching trees */
  int32_t root_kp;           /* 0: keypoints in the base frame; >0: re-root at that keypoint's link frame */
  int32_t root_step;         /* step whose frame carries the root link (-1: base) */
  const int32_t* step_type;  /* 1 revolute, 2 prismatic */
  const int32_t* step_parent;/* HRP_FK_PARENT_BASE, HRP_FK_PARENT_PREV or a slot index */
  const int32_t* step_save;  /* slot to save this step's frame into, or -1 */
  const int32_t* step_q;     /* configuration column */
  const float* step_mul;     /* mimic: q' = mul*q + off */
  const float* step_off;
  const float* step_origin;  /* n_steps x 12: parent->joint frame (fixed joints folded), row-major 3x4 */
  const float* step_axis;    /* n_steps x 3 */
  const int32_t* kp_step;    /* nkpt, ascending; -1 = base frame */
  const int32_t* kp_index;   /* output keypoint index */
  const float* kp_offset;    /* nkpt x 3, in the step frame */
  const float* root_fixed;   /* 12: step frame -> root link frame */
} hrp_fk_program;

typedef struct hrp_fk hrp_fk;

int hrp_fk_create(const hrp_fk_program* prog, hrp_fk** out);
void hrp_fk_destroy(hrp_fk* fk);

/* Batched FK + pinhole projection, one pose per thread.
 * q [N,dof], rot6d [N,6] (first two ROWS of R, lib/utils/geometries.py:100-132), trans [N,3], Kmat [N,3,3] row-major
 * -> xyz [N,nkpt,3] (metres, camera frame), uv [N,nkpt,2] (pixels; may be NULL). Replaces urdf_robot.py:95-118 /
 * 193-223 + transforms.py:17-21. All fp32 device pointers. */
int hrp_fk_project(const hrp_fk* fk, const float* q, const float* rot6d, const float* trans, const float* Kmat,
                   int64_t N, float* xyz, float* uv, void* stream);

/* ---- heatmap soft-argmax (integral layer) --------------------------------------------------------------------------
 * hm [B, K*D, H, W] fp32 logits in the reference layout (channel = k*D + d, lib/utils/integral.py:122,170);
 * softmax over D*H*W per (b,k), marginal first moments, /size - 0.5, fixroot, then uvd_to_xyz
 * (lib/utils/transforms.py:33-82) with the inverse pinhole of integral.py:56-73.
 * Kmat [B,3,3], root_z [B] absolute root depth in metres. uvd [B,K,3]; xyz [B,K,3] (may be NULL, then Kmat/root_z
 * may be NULL). workspace: hrp_softargmax3d_workspace() bytes of device scratch. */
size_t hrp_softargmax3d_workspace(int B, int K, int D, int H, int W);
int hrp_softargmax3d(const float* hm, int B, int K, int D, int H, int W, const float* Kmat, const float* root_z,
                     float depth_factor, float image_size, int rootid, int fixroot, float* uvd, float* xyz,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- single convolution (layer-level parity tests and micro-benchmarks) ----------------------------------------------
 * in NHWC [B,Hi,Wi,Cin] fp32; weight in the reference layout OIHW [Cout,Cin,KH,KW] fp32 (packed internally, not on a
 * timed path); bias [Cout] or NULL; residual NHWC [B,Ho,Wo,Cout] or NULL; out NHWC [B,Ho,Wo,Cout]. All device. */
int hrp_conv2d_nhwc(const float* in, const float* weight_oihw, const float* bias, const float* residual, float* out,
                    int B, int Hi, int Wi, int Cin, int Cout, int KH, int KW, int stride, int pad, int relu,
                    int precision, void* stream);

/* One HRNet BasicBlock -- relu(conv2(relu(conv1(x) + b1)) + b2 + x), both convs 3x3/s1/p1, C -> C (HRnet.py:28-57, BN
 * already folded by the caller) -- through the fused two-GEMM kernel of the bf16 family. x, out NHWC [B,H,W,C] fp32
 * (operands are rounded to bf16 inside), weights OIHW fp32, biases [C] or NULL. HRP_ERR_INVALID when the block shape is
 * not one the fused kernel takes (it never falls back silently). Layer-level parity tests. */
int hrp_basic_block_nhwc(const float* x, const float* w1_oihw, const float* b1, const float* w2_oihw, const float* b2,
                         float* out, int B, int H, int W, int C, void* stream);

/* A chain of `nblocks` (1..4) such BasicBlocks -- one branch of an HRNet module (HRnet.py:137-149) -- through the one-launch
 * branch kernel of the bf16 family (one image per CTA, activations resident in shared memory, weights streamed). w_oihw:
 * [2*nblocks][C][C][3][3] (conv1, conv2 of block 0, conv1 of block 1, ...), b: [2*nblocks][C] or NULL; x, out NHWC fp32.
 * HRP_ERR_INVALID when the shape is not one the kernel takes (C = 128 / 256 at the low HRNet resolutions). Parity tests. */
int hrp_basic_chain_nhwc(const float* x, const float* w_oihw, const float* b, int nblocks, float* out, int B, int H, int W, int C,
                         void* stream);

/* ---- full network ------------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t backbone;          /* hrp_backbone: keypoint-branch backbone; the DepthNet backbone is always HRNet-W32 */
  int32_t precision;         /* hrp_precision */
  int32_t n_iter;            /* refinement iterations of the pose / rotation heads (shipped: 4) */
  int32_t fix_root;          /* zero the root keypoint's d before back-projection (shipped: 1) */
  float image_size;          /* 256 */
  float depth_factor;        /* bbox_3d_shape[2] * 1e-3 (shipped: 1.3) */
  /* constructor variants outside the shipped configuration; all 0 = the shipped network */
  int32_t direct_reg_rot;        /* rotation from a seven-layer regressor instead of the refinement loop (full_net.py:107-117, 395-409) */
  int32_t rot_iterative_matmul;  /* each refinement step COMPOSES the regressed rotation with the current one (full_net.py:413-429) */
  int32_t add_fc;                /* DepthNet bottleneck MLP in front of depth_layer (full_net.py:156-163, 296-313) */
  int32_t depth_num;             /* multi_kp: len(kps_need_depth) depth outputs (>= 1), all returned in HRP_F_DEPTHS; 0 = single root depth */
  int32_t depth_root;            /* multi_kp: kps_need_depth.index(reference_keypoint_id) (full_net.py:328-329) */
  int32_t reg_joint_map;         /* joint angles from per-joint maps over the trunk's 8x8 grid + a 1-D soft-argmax instead of the
                                    refinement loop (ResNet keypoint backbone only; full_net.py:92-101, 376-379; integral.py:211-251) */
  int32_t joint_conv_dim[3];     /* channels of the three 3x3 conv + BN + ReLU layers in front of it (multiples of 32) */
  float joint_bounds[32];        /* dof x {lower, upper} (lib/dataset/const.py:239-284) */
} hrp_config;

typedef struct hrp_handle hrp_handle;

/* Replaces get_rootNetwithRegInt_model / RootNetwithRegInt.__init__ (full_net.py:37-212, 470-505). */
int hrp_create(const hrp_config* cfg, const hrp_fk_program* robot, int device, hrp_handle** out);
void hrp_destroy(hrp_handle* h);

/* The tensors the handle needs, in the reference's state-dict naming. */
int hrp_num_weights(const hrp_handle* h);
const char* hrp_weight_name(const hrp_handle* h, int i);
int hrp_weight_shape(const hrp_handle* h, int i, int64_t* shape /*[4]*/, int* ndim);

/* Replaces load_state_dict: hand over one HOST tensor (copied; BN folding / repacking happens in finalize).
 * Unknown names -> HRP_ERR_WEIGHT; `num_batches_tracked` tensors are accepted and ignored. */
int hrp_set_weight(hrp_handle* h, const char* name, const void* host_ptr, const int64_t* shape, int ndim, int dtype);
int hrp_finalize_weights(hrp_handle* h);

/* Output record: one flat fp32 device buffer, struct-of-arrays. Field f occupies
 * [offsets[f], offsets[f+1]) floats; offsets has HRP_NUM_FIELDS+1 entries. */
enum {
  HRP_F_POSE = 0,      /* [B,dof]    joint angles                      full_net.py:394 */
  HRP_F_ROT = 1,       /* [B,6]      rot6d                             full_net.py:444 */
  HRP_F_TRANS = 2,     /* [B,3]      root translation                  full_net.py:367 */
  HRP_F_ROOT_UV = 3,   /* [B,2]      root keypoint pixel               full_net.py:360 */
  HRP_F_DEPTH = 4,     /* [B,1]      root depth (m)                    full_net.py:334-336 */
  HRP_F_UVD = 5,       /* [B,nkpt,3]                                   integral.py:146-151 */
  HRP_F_XYZ_INT = 6,   /* [B,nkpt,3] integral keypoints                integral.py:157 */
  HRP_F_XYZ_FK = 7,    /* [B,nkpt,3] forward-kinematics keypoints      full_net.py:447-450 */
  HRP_F_KP2D_INT = 8,  /* [B,nkpt,2] projection of XYZ_INT             transforms.py:17-21 */
  HRP_F_KP2D_FK = 9,   /* [B,nkpt,2] projection of XYZ_FK */
  HRP_F_DEPTHS = 10,   /* [B,depth_num] every regressed depth (m), multi_kp only (width 0 otherwise)   full_net.py:319-327, 462-464 */
  HRP_NUM_FIELDS = 11
};
int hrp_output_offsets(const hrp_handle* h, int B, int64_t* offsets /*[HRP_NUM_FIELDS+1]*/);

/* Device scratch the handle owns for batch B (allocated on first use, outside any timed region after warm-up). */
size_t hrp_workspace_bytes(hrp_handle* h, int B);

/* The forward: x_reg, x_root NCHW fp32 [B,3,256,256] in [0,1]; k_value [B]; Kmat [B,3,3]; out: flat record.
 * Replaces RootNetwithRegInt.forward (full_net.py:262-466) + both caller-side projections (function.py:140-141). */
int hrp_forward(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                int B, float* out, void* stream);

/* The same forward with per-frame initial states for the two refinement heads -- the reference's `init_pose` /
 * `init_rot` keyword arguments (full_net.py:262, 268-272): init_pose [B,dof], init_rot [B,6] device fp32, either may be
 * NULL (then the module's `init_pose` / `init_rot` buffers are used, as in the reference). */
int hrp_forward_ex(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                   const float* init_pose, const float* init_rot, int B, float* out, void* stream);

/* The reference's `test_fps=True` mode (full_net.py:277-279, 337-345, 452-460; called by scripts/test.py:161-162): the
 * forward runs un-graphed in list order (DepthNet first) and SYNCHRONISES the stream, like the reference does, then
 * reports ms3 = {time_root, time_other, time_whole} in MILLISECONDS of device time (CUDA events: start -> depth head done
 * -> end of FK). init_pose / init_rot as in hrp_forward_ex. */
int hrp_forward_timed(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value, const float* Kmat,
                      const float* init_pose, const float* init_rot, int B, float* out, float* ms3 /*[3]*/, void* stream);

/* The forward fed with what the reference's DataLoader hands to the device: uint8 NCHW crops [B,3,256,256]
 * (`input_batch["root"]["images"]`, lib/dataset/dream.py:441-443); the `.float() / 255.` of scripts/test.py:93-96 /
 * lib/core/function.py happens on the device inside the stem's input pack. A quarter of the host->device bytes of
 * hrp_forward; results are bit-identical to hrp_forward on x = u8 / 255. */
int hrp_forward_u8(hrp_handle* h, const uint8_t* x_reg, const uint8_t* x_root, const float* k_value, const float* Kmat,
                   int B, float* out, void* stream);

/* Input side of the boundary on the device, replacing the reference's CPU data preparation for
 * inference: frames [B,Hf,Wf,3] uint8 HWC camera images, crop_box [B,4] int32 (wmin,hmin,wmax,hmax) inside the frame ->
 * crops [B,3,256,256] uint8 NCHW exactly as the DataLoader builds them: the box pasted into a zero square
 * (lib/dataset/roboutils.py:142-171), bilinear resize with align_corners=False on /255 floats, truncated back to uint8
 * (lib/dataset/augmentations.py:189-262); K_in [B,3,3] -> K_out through the paste shift + get_K_crop_resize
 * (lib/utils/geometries.py:360-402); and, when k_value != NULL, k_box [B,4] float (the strict robot box in FRAME
 * coordinates) -> bbox_transform + clipping (lib/dataset/dream.py:445-449, roboutils.py:248-263) ->
 * k_value[b] = sqrt(fx fy 1000^2 / max(|dx|,|dy|)^2) with the new fx, fy (lib/core/function.py:98-110). All device. */
int hrp_crop_resize_u8(const uint8_t* frames, int B, int Hf, int Wf, const int32_t* crop_box, const float* k_box,
                       const float* K_in, uint8_t* crops, float* K_out, float* k_value, void* stream);

/* Evaluation tail on the device, replacing compute_metrics_batch (lib/utils/metrics.py:8-118; callers
 * lib/core/function.py:158-172, scripts/test.py:167-181). pred_xyz [B,nkpt,3] / pred_uv [B,nkpt,2]: the predicted pose through
 * hrp_fk_project with the ORIGINAL camera matrix (metrics.py:29-42) -- or pred_uv NULL and K_original [B,3,3] given: the points are
 * projected here (metrics.py:22-26, 42: predictions that are 3-D points, not a pose) --; pred_joint [B,dof] or NULL (metrics.py:90-92: zeros);
 * gt_xyz, gt_uv, gt_joint: ground truth of the batch. joint_cols: columns averaged per frame (dof, Panda dof-1:
 * metrics.py:87-88). Outputs (device): per_frame [B,6] = error3d, error2d (mean over keypoints with gt inside the 640x480 frame;
 * NaN when none is), mean_jointerror, error_depth, batch_error_relative, error3d_relative; dis3d [nkpt], dis2d [nkpt],
 * l1_joint [dof] = the batch means per keypoint / joint. */
int hrp_metrics_batch(const float* pred_xyz, const float* pred_uv, const float* K_original, const float* pred_joint, const float* gt_xyz,
                      const float* gt_uv, const float* gt_joint, int B, int nkpt, int dof, int root_kp, int joint_cols,
                      float* per_frame, float* dis3d, float* dis2d, float* l1_joint, void* stream);

/* summary_add_pck (lib/utils/metrics.py:121-162; scripts/test.py:232-233) over the per-frame error lists of a whole run, kept on
 * the device: dis3d [N] metres, dis2d [N] pixels -> out22 (device, fp64): for dis3d then dis2d, {mean, median, AUC (trapezoid of
 * the fraction under np.arange(0, 0.1, 1e-5) metres / np.arange(0, 20, 0.01) pixels, over the range), fraction <= each of the
 * eight table thresholds (1,5,10,20,40,60,80,100 mm / 2.5,...,20 px)}. workspace: hrp_summary_workspace(N) bytes. */
size_t hrp_summary_workspace(int64_t N);
int hrp_summary_add_pck(const float* dis3d, const float* dis2d, int64_t N, double* out22, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- output gather over peer memory (one box, one process per GPU) ---------------------------------------------
 * The path's only exchange step: every rank's packed output record to every rank (the reference's counterpart is
 * nn.DataParallel's gather, scripts/test.py:159). Each rank owns a window in its HBM that the peers map through CUDA IPC; an
 * all-gather is two kernels on the caller's stream that store the record straight into the peers' windows over NVLink and
 * copy the arrived records out -- no NCCL call, no host synchronisation (csrc/p2p_gather.cu).
 * hrp_p2p_create (allocates the window for records of bytes_per_rank, rounded up to 16) -> hrp_p2p_handle (64 bytes to hand to
 * every peer, e.g. with one torch.distributed all_gather at start-up) -> hrp_p2p_connect (all handles, rank order) ->
 * hrp_p2p_all_gather (src: this rank's record, dst: world x bytes, both device, 16-byte aligned; every rank must call it the
 * same number of times in the same order) -> hrp_p2p_status (synchronises; HRP_ERR_STATE if a peer timed out). */
typedef struct hrp_p2p hrp_p2p;
int hrp_p2p_create(int rank, int world, size_t bytes_per_rank, int device, hrp_p2p** out);
int hrp_p2p_handle(hrp_p2p* g, void* handle64);
int hrp_p2p_connect(hrp_p2p* g, const void* handles);
int hrp_p2p_all_gather(hrp_p2p* g, const void* src, size_t bytes, void* dst, void* stream);
int hrp_p2p_status(hrp_p2p* g);
void hrp_p2p_destroy(hrp_p2p* g);

/* Plans (workspace + CUDA graph) are cached per batch size: at most "max_cached_batches" distinct sizes (default 4, least
 * recently used dropped first), and a second / third plan of a size only when forwards of that size arrive on different
 * streams. hrp_release_plans frees every cached plan now (waits for the forwards that use them); the next forward
 * re-plans. */
int hrp_release_plans(hrp_handle* h);

/* Options: "cuda_graph" (0/1, default 1: replay one graph per batch size), "lanes" (0/1, default 1: capture the graph
 * over several streams so independent sub-networks overlap), "slots" (1..8, default 4: plans = workspace + graph kept per
 * batch size and used round-robin, so consecutive forwards enqueued on different streams overlap), "lane_share_pct"
 * (5..100: share of the CTA slots one conv launch may take; default 25 for batches of 32 frames and more -- it maximises
 * the throughput of overlapping forwards -- and 50 below, where the latency of a single forward matters). Graph-shaping options take effect for graphs not yet captured.
 * "max_cached_batches" (1..64, default 4): see hrp_release_plans. Unknown option -> HRP_ERR_INVALID. */
int hrp_set_option(hrp_handle* h, const char* name, int64_t value);

/* Introspection for the bench / tests. */
int64_t hrp_launch_count(const hrp_handle* h);            /* kernels one forward enqueues (plan of the most recent hrp_forward) */
/* name: "xf", "img_feat" [B,2048]; "logits" (fp32 family) [B,nkpt*64,64,64]; "head_iters" [B,n_iter,dof+6]: every iterate of
 * the pose (first dof columns) and rot6d refinement (full_net.py:381-394, 431-444). */
int hrp_debug_tensor(hrp_handle* h, const char* name, int B, float* dst_device, int64_t* numel, void* stream);
/* Per-kernel-class device time of one un-graphed forward (CUDA events around every launch; for profiling only).
 * classes: 0 conv-tensor, 1 conv-fp32, 2 stem, 3 pool/fuse elementwise, 4 heads, 5 softargmax, 6 fk. */
#define HRP_NUM_CLASSES 7
int hrp_forward_profile(hrp_handle* h, const float* x_reg, const float* x_root, const float* k_value,
                        const float* Kmat, int B, float* out, float* ms_by_class, int64_t* launches_by_class,
                        double* flops_by_class, void* stream);

/* Micro-benchmark of one tensor-core conv layer (tf32 / bf16 families) on random operands: average device time of
 * `iters` back-to-back launches. with_residual: 0 / 1, or 8 = time the 8-conv branch chain kernel (hrp_basic_chain_nhwc's)
 * on this shape (bf16, k 3, stride 1, Cin == Cout). Tuning aid (scripts/conv_bench.py); not used on the forward path. */
int hrp_conv_bench(int precision, int B, int H, int W, int Cin, int Cout, int k, int stride, int with_residual,
                   int iters, float* ms_per_launch, void* stream);

const char* hrp_last_error(void);
const char* hrp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HRP_B200_H_ */
