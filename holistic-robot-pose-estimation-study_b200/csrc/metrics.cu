// Evaluation tail on the device: the per-batch error measures and the end-of-run ADD / PCK summary the reference computes in
// numpy after copying every prediction to the host (lib/utils/metrics.py:8-118 compute_metrics_batch, 121-162
// summary_add_pck; callers lib/core/function.py:158-172, scripts/test.py:167-181, 229-262). Inputs are what the forward left in
// HBM (the FK keypoints and their projection with the ORIGINAL camera matrix come from hrp_fk_project, as metrics.py:29-42 does
// through robot.get_keypoints_root + point_projection_from_3d) plus the ground truth of the batch; nothing goes through the host.
// HBM-trivial work (a few hundred bytes per frame): plain coalesced kernels, deterministic reductions (no atomics), fp32
// arithmetic per element like numpy on float32 arrays, fp64 for the sums numpy reports in fp64.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/hrp_b200.h"
#include "common.h"

namespace hrp {
namespace {

constexpr int MB_THREADS = 128;

// predicted pixel of keypoint k of frame b: taken from puv, or (puv == null) projected here with the frame's camera matrix
// exactly as point_projection_from_3d does (transforms.py:7-15: K @ p, divided by its last component)
__device__ __forceinline__ void pred_pixel(const float* __restrict__ puv, const float* __restrict__ Kmat, const float* p3, int b, int nk, int k,
                                           float* u, float* v) {
  if (puv != nullptr) { *u = puv[((size_t)b * nk + k) * 2]; *v = puv[((size_t)b * nk + k) * 2 + 1]; return; }
  const float* K = Kmat + (size_t)b * 9;
  const float x = p3[0], y = p3[1], z = p3[2];
  const float hx = K[0] * x + K[1] * y + K[2] * z, hy = K[3] * x + K[4] * y + K[5] * z, hz = K[6] * x + K[7] * y + K[8] * z;
  *u = hx / hz; *v = hy / hz;
}

// one thread per frame: metrics.py:55-69 (error3d, error2d over in-frame keypoints), 83-94 (joint error), 98-116 (root depth,
// root-relative depth, root-relative ADD)
__global__ void metrics_frame_kernel(const float* __restrict__ pxyz, const float* __restrict__ puv, const float* __restrict__ Kmat, const float* __restrict__ pq,
                                     const float* __restrict__ gxyz, const float* __restrict__ guv, const float* __restrict__ gq,
                                     int B, int nk, int dof, int root, int joint_cols, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p3 = pxyz + (size_t)b * nk * 3;
  const float* g3 = gxyz + (size_t)b * nk * 3;
  const float* g2 = guv + (size_t)b * nk * 2;
  const float pzr = p3[root * 3 + 2], gzr = g3[root * 3 + 2];
  float s3 = 0.f, s2 = 0.f, srel = 0.f, s3rel = 0.f;
  int nvalid = 0;
  for (int k = 0; k < nk; ++k) {
    const float dx = p3[k * 3] - g3[k * 3], dy = p3[k * 3 + 1] - g3[k * 3 + 1], dz = p3[k * 3 + 2] - g3[k * 3 + 2];
    s3 += sqrtf(dx * dx + dy * dy + dz * dz);
    float pu, pv;
    pred_pixel(puv, Kmat, p3 + k * 3, b, nk, k, &pu, &pv);
    const float ux = pu - g2[k * 2], uy = pv - g2[k * 2 + 1];
    const float gx = g2[k * 2], gy = g2[k * 2 + 1];
    const bool valid = gx <= 640.0f && gx >= 0.f && gy <= 480.0f && gy >= 0.f;     // metrics.py:63 (literal frame size)
    if (valid) { s2 += sqrtf(ux * ux + uy * uy); ++nvalid; }
    const float dr = (p3[k * 3 + 2] - pzr) - (g3[k * 3 + 2] - gzr);
    srel += fabsf(dr);
    s3rel += sqrtf(dx * dx + dy * dy + dr * dr);
  }
  float sj = 0.f;
  if (pq != nullptr)
    for (int j = 0; j < joint_cols; ++j) sj += fabsf(gq[(size_t)b * dof + j] - pq[(size_t)b * dof + j]);
  float* o = out + (size_t)b * 6;
  o[0] = s3 / (float)nk;                         // error3d
  o[1] = s2 / (float)nvalid;                     // error2d (0/0 = NaN when no keypoint is in the frame, as numpy)
  o[2] = pq != nullptr ? sj / (float)joint_cols : 0.f;   // mean_jointerror (Panda: without the finger joint, metrics.py:87-88)
  o[3] = fabsf(pzr - gzr);                       // error_depth
  o[4] = srel / (float)nk;                       // batch_error_relative
  o[5] = s3rel / (float)nk;                      // error3d_relative
}

// one block per column (keypoint 0..nk-1: dis3d, dis2d; joint nk..nk+dof-1: l1_jointerror): batch means, metrics.py:72-76, 86
__global__ void metrics_column_kernel(const float* __restrict__ pxyz, const float* __restrict__ puv, const float* __restrict__ Kmat, const float* __restrict__ pq,
                                      const float* __restrict__ gxyz, const float* __restrict__ guv, const float* __restrict__ gq,
                                      int B, int nk, int dof, float* __restrict__ dis3d, float* __restrict__ dis2d,
                                      float* __restrict__ l1joint) {
  __shared__ float sh[3][MB_THREADS];
  const int c = blockIdx.x, t = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (c < nk) {
    for (int b = t; b < B; b += MB_THREADS) {
      const float* p3 = pxyz + ((size_t)b * nk + c) * 3;
      const float* g3 = gxyz + ((size_t)b * nk + c) * 3;
      const float dx = p3[0] - g3[0], dy = p3[1] - g3[1], dz = p3[2] - g3[2];
      a0 += sqrtf(dx * dx + dy * dy + dz * dz);
      const float* g2 = guv + ((size_t)b * nk + c) * 2;
      if (g2[0] <= 640.0f && g2[0] >= 0.f && g2[1] <= 480.0f && g2[1] >= 0.f) {
        float pu, pv;
        pred_pixel(puv, Kmat, p3, b, nk, c, &pu, &pv);
        const float ux = pu - g2[0], uy = pv - g2[1];
        a1 += sqrtf(ux * ux + uy * uy);
        a2 += 1.f;
      }
    }
  } else if (pq != nullptr) {
    const int j = c - nk;
    for (int b = t; b < B; b += MB_THREADS) a0 += fabsf(gq[(size_t)b * dof + j] - pq[(size_t)b * dof + j]);
  }
  sh[0][t] = a0; sh[1][t] = a1; sh[2][t] = a2;
  __syncthreads();
  for (int s = MB_THREADS / 2; s > 0; s >>= 1) {
    if (t < s) { sh[0][t] += sh[0][t + s]; sh[1][t] += sh[1][t + s]; sh[2][t] += sh[2][t + s]; }
    __syncthreads();
  }
  if (t == 0) {
    if (c < nk) { dis3d[c] = sh[0][0] / (float)B; dis2d[c] = sh[1][0] / sh[2][0]; }
    else l1joint[c - nk] = sh[0][0] / (float)B;
  }
}

// ---- summary_add_pck ----------------------------------------------------------------------------------------------------
// Per element: how many of the n thresholds t_j = j*delta (np.arange(0, auc_threshold, delta): exactly j*delta in fp64)
// satisfy d <= t_j, the eight table thresholds, the value itself for the mean; per block partial sums in fp64.
constexpr int SM_THREADS = 256;
constexpr int SM_ACC = 12;      // 0 sum, 1 threshold hits, 2 hits at t_0, 3 hits at t_{n-1}, 4..11 table thresholds

__device__ __forceinline__ long long thresholds_at_or_above(double d, double delta, int n) {
  if (!(d == d)) return 0;                          // NaN compares false with everything
  if (d <= 0.0) return n;
  long long j = (long long)ceil(d / delta);         // first j with j*delta >= d, give or take one ulp of the division
  if (j > 0 && (double)(j - 1) * delta >= d) --j;
  if ((double)j * delta < d) ++j;
  return j >= n ? 0 : (long long)n - j;
}

__global__ void summary_partial_kernel(const float* __restrict__ v, long long N, double delta, int n, const double* __restrict__ table,
                                       double* __restrict__ partial) {
  __shared__ double sh[SM_ACC][SM_THREADS];
  double acc[SM_ACC];
#pragma unroll
  for (int i = 0; i < SM_ACC; ++i) acc[i] = 0.0;
  const double t_last = (double)(n - 1) * delta;
  for (long long i = (long long)blockIdx.x * SM_THREADS + threadIdx.x; i < N; i += (long long)gridDim.x * SM_THREADS) {
    const float f = v[i];
    const double d = (double)f;
    acc[0] += d;
    acc[1] += (double)thresholds_at_or_above(d, delta, n);
    acc[2] += d <= 0.0 ? 1.0 : 0.0;
    acc[3] += d <= t_last ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[4 + k] += d <= table[k] ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < SM_ACC; ++i) sh[i][threadIdx.x] = acc[i];
  __syncthreads();
  for (int s = SM_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int i = 0; i < SM_ACC; ++i) sh[i][threadIdx.x] += sh[i][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < SM_ACC) partial[(size_t)blockIdx.x * SM_ACC + threadIdx.x] = sh[threadIdx.x][0];
}

// order statistics by rank counting (exact, stable on ties; N is a test set of 10^3..10^5 frames): element i has rank
// #{j: v_j < v_i} + #{j < i: v_j == v_i}; the two middle ranks are written out. NaNs sort last like np.sort.
__device__ __forceinline__ bool less_nan_last(float a, float b) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return !an && bn;
  return a < b;
}
__global__ void median_rank_kernel(const float* __restrict__ v, long long N, long long k_lo, long long k_hi, float* __restrict__ mid) {
  __shared__ float tile[SM_THREADS];
  const long long i = (long long)blockIdx.x * SM_THREADS + threadIdx.x;
  const float mine = i < N ? v[i] : 0.f;
  long long rank = 0;
  for (long long j0 = 0; j0 < N; j0 += SM_THREADS) {
    const long long j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < N ? v[j] : 0.f;
    __syncthreads();
    const int lim = (int)((N - j0) < SM_THREADS ? (N - j0) : SM_THREADS);
    if (i < N)
      for (int q = 0; q < lim; ++q) {
        const float o = tile[q];
        const bool eq = (o == mine) || (o != o && mine != mine);
        rank += (less_nan_last(o, mine) || (eq && j0 + q < i)) ? 1 : 0;
      }
    __syncthreads();
  }
  if (i < N) {
    if (rank == k_lo) mid[0] = mine;
    if (rank == k_hi) mid[1] = mine;
  }
}

// out[0..10] for one error list: mean, median, AUC, then the eight table fractions
__global__ void summary_final_kernel(const double* __restrict__ partial, int blocks, const float* __restrict__ mid, long long N, double delta,
                                     int n, double auc_threshold, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc[SM_ACC];
  for (int i = 0; i < SM_ACC; ++i) acc[i] = 0.0;
  for (int b = 0; b < blocks; ++b)
    for (int i = 0; i < SM_ACC; ++i) acc[i] += partial[(size_t)b * SM_ACC + i];
  const double inv = 1.0 / (double)N;
  out[0] = acc[0] * inv;                                                   // np.mean
  // np.median of a float32 array: mean of the two middle values, in float32 (NaN if any value is NaN)
  const float m = (mid[0] + mid[1]) * 0.5f;
  out[1] = (mid[1] != mid[1]) ? (double)NAN : (double)m;
  // np.trapz(counts, dx=delta) / auc_threshold with counts_j = mean(d <= t_j): delta * (sum_j c_j - (c_0 + c_{n-1}) / 2)
  out[2] = delta * (acc[1] * inv - 0.5 * (acc[2] * inv + acc[3] * inv)) / auc_threshold;
  for (int k = 0; k < 8; ++k) out[3 + k] = acc[4 + k] * inv;
}

}  // namespace
}  // namespace hrp

using namespace hrp;

extern "C" int hrp_metrics_batch(const float* pred_xyz, const float* pred_uv, const float* K_original, const float* pred_joint, const float* gt_xyz,
                                 const float* gt_uv, const float* gt_joint, int B, int nkpt, int dof, int root_kp, int joint_cols,
                                 float* per_frame, float* dis3d, float* dis2d, float* l1_joint, void* stream) {
  if (B < 0 || nkpt <= 0 || nkpt > HRP_FK_MAX_KP || dof <= 0 || root_kp < 0 || root_kp >= nkpt || joint_cols <= 0 || joint_cols > dof)
    return fail(HRP_ERR_INVALID, "hrp_metrics_batch: bad sizes (B=%d nkpt=%d dof=%d root=%d joint_cols=%d)", B, nkpt, dof, root_kp, joint_cols);
  if (B == 0) return HRP_OK;
  if (!pred_xyz || (!pred_uv && !K_original) || !gt_xyz || !gt_uv || !per_frame || !dis3d || !dis2d || !l1_joint || (pred_joint && !gt_joint))
    return fail(HRP_ERR_INVALID, "hrp_metrics_batch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  metrics_frame_kernel<<<(B + MB_THREADS - 1) / MB_THREADS, MB_THREADS, 0, st>>>(pred_xyz, pred_uv, K_original, pred_joint, gt_xyz, gt_uv, gt_joint, B, nkpt, dof,
                                                                                root_kp, joint_cols, per_frame);
  HRP_CHECK_LAUNCH("metrics_frame_kernel");
  metrics_column_kernel<<<nkpt + dof, MB_THREADS, 0, st>>>(pred_xyz, pred_uv, K_original, pred_joint, gt_xyz, gt_uv, gt_joint, B, nkpt, dof, dis3d, dis2d, l1_joint);
  HRP_CHECK_LAUNCH("metrics_column_kernel");
  if (!pred_joint) HRP_CUDA(cudaMemsetAsync(l1_joint, 0, (size_t)dof * 4, st));        // metrics.py:91-92
  return HRP_OK;
}

extern "C" size_t hrp_summary_workspace(int64_t N) {
  const long long blocks = N <= 0 ? 1 : (N + SM_THREADS - 1) / SM_THREADS;
  const long long b = blocks < 1024 ? blocks : 1024;
  return (size_t)(2 * b * SM_ACC * 8 + 16 * 8 + 16);
}

extern "C" int hrp_summary_add_pck(const float* dis3d, const float* dis2d, int64_t N, double* out22, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (N <= 0) return fail(HRP_ERR_INVALID, "hrp_summary_add_pck: empty error lists");
  if (!dis3d || !dis2d || !out22 || !workspace) return fail(HRP_ERR_INVALID, "hrp_summary_add_pck: null argument");
  if (workspace_bytes < hrp_summary_workspace(N)) return fail(HRP_ERR_INVALID, "hrp_summary_add_pck: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<long long>((N + SM_THREADS - 1) / SM_THREADS, 1024);
  double* partial = static_cast<double*>(workspace);
  double* table = partial + (size_t)2 * blocks * SM_ACC;
  float* mid = reinterpret_cast<float*>(table + 16);
  // metrics.py:127-128, 159-162: ADD thresholds in millimetres (th * 1e-3 in fp64), PCK thresholds in pixels
  const double host_table[16] = {1 * 1e-3, 5 * 1e-3, 10 * 1e-3, 20 * 1e-3, 40 * 1e-3, 60 * 1e-3, 80 * 1e-3, 100 * 1e-3,
                                 2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0};
  HRP_CUDA(cudaMemcpyAsync(table, host_table, sizeof(host_table), cudaMemcpyHostToDevice, st));
  const long long k_hi = N / 2, k_lo = (N % 2) ? k_hi : k_hi - 1;
  const int rb = (int)((N + SM_THREADS - 1) / SM_THREADS);
  for (int which = 0; which < 2; ++which) {
    const float* v = which ? dis2d : dis3d;
    const double delta = which ? 0.01 : 0.00001, thr = which ? 20.0 : 0.1;   // metrics.py:131-133, 143-145
    const int n = which ? 2000 : 10000;                                       // len(np.arange(0, thr, delta))
    double* part = partial + (size_t)which * blocks * SM_ACC;
    summary_partial_kernel<<<blocks, SM_THREADS, 0, st>>>(v, N, delta, n, table + 8 * which, part);
    HRP_CHECK_LAUNCH("summary_partial_kernel");
    median_rank_kernel<<<rb, SM_THREADS, 0, st>>>(v, N, k_lo, k_hi, mid + 2 * which);
    HRP_CHECK_LAUNCH("median_rank_kernel");
    summary_final_kernel<<<1, 32, 0, st>>>(part, blocks, mid + 2 * which, N, delta, n, thr, out22 + 11 * which);
    HRP_CHECK_LAUNCH("summary_final_kernel");
  }
  return HRP_OK;
}
