// Heatmap integral layer: softmax over D*H*W per (frame, keypoint) + first moments of the three marginals, single pass.
//
// Replaces HeatmapIntegralPose.forward (lib/utils/integral.py:102-208: F.softmax over 262 144 bins, /sum, three
// marginal sums, three expectations, /size - 0.5, cat, fixroot) and uvd_to_xyz (lib/utils/transforms.py:33-82) with
// the inverse pinhole of integral.py:56-73. The reference makes >= 6 full passes over the logits in >= 12 launches;
// here the logits are read exactly once (HBM roofline: B*K*D*H*W*4 bytes):
//   kernel 1: grid (chunks, B*K). Each CTA streams one 64 KB chunk of one heatmap with 16-byte coalesced loads, four in
//             flight per thread, keeps a per-thread online-softmax state (running max m, sum l, moments sx, sy, sz),
//             combines it with warp shuffles + one shared-memory hop, and writes a 5-float partial.
//   kernel 2: one warp per (frame, keypoint) merges the partials (split-softmax merge), normalises, applies fixroot and
//             the camera back-projection.
// The expectation sum_i i*p_i over a marginal equals sum over all bins of coord*p, so no marginal is materialised.
#include "common.h"
#include "sa_state.h"

namespace hrp {

constexpr int SA_THREADS = 256;
constexpr int SA_CHUNK = 16384;       // floats per CTA (64 KB)
constexpr int SA_UNROLL = 4;          // float4 loads in flight per thread

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// hm: [rows][D*H*W]; partial: [rows][chunks][5]
template <bool POW2>
__global__ void __launch_bounds__(SA_THREADS)
softargmax_partial_kernel(const float* __restrict__ hm, float* __restrict__ partial, int total, int W, int H,
                          int lw, int lh, int chunks) {
  const int row = blockIdx.y, chunk = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(hm + (size_t)row * total);
  const int beg4 = chunk * (SA_CHUNK / 4);
  const int end4 = min(beg4 + SA_CHUNK / 4, total >> 2);
  SaState st{-INFINITY, 0.f, 0.f, 0.f, 0.f};

  for (int i4 = beg4 + threadIdx.x; i4 < end4; i4 += SA_THREADS * SA_UNROLL) {
    float4 v[SA_UNROLL];
#pragma unroll
    for (int u = 0; u < SA_UNROLL; ++u) {
      const int j = i4 + u * SA_THREADS;
      v[u] = (j < end4) ? ld_stream(src + j) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < SA_UNROLL; ++u) bm = fmaxf(bm, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
    if (bm > st.m) {  // rescale the running sums to the new maximum
      const float f = __expf(st.m - bm);   // exp(-inf) = 0 on the first batch
      st.l *= f; st.sx *= f; st.sy *= f; st.sz *= f;
      st.m = bm;
    }
#pragma unroll
    for (int u = 0; u < SA_UNROLL; ++u) {
      const int idx = (i4 + u * SA_THREADS) << 2;   // first bin of this float4: same (d,h), w0..w0+3
      int w0, h, d;
      if (POW2) {
        w0 = idx & (W - 1); h = (idx >> lw) & (H - 1); d = idx >> (lw + lh);
      } else {
        w0 = idx % W; const int r = idx / W; h = r % H; d = r / H;
      }
      const float e0 = __expf(v[u].x - st.m), e1 = __expf(v[u].y - st.m);
      const float e2 = __expf(v[u].z - st.m), e3 = __expf(v[u].w - st.m);
      const float s4 = (e0 + e1) + (e2 + e3);
      st.l += s4;
      st.sx += (float)w0 * s4 + (e1 + 2.f * e2 + 3.f * e3);
      st.sy += (float)h * s4;
      st.sz += (float)d * s4;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    SaState o = sa_shfl_xor(st, off);
    sa_merge(st, o);
  }
  __shared__ SaState warp_state[SA_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_state[wid] = st;
  __syncthreads();
  if (wid == 0) {
    st = (lane < SA_THREADS / 32) ? warp_state[lane] : SaState{-INFINITY, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      SaState o = sa_shfl_xor(st, off);
      sa_merge(st, o);
    }
    if (lane == 0) {
      float* p = partial + ((size_t)row * chunks + chunk) * 5;
      p[0] = st.m; p[1] = st.l; p[2] = st.sx; p[3] = st.sy; p[4] = st.sz;
    }
  }
}

struct SaTail {           // optional in-network outputs (NULL in the stand-alone op)
  float* root_uv;         // [B,2]   (uvd[:,ref,:2]+0.5)*image_size            full_net.py:360
  float* trans;           // [B,3]   uvz2xyz_singlepoint                        transforms.py:142-153
  float* kp2d;            // [B,K,2] projection of xyz                          transforms.py:17-21
};

__global__ void __launch_bounds__(128)
softargmax_finalize_kernel(const float* __restrict__ partial, int rows, int K, int chunks, int W, int H, int D,
                           const float* __restrict__ Kmat, const float* __restrict__ root_z, float depth_factor,
                           float image_size, int rootid, int fixroot, float* __restrict__ uvd,
                           float* __restrict__ xyz, SaTail tail) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  SaState st{-INFINITY, 0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < chunks; c += 32) {
    const float* p = partial + ((size_t)row * chunks + c) * 5;
    SaState o{p[0], p[1], p[2], p[3], p[4]};
    sa_merge(st, o);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    SaState o = sa_shfl_xor(st, off);
    sa_merge(st, o);
  }
  if (lane != 0) return;
  const int b = row / K, k = row % K;
  const float u = (st.sx / st.l) / (float)W - 0.5f;      // integral.py:138-146
  const float v = (st.sy / st.l) / (float)H - 0.5f;
  float d = (st.sz / st.l) / (float)D - 0.5f;
  if (fixroot && k == rootid) d = 0.f;                   // integral.py:150-151
  uvd[(size_t)row * 3 + 0] = u;
  uvd[(size_t)row * 3 + 1] = v;
  uvd[(size_t)row * 3 + 2] = d;
  if (xyz == nullptr) return;
  const float* Kb = Kmat + (size_t)b * 9;
  // inverse pinhole from fx, fy, cx, cy only, evaluated in double then rounded to fp32 (integral.py:60-65)
  const double fx = (double)Kb[0], fy = (double)Kb[4];
  const float i00 = (float)(1.0 / fx), i02 = (float)(-(double)Kb[2] / fx);
  const float i11 = (float)(1.0 / fy), i12 = (float)(-(double)Kb[5] / fy);
  const float up = (u + 0.5f) * image_size, vp = (v + 0.5f) * image_size;   // transforms.py:45-46
  const float z = d * depth_factor + root_z[b];                              // transforms.py:48, 71
  const float X = (i00 * up + i02) * z, Y = (i11 * vp + i12) * z, Z = z;
  xyz[(size_t)row * 3 + 0] = X;
  xyz[(size_t)row * 3 + 1] = Y;
  xyz[(size_t)row * 3 + 2] = Z;
  if (tail.kp2d != nullptr) {
    const float hx = Kb[0] * X + Kb[1] * Y + Kb[2] * Z;
    const float hy = Kb[3] * X + Kb[4] * Y + Kb[5] * Z;
    const float hz = Kb[6] * X + Kb[7] * Y + Kb[8] * Z;
    tail.kp2d[(size_t)row * 2 + 0] = hx / hz;
    tail.kp2d[(size_t)row * 2 + 1] = hy / hz;
  }
  if (tail.root_uv != nullptr && k == rootid) {
    const float zr = root_z[b];
    tail.root_uv[b * 2 + 0] = up;
    tail.root_uv[b * 2 + 1] = vp;
    tail.trans[b * 3 + 0] = i00 * (up * zr) + i02 * zr;   // inv_k @ [u*z, v*z, z]
    tail.trans[b * 3 + 1] = i11 * (vp * zr) + i12 * zr;
    tail.trans[b * 3 + 2] = zr;
  }
}

static inline int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

size_t softargmax_workspace(int B, int K, int D, int H, int W) {
  const long long total = (long long)D * H * W;
  const long long chunks = (total + SA_CHUNK - 1) / SA_CHUNK;
  return (size_t)B * K * chunks * 5 * sizeof(float);
}

int softargmax_launch(const float* hm, int B, int K, int D, int H, int W, const float* Kmat, const float* root_z,
                      float depth_factor, float image_size, int rootid, int fixroot, float* uvd, float* xyz,
                      void* ws, size_t ws_bytes, float* root_uv, float* trans, float* kp2d, cudaStream_t stream,
                      int* launches) {
  if (B <= 0) return HRP_OK;
  const long long total = (long long)D * H * W;
  if (K <= 0 || D <= 0 || H <= 0 || W <= 0 || total > (1LL << 30) || (W & 3) != 0)
    return fail(HRP_ERR_INVALID, "softargmax: unsupported heatmap shape K=%d D=%d H=%d W=%d (W must be a multiple of 4)", K, D, H, W);
  if ((reinterpret_cast<uintptr_t>(hm) & 15) != 0) return fail(HRP_ERR_INVALID, "softargmax: heatmap pointer must be 16-byte aligned");
  const int rows = B * K;
  const int chunks = (int)((total + SA_CHUNK - 1) / SA_CHUNK);
  if (ws_bytes < softargmax_workspace(B, K, D, H, W)) return fail(HRP_ERR_INVALID, "softargmax: workspace too small");
  if (rows > 65535) return fail(HRP_ERR_INVALID, "softargmax: B*K > 65535 not supported in one call");
  if (xyz != nullptr && (Kmat == nullptr || root_z == nullptr)) return fail(HRP_ERR_INVALID, "softargmax: xyz requested without Kmat/root_z");
  float* partial = static_cast<float*>(ws);
  const int lw = ilog2_exact(W), lh = ilog2_exact(H);
  dim3 grid(chunks, rows);
  if (lw >= 0 && lh >= 0)
    softargmax_partial_kernel<true><<<grid, SA_THREADS, 0, stream>>>(hm, partial, (int)total, W, H, lw, lh, chunks);
  else
    softargmax_partial_kernel<false><<<grid, SA_THREADS, 0, stream>>>(hm, partial, (int)total, W, H, 0, 0, chunks);
  HRP_CHECK_LAUNCH("softargmax_partial_kernel");
  SaTail tail{root_uv, trans, kp2d};
  softargmax_finalize_kernel<<<ceil_div(rows, 4), 128, 0, stream>>>(partial, rows, K, chunks, W, H, D, Kmat, root_z,
                                                                    depth_factor, image_size, rootid, fixroot, uvd, xyz, tail);
  HRP_CHECK_LAUNCH("softargmax_finalize_kernel");
  if (launches) *launches += 2;
  return HRP_OK;
}

// Second half only: merge `chunks` partial states per (frame, keypoint) that something else produced (the final conv's
// epilogue, conv_tc.cu) and finish exactly as the stand-alone op does.
int softargmax_finalize_launch(const float* partial, int B, int K, int chunks, int D, int H, int W, const float* Kmat,
                               const float* root_z, float depth_factor, float image_size, int rootid, int fixroot,
                               float* uvd, float* xyz, float* root_uv, float* trans, float* kp2d, cudaStream_t stream) {
  if (B <= 0) return HRP_OK;
  const int rows = B * K;
  SaTail tail{root_uv, trans, kp2d};
  softargmax_finalize_kernel<<<ceil_div(rows, 4), 128, 0, stream>>>(partial, rows, K, chunks, W, H, D, Kmat, root_z,
                                                                    depth_factor, image_size, rootid, fixroot, uvd, xyz, tail);
  HRP_CHECK_LAUNCH("softargmax_finalize_kernel");
  return HRP_OK;
}

}  // namespace hrp

extern "C" size_t hrp_softargmax3d_workspace(int B, int K, int D, int H, int W) {
  if (B <= 0 || K <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  return hrp::softargmax_workspace(B, K, D, H, W);
}

extern "C" int hrp_softargmax3d(const float* hm, int B, int K, int D, int H, int W, const float* Kmat,
                                const float* root_z, float depth_factor, float image_size, int rootid, int fixroot,
                                float* uvd, float* xyz, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace hrp;
  if (B < 0) return fail(HRP_ERR_INVALID, "hrp_softargmax3d: negative batch");
  if (B == 0) return HRP_OK;
  if (!hm || !uvd || !workspace) return fail(HRP_ERR_INVALID, "hrp_softargmax3d: null argument");
  return softargmax_launch(hm, B, K, D, H, W, Kmat, root_z, depth_factor, image_size, rootid, fixroot, uvd, xyz,
                           workspace, workspace_bytes, nullptr, nullptr, nullptr, (cudaStream_t)stream, nullptr);
}
