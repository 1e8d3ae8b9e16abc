// Batched URDF forward kinematics + pinhole projection, one pose per thread.
//
// Replaces URDFRobot.get_keypoints / get_keypoints_root (lib/utils/urdf_robot.py:95-118, 193-223), URDF.link_fk_batch
// (lib/utils/urdfpytorch/urdf.py:3064-3167), Joint.get_child_poses / _rotation_matrices (urdf.py:2345-2398, 2429-2464),
// rot6d_to_rotmat (lib/utils/geometries.py:100-115) and point_projection_from_3d_tensor (lib/utils/transforms.py:17-21).
//
// HBM-bound by design (244 / 260 / 472 B per pose for Panda / Kuka / Baxter):
//   * the kinematic program (compiled URDF) travels as a __grid_constant__ kernel parameter, so every table read is a
//     uniform constant-bank access;
//   * the CTA's slice of each AoS input array is contiguous in memory: it is copied with 16-byte coalesced loads into
//     shared memory and each thread then picks its own row; results go back the same way (row -> shared -> 16-byte
//     coalesced stores), so DRAM sees only full sectors;
//   * link frames that a branching tree needs later are parked in shared memory (one column per thread, conflict-free).
//
// Two kernels: fk_project_kernel interprets any compiled URDF program; fk_gen_kernel<ROBOT> runs the straight-line,
// constant-folded chain that scripts/gen_fk_programs.py generated from the packaged Panda / Kuka / Baxter URDFs
// (fk_programs_gen.h) inside a persistent CTA whose next input tile is already in flight (cp.async) while the current
// one is evaluated. A generated chain is used only when the incoming program is bitwise the one it was generated from.
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace hrp {

struct FkTables {
  int dof, nkpt, n_steps, n_slots, root_kp, root_step;
  int step_type[HRP_FK_MAX_STEPS];
  int step_parent[HRP_FK_MAX_STEPS];
  int step_save[HRP_FK_MAX_STEPS];
  int step_q[HRP_FK_MAX_STEPS];
  float step_mul[HRP_FK_MAX_STEPS];
  float step_off[HRP_FK_MAX_STEPS];
  float step_origin[HRP_FK_MAX_STEPS][12];
  float step_axis[HRP_FK_MAX_STEPS][3];
  int kp_step[HRP_FK_MAX_KP];
  int kp_index[HRP_FK_MAX_KP];
  float kp_offset[HRP_FK_MAX_KP][3];
  float root_fixed[12];
};

constexpr int FK_THREADS = 128;

// Raw program tables of the generated chains (bit patterns), matched against an incoming hrp_fk_program.
struct FkRaw {
  int dof, nkpt, n_steps, n_slots, root_kp, root_step;
  int step_type[HRP_FK_MAX_STEPS], step_parent[HRP_FK_MAX_STEPS], step_save[HRP_FK_MAX_STEPS], step_q[HRP_FK_MAX_STEPS];
  uint32_t step_mul[HRP_FK_MAX_STEPS], step_off[HRP_FK_MAX_STEPS], step_origin[HRP_FK_MAX_STEPS * 12], step_axis[HRP_FK_MAX_STEPS * 3];
  int kp_step[HRP_FK_MAX_KP], kp_index[HRP_FK_MAX_KP];
  uint32_t kp_offset[HRP_FK_MAX_KP * 3], root_fixed[12];
};
enum { FK_GENERIC = 0, FK_PANDA = 1, FK_KUKA = 2, FK_BAXTER = 3 };
template <int ROBOT> struct FkGen;

// sin/cos for joint angles: three-constant Cody-Waite reduction by pi/2 + the classic single-precision minimax kernels
// (max abs error 8e-8 for |x| <= 50 against float64). One shared out-of-line copy of sincosf serves |x| > 1000, instead of
// one inlined Payne-Hanek slow path per joint.
__device__ __noinline__ float2 fk_sincos_slow(float x) { float2 r; sincosf(x, &r.x, &r.y); return r; }
__device__ __forceinline__ void fk_sincos(float x, float& s, float& c) {
  if (fabsf(x) > 1000.f) { const float2 r = fk_sincos_slow(x); s = r.x; c = r.y; return; }
  const float kf = rintf(x * 0.636619772f);
  const int k = (int)kf;
  float r = fmaf(kf, -0x1.921fb6p+0f, x);
  r = fmaf(kf, 0x1.777a5cp-25f, r);
  r = fmaf(kf, 0x1.0p-49f, r);
  const float z = r * r;
  const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
  const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f) * z, z, fmaf(-0.5f, z, 1.f));
  const float a = (k & 1) ? cp : sp, b = (k & 1) ? sp : cp;
  s = (k & 2) ? -a : a;
  c = ((k + 1) & 2) ? -b : b;
}
#include "fk_programs_gen.h"

struct Rt {  // rigid transform, row-major 3x3 + translation
  float r[9];
  float t[3];
};

__device__ __forceinline__ void mul(const Rt& a, const Rt& b, Rt& c) {  // c = a * b
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.r[i * 3 + j] = a.r[i * 3 + 0] * b.r[0 * 3 + j] + a.r[i * 3 + 1] * b.r[1 * 3 + j] + a.r[i * 3 + 2] * b.r[2 * 3 + j];
    c.t[i] = a.r[i * 3 + 0] * b.t[0] + a.r[i * 3 + 1] * b.t[1] + a.r[i * 3 + 2] * b.t[2] + a.t[i];
  }
}

// Cooperative contiguous copy global -> shared of `count` floats starting at g (16-byte path when aligned).
__device__ __forceinline__ void stage_in(float* s, const float* __restrict__ g, int count) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = count >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(s);
    for (int i = threadIdx.x; i < n4; i += FK_THREADS) s4[i] = __ldg(g4 + i);
    for (int i = (n4 << 2) + threadIdx.x; i < count; i += FK_THREADS) s[i] = __ldg(g + i);
  } else {
    for (int i = threadIdx.x; i < count; i += FK_THREADS) s[i] = __ldg(g + i);
  }
}

__device__ __forceinline__ void stage_out(float* __restrict__ g, const float* s, int count) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = count >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = threadIdx.x; i < n4; i += FK_THREADS) g4[i] = s4[i];
    for (int i = (n4 << 2) + threadIdx.x; i < count; i += FK_THREADS) g[i] = s[i];
  } else {
    for (int i = threadIdx.x; i < count; i += FK_THREADS) g[i] = s[i];
  }
}

__global__ void __launch_bounds__(FK_THREADS)
fk_project_kernel(const __grid_constant__ FkTables P, const float* __restrict__ q, const float* __restrict__ rot6d,
                  const float* __restrict__ trans, const float* __restrict__ Kmat, long long N,
                  float* __restrict__ xyz, float* __restrict__ uv) {
  extern __shared__ __align__(16) float smem[];
  const int dof = P.dof, nk = P.nkpt;
  // region A (inputs) is reused for the outputs after the chain has been evaluated
  const int in_q = 0;
  const int in_rot = in_q + ((FK_THREADS * dof + 3) & ~3);
  const int in_tr = in_rot + FK_THREADS * 6;
  const int in_k = in_tr + FK_THREADS * 3;
  const int in_end = in_k + FK_THREADS * 9;
  const int out_xyz = 0;
  const int out_uv = FK_THREADS * nk * 3;
  const int out_end = out_uv + FK_THREADS * nk * 2;
  const int io_end = (in_end > out_end ? in_end : out_end);
  float* s_kp = smem + ((io_end + 3) & ~3);                       // [nk*3][FK_THREADS] base-frame keypoints
  float* s_slot = s_kp + nk * 3 * FK_THREADS;                     // [n_slots*12][FK_THREADS]

  const long long base = (long long)blockIdx.x * FK_THREADS;
  const int valid = (int)min((long long)FK_THREADS, N - base);
  const int t = threadIdx.x;

  stage_in(smem + in_q, q + base * dof, valid * dof);
  stage_in(smem + in_rot, rot6d + base * 6, valid * 6);
  stage_in(smem + in_tr, trans + base * 3, valid * 3);
  stage_in(smem + in_k, Kmat + base * 9, valid * 9);
  __syncthreads();

  float Kc[9];
  Rt A;  // base (or root link) -> camera
  if (t < valid) {
    const float* sq = smem + in_q + t * dof;
    // ---- base-frame chain -------------------------------------------------------------------------------------
    Rt T;       // current link frame
    Rt Troot;   // frame of the re-rooting link
#pragma unroll
    for (int i = 0; i < 9; ++i) { T.r[i] = (i % 4 == 0) ? 1.f : 0.f; Troot.r[i] = T.r[i]; }
    T.t[0] = T.t[1] = T.t[2] = 0.f;
    Troot.t[0] = Troot.t[1] = Troot.t[2] = 0.f;
    int kp = 0;
    for (; kp < nk && P.kp_step[kp] < 0; ++kp) {   // keypoints rigidly attached to the base
      const int o = P.kp_index[kp] * 3;
      s_kp[(o + 0) * FK_THREADS + t] = P.kp_offset[kp][0];
      s_kp[(o + 1) * FK_THREADS + t] = P.kp_offset[kp][1];
      s_kp[(o + 2) * FK_THREADS + t] = P.kp_offset[kp][2];
    }
    for (int s = 0; s < P.n_steps; ++s) {
      Rt M;  // origin * motion(q)   (urdf.py:2385-2392)
      const float qv = P.step_mul[s] * sq[P.step_q[s]] + P.step_off[s];
      const float* O = P.step_origin[s];
      const float ax = P.step_axis[s][0], ay = P.step_axis[s][1], az = P.step_axis[s][2];
      if (P.step_type[s] == 1) {
        float sn, cs;
        sincosf(qv, &sn, &cs);
        const float oc = 1.f - cs;
        // cos*I + (1-cos)*a a^T + sin*[a]x   (urdf.py:2451-2463)
        float R[9];
        R[0] = cs + ax * ax * oc;      R[1] = ax * ay * oc - az * sn; R[2] = ax * az * oc + ay * sn;
        R[3] = ay * ax * oc + az * sn; R[4] = cs + ay * ay * oc;      R[5] = ay * az * oc - ax * sn;
        R[6] = az * ax * oc - ay * sn; R[7] = az * ay * oc + ax * sn; R[8] = cs + az * az * oc;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
          for (int j = 0; j < 3; ++j)
            M.r[i * 3 + j] = O[i * 4 + 0] * R[j] + O[i * 4 + 1] * R[3 + j] + O[i * 4 + 2] * R[6 + j];
          M.t[i] = O[i * 4 + 3];
        }
      } else {  // prismatic: translate along the (un-normalised) axis
        const float dx = ax * qv, dy = ay * qv, dz = az * qv;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          M.r[i * 3 + 0] = O[i * 4 + 0]; M.r[i * 3 + 1] = O[i * 4 + 1]; M.r[i * 3 + 2] = O[i * 4 + 2];
          M.t[i] = O[i * 4 + 0] * dx + O[i * 4 + 1] * dy + O[i * 4 + 2] * dz + O[i * 4 + 3];
        }
      }
      const int par = P.step_parent[s];
      if (par == HRP_FK_PARENT_BASE) {
        T = M;
      } else {
        Rt Pm;
        if (par == HRP_FK_PARENT_PREV) {
          Pm = T;
        } else {
#pragma unroll
          for (int i = 0; i < 9; ++i) Pm.r[i] = s_slot[(par * 12 + i) * FK_THREADS + t];
#pragma unroll
          for (int i = 0; i < 3; ++i) Pm.t[i] = s_slot[(par * 12 + 9 + i) * FK_THREADS + t];
        }
        mul(Pm, M, T);  // fk[child] = fk[parent] @ (origin @ motion)   (urdf.py:3152-3156)
      }
      const int sv = P.step_save[s];
      if (sv >= 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) s_slot[(sv * 12 + i) * FK_THREADS + t] = T.r[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) s_slot[(sv * 12 + 9 + i) * FK_THREADS + t] = T.t[i];
      }
      if (s == P.root_step) Troot = T;
      for (; kp < nk && P.kp_step[kp] == s; ++kp) {   // pts = R*offset + t   (urdf_robot.py:117)
        const float ox = P.kp_offset[kp][0], oy = P.kp_offset[kp][1], oz = P.kp_offset[kp][2];
        const int o = P.kp_index[kp] * 3;
        s_kp[(o + 0) * FK_THREADS + t] = T.r[0] * ox + T.r[1] * oy + T.r[2] * oz + T.t[0];
        s_kp[(o + 1) * FK_THREADS + t] = T.r[3] * ox + T.r[4] * oy + T.r[5] * oz + T.t[1];
        s_kp[(o + 2) * FK_THREADS + t] = T.r[6] * ox + T.r[7] * oy + T.r[8] * oz + T.t[2];
      }
    }
    // ---- camera pose of the base / root link: rot6d -> R (rows x, y, z), geometries.py:100-115 -----------------
    const float* sr = smem + in_rot + t * 6;
    float x0 = sr[0], x1 = sr[1], x2 = sr[2];
    const float y0 = sr[3], y1 = sr[4], y2 = sr[5];
    const float nx = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
    x0 = x0 / nx; x1 = x1 / nx; x2 = x2 / nx;
    float z0 = x1 * y2 - x2 * y1, z1 = x2 * y0 - x0 * y2, z2 = x0 * y1 - x1 * y0;
    const float nz = sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
    z0 = z0 / nz; z1 = z1 / nz; z2 = z2 / nz;
    Rt C;
    C.r[0] = x0; C.r[1] = x1; C.r[2] = x2;
    C.r[3] = z1 * x2 - z2 * x1; C.r[4] = z2 * x0 - z0 * x2; C.r[5] = z0 * x1 - z1 * x0;
    C.r[6] = z0; C.r[7] = z1; C.r[8] = z2;
    const float* st = smem + in_tr + t * 3;
    C.t[0] = st[0]; C.t[1] = st[1]; C.t[2] = st[2];
    if (P.root_kp != 0) {
      // TWL = base2cam @ inv(TWL_base[root]) @ TWL_base (urdf_robot.py:218-221); rigid inverse in closed form
      Rt F, Tr, Ti;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        F.r[i * 3 + 0] = P.root_fixed[i * 4 + 0]; F.r[i * 3 + 1] = P.root_fixed[i * 4 + 1];
        F.r[i * 3 + 2] = P.root_fixed[i * 4 + 2]; F.t[i] = P.root_fixed[i * 4 + 3];
      }
      mul(Troot, F, Tr);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Ti.r[i * 3 + j] = Tr.r[j * 3 + i];
        Ti.t[i] = -(Tr.r[0 * 3 + i] * Tr.t[0] + Tr.r[1 * 3 + i] * Tr.t[1] + Tr.r[2 * 3 + i] * Tr.t[2]);
      }
      mul(C, Ti, A);
    } else {
      A = C;
    }
    const float* sk = smem + in_k + t * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) Kc[i] = sk[i];
  }
  __syncthreads();  // inputs consumed: region A becomes the output staging area
  if (t < valid) {
    float* so = smem + out_xyz + t * nk * 3;
    float* su = smem + out_uv + t * nk * 2;
    for (int k = 0; k < nk; ++k) {
      const float px = s_kp[(k * 3 + 0) * FK_THREADS + t];
      const float py = s_kp[(k * 3 + 1) * FK_THREADS + t];
      const float pz = s_kp[(k * 3 + 2) * FK_THREADS + t];
      const float cx = A.r[0] * px + A.r[1] * py + A.r[2] * pz + A.t[0];
      const float cy = A.r[3] * px + A.r[4] * py + A.r[5] * pz + A.t[1];
      const float cz = A.r[6] * px + A.r[7] * py + A.r[8] * pz + A.t[2];
      so[k * 3 + 0] = cx; so[k * 3 + 1] = cy; so[k * 3 + 2] = cz;
      // hnormalized(K @ p)   (transforms.py:7-9, 17-21)
      const float hx = Kc[0] * cx + Kc[1] * cy + Kc[2] * cz;
      const float hy = Kc[3] * cx + Kc[4] * cy + Kc[5] * cz;
      const float hz = Kc[6] * cx + Kc[7] * cy + Kc[8] * cz;
      su[k * 2 + 0] = hx / hz;
      su[k * 2 + 1] = hy / hz;
    }
  }
  __syncthreads();
  stage_out(xyz + base * nk * 3, smem + out_xyz, valid * nk * 3);
  if (uv != nullptr) stage_out(uv + base * nk * 2, smem + out_uv, valid * nk * 2);
}

// ---- generated-chain kernel ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void load_row(const float* s, float (&r)[N]) {     // s is 4*N-byte-strided, base 16-byte aligned
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) { const float4 v = reinterpret_cast<const float4*>(s)[i]; r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w; }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { const float2 v = reinterpret_cast<const float2*>(s)[i]; r[2 * i] = v.x; r[2 * i + 1] = v.y; }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = s[i];
  }
}
template <int N>
__device__ __forceinline__ void store_row(float* s, const float (&r)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) reinterpret_cast<float4*>(s)[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) reinterpret_cast<float2*>(s)[i] = make_float2(r[2 * i], r[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = r[i];
  }
}

__device__ __forceinline__ void fk_cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// Persistent CTA over 128-pose tiles. Shared memory: one input tile + one output staging tile. A tile's rows move to
// registers first; the input buffer is then free and the NEXT tile's cp.async copies are issued before the chain is
// evaluated, so they have the whole evaluation + copy-out of this tile to land. Base-frame keypoints stay in registers
// (the chain is straight-line code, every index a compile-time constant).
template <int ROBOT>
__global__ void __launch_bounds__(FK_THREADS, FkGen<ROBOT>::MIN_CTAS)
fk_gen_kernel(const float* __restrict__ q, const float* __restrict__ rot6d, const float* __restrict__ trans,
              const float* __restrict__ Kmat, long long N, float* __restrict__ xyz, float* __restrict__ uv, int async_ok) {
  using G = FkGen<ROBOT>;
  constexpr int DOF = G::DOF, NK = G::NK;
  constexpr int IN_Q = 0, IN_ROT = IN_Q + FK_THREADS * DOF, IN_TR = IN_ROT + FK_THREADS * 6, IN_K = IN_TR + FK_THREADS * 3,
                IN_END = IN_K + FK_THREADS * 9;
  constexpr int OUT_XYZ = IN_END, OUT_UV = OUT_XYZ + FK_THREADS * NK * 3;
  extern __shared__ __align__(16) float smem[];
  const int t = threadIdx.x;
  const long long ntiles = (N + FK_THREADS - 1) / FK_THREADS;

  // (no pointer / count arrays here: indexing them, even in an unrolled loop, put a 32-byte frame into local memory)
  auto copy_in = [&](float* dst, const float* src, int n16) {
    for (int i = t; i < n16; i += FK_THREADS) fk_cp_async16(dst + 4 * i, src + 4 * i);
  };
  auto prefetch = [&](long long tile) {      // full tiles only; always commits a group so the counting stays uniform
    if (async_ok && tile < ntiles && (tile + 1) * FK_THREADS <= N) {
      const long long base = tile * FK_THREADS;
      copy_in(smem + IN_Q, q + base * DOF, FK_THREADS * DOF / 4);
      copy_in(smem + IN_ROT, rot6d + base * 6, FK_THREADS * 6 / 4);
      copy_in(smem + IN_TR, trans + base * 3, FK_THREADS * 3 / 4);
      copy_in(smem + IN_K, Kmat + base * 9, FK_THREADS * 9 / 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  long long tile = blockIdx.x;
  prefetch(tile);
  for (; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * FK_THREADS;
    const int valid = (int)min((long long)FK_THREADS, N - base);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!async_ok || valid < FK_THREADS) {                 // unaligned arrays or the ragged last tile: synchronous staging
      stage_in(smem + IN_Q, q + base * DOF, valid * DOF);
      stage_in(smem + IN_ROT, rot6d + base * 6, valid * 6);
      stage_in(smem + IN_TR, trans + base * 3, valid * 3);
      stage_in(smem + IN_K, Kmat + base * 9, valid * 9);
    }
    __syncthreads();                                       // inputs landed; previous tile's staging has been copied out
    float qr[DOF], sr[6], st[3], Kc[9];
    if (t < valid) {
      load_row<DOF>(smem + IN_Q + t * DOF, qr);
      load_row<6>(smem + IN_ROT + t * 6, sr);
      load_row<3>(smem + IN_TR + t * 3, st);
      load_row<9>(smem + IN_K + t * 9, Kc);
    }
    __syncthreads();                                       // every row is in registers: refill the input buffer now
    prefetch(tile + gridDim.x);
    if (t < valid) {
      float kp[NK * 3], TrG[12];
      G::chain(qr, kp, TrG);
      // camera pose of the base / root link: rot6d -> R (rows x, y, z), geometries.py:100-115
      float x0 = sr[0], x1 = sr[1], x2 = sr[2];
      const float y0 = sr[3], y1 = sr[4], y2 = sr[5];
      const float inx = 1.f / sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
      x0 *= inx; x1 *= inx; x2 *= inx;
      float z0 = x1 * y2 - x2 * y1, z1 = x2 * y0 - x0 * y2, z2 = x0 * y1 - x1 * y0;
      const float inz = 1.f / sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
      z0 *= inz; z1 *= inz; z2 *= inz;
      Rt C, A;
      C.r[0] = x0; C.r[1] = x1; C.r[2] = x2;
      C.r[3] = z1 * x2 - z2 * x1; C.r[4] = z2 * x0 - z0 * x2; C.r[5] = z0 * x1 - z1 * x0;
      C.r[6] = z0; C.r[7] = z1; C.r[8] = z2;
      C.t[0] = st[0]; C.t[1] = st[1]; C.t[2] = st[2];
      if constexpr (G::ROOT_KP != 0) {
        // TWL = base2cam @ inv(TWL_base[root]) @ TWL_base (urdf_robot.py:218-221); rigid inverse in closed form
        Rt Ti;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
          for (int j = 0; j < 3; ++j) Ti.r[i * 3 + j] = TrG[j * 3 + i];
          Ti.t[i] = -(TrG[0 * 3 + i] * TrG[9] + TrG[1 * 3 + i] * TrG[10] + TrG[2 * 3 + i] * TrG[11]);
        }
        mul(C, Ti, A);
      } else {
        A = C;
      }
      float oxyz[NK * 3], ouv[NK * 2];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const float px = kp[k * 3], py = kp[k * 3 + 1], pz = kp[k * 3 + 2];
        const float cx = A.r[0] * px + A.r[1] * py + A.r[2] * pz + A.t[0];
        const float cy = A.r[3] * px + A.r[4] * py + A.r[5] * pz + A.t[1];
        const float cz = A.r[6] * px + A.r[7] * py + A.r[8] * pz + A.t[2];
        oxyz[k * 3] = cx; oxyz[k * 3 + 1] = cy; oxyz[k * 3 + 2] = cz;
        // hnormalized(K @ p)   (transforms.py:7-9, 17-21)
        const float hx = Kc[0] * cx + Kc[1] * cy + Kc[2] * cz;
        const float hy = Kc[3] * cx + Kc[4] * cy + Kc[5] * cz;
        const float ihz = 1.f / (Kc[6] * cx + Kc[7] * cy + Kc[8] * cz);
        ouv[k * 2] = hx * ihz;
        ouv[k * 2 + 1] = hy * ihz;
      }
      store_row<NK * 3>(smem + OUT_XYZ + t * NK * 3, oxyz);
      store_row<NK * 2>(smem + OUT_UV + t * NK * 2, ouv);
    }
    __syncthreads();
    stage_out(xyz + base * NK * 3, smem + OUT_XYZ, valid * NK * 3);
    if (uv != nullptr) stage_out(uv + base * NK * 2, smem + OUT_UV, valid * NK * 2);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int ROBOT>
constexpr size_t fk_gen_smem() {
  using G = FkGen<ROBOT>;
  return sizeof(float) * (size_t)FK_THREADS * (G::DOF + 18 + G::NK * 5);
}

template <int ROBOT>
int fk_gen_launch(const float* q, const float* rot6d, const float* trans, const float* Kmat, int64_t N, float* xyz, float* uv, cudaStream_t stream) {
  static int ctas_per_sm = 0;
  constexpr size_t smem = fk_gen_smem<ROBOT>();
  if (ctas_per_sm == 0) {
    HRP_CUDA(cudaFuncSetAttribute(fk_gen_kernel<ROBOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    HRP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fk_gen_kernel<ROBOT>, FK_THREADS, smem));
    ctas_per_sm = n > 0 ? n : 1;
  }
  const int64_t tiles = ceil_div64(N, FK_THREADS);
  const int64_t grid = tiles < (int64_t)sm_count() * ctas_per_sm ? tiles : (int64_t)sm_count() * ctas_per_sm;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(rot6d) | reinterpret_cast<uintptr_t>(trans) | reinterpret_cast<uintptr_t>(Kmat);
  fk_gen_kernel<ROBOT><<<(unsigned)grid, FK_THREADS, smem, stream>>>(q, rot6d, trans, Kmat, (long long)N, xyz, uv, (bits & 15) == 0 ? 1 : 0);
  HRP_CHECK_LAUNCH("fk_gen_kernel");
  return HRP_OK;
}

static uint32_t f32_bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }

static bool fk_matches(const FkTables& T, const FkRaw& R) {
  if (T.dof != R.dof || T.nkpt != R.nkpt || T.n_steps != R.n_steps || T.n_slots != R.n_slots || T.root_kp != R.root_kp || T.root_step != R.root_step) return false;
  for (int s = 0; s < T.n_steps; ++s) {
    if (T.step_type[s] != R.step_type[s] || T.step_parent[s] != R.step_parent[s] || T.step_save[s] != R.step_save[s] || T.step_q[s] != R.step_q[s]) return false;
    if (f32_bits(T.step_mul[s]) != R.step_mul[s] || f32_bits(T.step_off[s]) != R.step_off[s]) return false;
    for (int i = 0; i < 12; ++i) if (f32_bits(T.step_origin[s][i]) != R.step_origin[s * 12 + i]) return false;
    for (int i = 0; i < 3; ++i) if (f32_bits(T.step_axis[s][i]) != R.step_axis[s * 3 + i]) return false;
  }
  for (int k = 0; k < T.nkpt; ++k) {
    if (T.kp_step[k] != R.kp_step[k] || T.kp_index[k] != R.kp_index[k]) return false;
    for (int i = 0; i < 3; ++i) if (f32_bits(T.kp_offset[k][i]) != R.kp_offset[k * 3 + i]) return false;
  }
  for (int i = 0; i < 12; ++i) if (f32_bits(T.root_fixed[i]) != R.root_fixed[i]) return false;
  return true;
}

static size_t fk_smem_bytes(const FkTables& P) {
  const int dof = P.dof, nk = P.nkpt;
  const int in_end = ((FK_THREADS * dof + 3) & ~3) + FK_THREADS * (6 + 3 + 9);
  const int out_end = FK_THREADS * nk * 5;
  const int io_end = ((in_end > out_end ? in_end : out_end) + 3) & ~3;
  return sizeof(float) * (size_t)(io_end + nk * 3 * FK_THREADS + P.n_slots * 12 * FK_THREADS);
}

}  // namespace hrp

struct hrp_fk {
  hrp::FkTables tab;
  size_t smem;
  int robot;      // FK_GENERIC, or the generated chain this program is bitwise identical to
};

namespace hrp {
int fk_launch(const hrp_fk* fk, const float* q, const float* rot6d, const float* trans, const float* Kmat, int64_t N,
              float* xyz, float* uv, cudaStream_t stream) {
  if (N <= 0) return HRP_OK;
  switch (fk->robot) {
    case FK_PANDA: return fk_gen_launch<FK_PANDA>(q, rot6d, trans, Kmat, N, xyz, uv, stream);
    case FK_KUKA: return fk_gen_launch<FK_KUKA>(q, rot6d, trans, Kmat, N, xyz, uv, stream);
    case FK_BAXTER: return fk_gen_launch<FK_BAXTER>(q, rot6d, trans, Kmat, N, xyz, uv, stream);
    default: break;
  }
  const int64_t blocks = ceil_div64(N, FK_THREADS);
  if (blocks > 0x7fffffffLL) return fail(HRP_ERR_INVALID, "hrp_fk_project: N too large");
  fk_project_kernel<<<(unsigned)blocks, FK_THREADS, fk->smem, stream>>>(fk->tab, q, rot6d, trans, Kmat, (long long)N, xyz, uv);
  HRP_CHECK_LAUNCH("fk_project_kernel");
  return HRP_OK;
}
}  // namespace hrp

extern "C" int hrp_fk_create(const hrp_fk_program* p, hrp_fk** out) {
  using namespace hrp;
  if (!p || !out) return fail(HRP_ERR_INVALID, "hrp_fk_create: null argument");
  if (p->dof <= 0 || p->dof > 64 || p->nkpt <= 0 || p->nkpt > HRP_FK_MAX_KP || p->n_steps < 0 ||
      p->n_steps > HRP_FK_MAX_STEPS || p->n_slots < 0 || p->n_slots > HRP_FK_MAX_SLOTS || p->root_kp < 0 ||
      p->root_kp >= p->nkpt || p->root_step < -1 || p->root_step >= p->n_steps)
    return fail(HRP_ERR_INVALID, "hrp_fk_create: program out of range (dof=%d nkpt=%d steps=%d slots=%d)", p->dof,
                p->nkpt, p->n_steps, p->n_slots);
  hrp_fk* fk = new hrp_fk();
  FkTables& T = fk->tab;
  T.dof = p->dof; T.nkpt = p->nkpt; T.n_steps = p->n_steps; T.n_slots = p->n_slots;
  T.root_kp = p->root_kp; T.root_step = p->root_step;
  for (int s = 0; s < p->n_steps; ++s) {
    T.step_type[s] = p->step_type[s];
    T.step_parent[s] = p->step_parent[s];
    T.step_save[s] = p->step_save[s];
    T.step_q[s] = p->step_q[s];
    T.step_mul[s] = p->step_mul[s];
    T.step_off[s] = p->step_off[s];
    for (int i = 0; i < 12; ++i) T.step_origin[s][i] = p->step_origin[s * 12 + i];
    for (int i = 0; i < 3; ++i) T.step_axis[s][i] = p->step_axis[s * 3 + i];
    const bool ok = (T.step_type[s] == 1 || T.step_type[s] == 2) && T.step_q[s] >= 0 && T.step_q[s] < p->dof &&
                    T.step_parent[s] >= HRP_FK_PARENT_PREV && T.step_parent[s] < p->n_slots &&
                    T.step_save[s] >= -1 && T.step_save[s] < p->n_slots && !(s == 0 && T.step_parent[s] == HRP_FK_PARENT_PREV);
    if (!ok) { delete fk; return fail(HRP_ERR_INVALID, "hrp_fk_create: bad step %d", s); }
  }
  int prev = -1;
  for (int k = 0; k < p->nkpt; ++k) {
    T.kp_step[k] = p->kp_step[k];
    T.kp_index[k] = p->kp_index[k];
    for (int i = 0; i < 3; ++i) T.kp_offset[k][i] = p->kp_offset[k * 3 + i];
    if (T.kp_step[k] < prev || T.kp_step[k] >= p->n_steps || T.kp_index[k] < 0 || T.kp_index[k] >= p->nkpt) {
      delete fk; return fail(HRP_ERR_INVALID, "hrp_fk_create: bad keypoint entry %d", k);
    }
    prev = T.kp_step[k];
  }
  for (int i = 0; i < 12; ++i) T.root_fixed[i] = p->root_fixed[i];
  fk->smem = fk_smem_bytes(T);
  fk->robot = FK_GENERIC;
  if (!getenv("HRP_FK_GENERIC")) {   // development switch: force the table interpreter
    if (fk_matches(T, fk_raw_panda)) fk->robot = FK_PANDA;
    else if (fk_matches(T, fk_raw_kuka)) fk->robot = FK_KUKA;
    else if (fk_matches(T, fk_raw_baxter)) fk->robot = FK_BAXTER;
  }
  if (fk->smem > 200 * 1024) { const size_t need = fk->smem; delete fk; return fail(HRP_ERR_INVALID, "hrp_fk_create: program needs %zu B shared memory", need); }
  if (fk->smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fk_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fk->smem);
    if (e != cudaSuccess) {
      // no device in this process (CPU-only host): the attribute is set again lazily on first launch
      cudaGetLastError();
    }
  }
  *out = fk;
  return HRP_OK;
}

extern "C" void hrp_fk_destroy(hrp_fk* fk) { delete fk; }

extern "C" int hrp_fk_project(const hrp_fk* fk, const float* q, const float* rot6d, const float* trans,
                              const float* Kmat, int64_t N, float* xyz, float* uv, void* stream) {
  using namespace hrp;
  if (!fk) return fail(HRP_ERR_INVALID, "hrp_fk_project: null handle");
  if (N < 0) return fail(HRP_ERR_INVALID, "hrp_fk_project: negative N");
  if (N == 0) return HRP_OK;   // empty batch: nothing to do, pointers may be null
  if (!q || !rot6d || !trans || !Kmat || !xyz) return fail(HRP_ERR_INVALID, "hrp_fk_project: null argument");
  if (fk->smem > 48 * 1024)
    HRP_CUDA(cudaFuncSetAttribute(fk_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fk->smem));
  return fk_launch(fk, q, rot6d, trans, Kmat, N, xyz, uv, (cudaStream_t)stream);
}
