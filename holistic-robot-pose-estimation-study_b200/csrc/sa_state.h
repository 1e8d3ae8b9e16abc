// Online-softmax state of the heatmap integral layer (lib/utils/integral.py:102-208): running maximum m, sum of
// exponentials l and the three first moments; two states merge exactly like split softmax. Shared by softargmax.cu (the
// stand-alone single-pass kernel) and conv_tc.cu (the final 1x1 conv whose epilogue reduces its own logits).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace hrp {

struct SaState {
  float m, l, sx, sy, sz;
};

__device__ __forceinline__ void sa_merge(SaState& a, const SaState& b) {
  const float M = fmaxf(a.m, b.m);
  if (M == -INFINITY) return;  // both empty
  const float fa = __expf(a.m - M), fb = __expf(b.m - M);
  a.l = a.l * fa + b.l * fb;
  a.sx = a.sx * fa + b.sx * fb;
  a.sy = a.sy * fa + b.sy * fb;
  a.sz = a.sz * fa + b.sz * fb;
  a.m = M;
}

__device__ __forceinline__ SaState sa_shfl_xor(const SaState& s, int off) {
  SaState o;
  o.m = __shfl_xor_sync(0xffffffffu, s.m, off);
  o.l = __shfl_xor_sync(0xffffffffu, s.l, off);
  o.sx = __shfl_xor_sync(0xffffffffu, s.sx, off);
  o.sy = __shfl_xor_sync(0xffffffffu, s.sy, off);
  o.sz = __shfl_xor_sync(0xffffffffu, s.sz, off);
  return o;
}

}  // namespace hrp
