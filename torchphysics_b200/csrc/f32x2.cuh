// Packed fp32 arithmetic of sm_100 (add / mul / fma .f32x2: two IEEE round-to-nearest operations per instruction, one issue
// slot).  The thread-local jet formulas are independent per neuron, so neighbouring neurons are processed as pairs: the FP
// instruction count of the activation warps halves while every result stays bit-identical to the scalar formulas.
#pragma once
#include <cuda_runtime.h>

namespace tpx {

struct F2 {
  float2 v;
};
__device__ __forceinline__ F2 f2(float a, float b) { F2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ F2 f2(float a) { F2 r; r.v = make_float2(a, a); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ F2 neg(F2 a) { F2 r; r.v = make_float2(-a.v.x, -a.v.y); return r; }

}  // namespace tpx
