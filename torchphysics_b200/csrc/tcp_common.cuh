// Points-as-lanes tensor-core jet kernels (tcp_fwd.cu, tcp_bwd.cu): shared definitions.
//
// Round-2 design for hidden width <= 64 (the config-1 path).  A tile is 128 collocation points; TMEM lane = point, the
// streams (value, d_k, Laplacian) sit side by side in 64-column blocks.  The thread that reads lane p therefore owns every
// stream of point p for its 16 neurons: tanh, the jet formulas and their adjoints are thread-local and NOTHING is
// exchanged through shared memory (the round-1 kernels, rows = (stream, point), spent their shared-memory bandwidth on
// that exchange).  What makes the layout fit in 512 TMEM columns is a 16-bit operand format:
//
//   x * 2^e = hi + lo,  hi = fp16(x 2^e),  lo = fp16(x 2^e - hi)      (22 significant bits, as the tf32 split)
//   product = hi*hi + lo*hi + hi*lo  with kind::f16 MMAs (fp32 accumulate): half the operand bytes and twice the
//   rate of kind::tf32.  fp16 has a 5-bit exponent, so every operand carries a power-of-two scale 2^e chosen from a
//   rigorous BOUND of its magnitude (exact maxima of the previous stage, exchanged lazily, times operator norms of the
//   weights): no overflow by construction, conversions saturate anyway, and the fp16 subnormal range leaves > 2^15 of
//   slack for the looseness of the bounds.
//
// Measured on B200 (scripts/umma_probe3.cu): A from TMEM as two fp16 per 32-bit column (even K in the low half), B K-major
// SWIZZLE_128B: M=128 N=64 K=16 at its 32-cycle floor (2.17 PFLOP/s chip); A K-major + B MN-major and A MN-major +
// B MN-major (M = 128 as two 64-wide blocks LBO apart) from shared memory: 48 cycles at N = 64 (shared-memory port),
// 64 cycles at N = 128 (floor).  MN-major is what lets ONE image of row-major W serve the forward product (K-major B)
// and the adjoint product (MN-major B), and one image of Zbar (rows = points) serve the adjoint product (K-major A) and
// the gradient product (MN-major A, K = points).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace tpx {
namespace tcp {

using namespace tpx::tc;   // mbarrier / TMEM / fence helpers, tanh_fast

constexpr int HP = 64;              // padded hidden width
constexpr int TP = 128;             // points per tile = TMEM lanes
constexpr int CHN = 16;             // neurons per activation warp
constexpr int NCH = HP / CHN;       // neuron chunks = warps per TMEM quadrant
constexpr int ACT_WARPS = 4 * NCH;  // warp = chunk * 4 + quadrant
constexpr int THREADS = 32 * (ACT_WARPS + 1);
constexpr int MAXL = 8;             // hidden layers the image has scale slots for
constexpr int MAXIO = 4;

// ---- packed image (floats): [small fp32 block][per hidden matrix l = 1..L-1: W_hi, W_lo as fp16 SWIZZLE_128B images]
//   small: [W0: d_in x 64][bias: L x 64][WL: d_out x 64][bL: 8][norm: 4 x MAXL]  padded to 256 floats
//   norm(l): 0 = 1 / tau_l (tau_l = power of two that brings max|W_l| into [2^10, 2^11)), 1 = ||W_l||_inf (max row sum of |W|),
//            2 = ||W_l||_1 (max column sum), 3 = unused;  slot 0 holds (-, max|W0| per direction ...) see w0max()
//   image of W_l: row j (output neuron) = 128 bytes = W_l[j][0..63] * tau_l; 16-byte chunk c of row j sits at chunk
//   c ^ (j & 7) (the 128-byte swizzle); 8 KB per part = 2048 floats
struct Layout {
  int L, d_in, d_out;
  __host__ __device__ size_t w0() const { return 0; }
  __host__ __device__ size_t bias(int l) const { return (size_t)d_in * HP + (size_t)l * HP; }
  __host__ __device__ size_t wl() const { return (size_t)d_in * HP + (size_t)L * HP; }
  __host__ __device__ size_t bl() const { return wl() + (size_t)d_out * HP; }
  __host__ __device__ size_t norm(int l) const { return bl() + 8 + 4 * (size_t)l; }
  __host__ __device__ size_t w0max() const { return norm(MAXL); }              // max_j |W0[j][c]| per input column c (8 slots)
  __host__ __device__ size_t wlmax() const { return w0max() + 8; }             // max |WL| (1 slot, 8 reserved)
  __host__ __device__ size_t small() const { return (wlmax() + 8 + 255) / 256 * 256; }
  __host__ __device__ size_t mat(int l, int part) const { return small() + ((size_t)(l - 1) * 2 + part) * 2048; }
  __host__ __device__ size_t total() const { return small() + (size_t)(L - 1) * 2 * 2048; }
};

// ---- descriptors
// shared-memory operand, SWIZZLE_128B, 128-byte rows; K-major: rows = M/N index, 8-row groups SBO apart;
// MN-major: rows = K index (8-row groups SBO apart), 64-element blocks of the M/N index LBO apart
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {   // fp16 x fp16 -> fp32
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {  // bf16 x bf16 -> fp32
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// 16 consecutive columns without the wait (several loads in flight, one tmem_wait_ld() for all)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16f_nowait(uint32_t taddr, float (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                 "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// register reallocation between warpgroups (the activation warps need more than the 96 registers a 20-warp CTA starts with;
// the MMA warpgroup needs far fewer)
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- power-of-two scales from bounds
// scale 2^e with bound * 2^e in [2^14, 2^15) (fp16 max is 65504); returned as the biased exponent field of the scale
__device__ __forceinline__ int scale_exp_for(float bound) {
  const int E = (int)(__float_as_uint(bound) >> 23) & 0xFF;     // bound in [2^(E-127), 2^(E-126))
  int es = 268 - E;
  es = es < 1 ? 1 : (es > 253 ? 253 : es);
  return es;
}
__device__ __forceinline__ float exp_to_float(int biased) {     // 2^(biased - 127), flushes to 0 / saturates outside the normal range
  biased = biased < 0 ? 0 : (biased > 254 ? 254 : biased);
  return __uint_as_float((uint32_t)biased << 23);
}
__device__ __forceinline__ float inv_of_exp(int biased) { return exp_to_float(254 - biased); }

// two scaled values -> (hi, lo) fp16 pairs, element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half2 h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(*reinterpret_cast<uint32_t*>(&h)) : "f"(x1), "f"(x0));
  const float2 f = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// two values -> (hi, lo) bf16 pairs (x = hi + lo to 16 significant bits, fp32 exponent range), element 0 in the low half
__device__ __forceinline__ void split2_bf16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

// warp-wide maximum of non-negative floats in one instruction
__device__ __forceinline__ float warp_max_nonneg(float v) {
  uint32_t r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_uint(v)));
  return __uint_as_float(r);
}

// activation derivatives from the kept quantity (tanh keeps t, sin keeps z)
template <int ACT>
__device__ __forceinline__ void act_from_keep(float keep, float& t, float& s1, float& s2) {
  if constexpr (ACT == 0) {
    t = keep;
    s1 = fmaf(-t, t, 1.0f);
    s2 = -2.0f * t * s1;
  } else {
    sincosf(keep, &t, &s1);
    s2 = -t;
  }
}
template <int ACT> __host__ __device__ constexpr float s2_max() { return ACT == 0 ? 0.7698004f : 1.0f; }   // max |s''|: 4 / (3 sqrt 3)
template <int ACT> __host__ __device__ constexpr float s3_max() { return ACT == 0 ? 2.0f : 1.0f; }

// statistics the forward sweep leaves per tile and layer for the scales of the reverse sweep (exact tile maxima)
//   [0..3] max |a_s| of the activation streams, [4..6] max |d_k z|, [7] max |L z|
constexpr int STATS = 8;

}  // namespace tcp
}  // namespace tpx
