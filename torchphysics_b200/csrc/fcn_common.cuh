// Shared device code of the FP32 (CUDA-core) jet kernels: tiling configuration, the
// shared-memory activation layout, the weight-chunk pipeline and the activation jets.
//
// Mathematics (see oracle/jet_numpy.py): per layer z = W h + b is applied
// to every stream (value, d_k, L); through the activation s
//     h+ = s(z),  d_k h+ = s'(z) d_k z,  L h+ = s''(z) sum_k c_k (d_k z)^2 + s'(z) L z.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tpx.h"

namespace tpx {

constexpr int NT = 256;  // threads per CTA

// Packed parameter image (floats).  Hidden width is padded to HP for every layer.
//   [W0t: d_in x HP (column c of the input is row c)] [b0: HP]
//   for l = 1..L-1: [Wt_l: HP x HP, row = input neuron i, col = output neuron j]
//                   [W_l : HP x HP, row = output neuron j, col = input neuron i]
//                   [b_l : HP]
//   [WL: d_out x HP (row o)] [bL: d_out rounded up to 4]
struct Layout {
  int L, HP, d_in, d_out;
  __host__ __device__ size_t w0() const { return 0; }
  __host__ __device__ size_t b0() const { return (size_t)d_in * HP; }
  __host__ __device__ size_t hidden_base(int l) const {  // l in 1..L-1
    return (size_t)d_in * HP + HP + (size_t)(l - 1) * (2 * (size_t)HP * HP + HP);
  }
  // 9 or more outputs on a 128-wide image (DeepONet trunks; the DeepONet inner product folded into the output layer): the
  // output layer is stored like a hidden matrix -- Wt [HP][DOP] | W [DOP][HP] | b [DOP], DOP = d_out padded to 128-neuron
  // blocks, padding zero -- so that the layer-major tensor-core kernels treat it as matrix L, block by block
  __host__ __device__ bool wide_out() const { return HP >= 128 && d_out > 8; }
  __host__ __device__ size_t dop() const { return (size_t)((d_out + 127) / 128) * 128; }
  __host__ __device__ size_t wt(int l) const { return hidden_base(l); }
  __host__ __device__ size_t wts(int l) const { return (l == L && wide_out()) ? dop() : (size_t)HP; }   // row stride of wt(l)
  __host__ __device__ size_t w(int l) const { return l == L ? wl() : hidden_base(l) + (size_t)HP * HP; }
  __host__ __device__ size_t b(int l) const { return l == L ? bl() : hidden_base(l) + 2 * (size_t)HP * HP; }
  __host__ __device__ size_t wl() const { return wide_out() ? hidden_base(L) + (size_t)HP * dop() : hidden_base(L); }
  __host__ __device__ size_t bl() const { return wl() + (wide_out() ? dop() : (size_t)d_out) * HP; }
  __host__ __device__ size_t total() const { return bl() + (wide_out() ? dop() : (size_t)((d_out + 3) / 4) * 4); }
};

struct JetArgs {
  int nd;
  int dcol[TPX_MAX_ND];
  float lap_w[TPX_MAX_ND];
  int out_major;         // outputs and their adjoints as (d_out, n) / (nd, d_out, n): wide tensor-core path with wide_out only
  long long out_ld;      // ... their row stride = the caller's n (a launch may cover a chunk of the points)
};

template <int HP_, int ND_, int LAP_, int TP_>
struct Cfg {
  static constexpr int HP = HP_, ND = ND_, LAP = LAP_, TP = TP_;
  static constexpr int S = 1 + ND + LAP;   // streams: value, ND first derivatives, Laplacian
  static constexpr int NJ = HP / 4;        // neuron groups of 4
  static constexpr int NPG = NT / NJ;      // point groups
  static constexpr int P = TP * NPG;       // points per tile
  static constexpr int P4 = P / 4;
  static constexpr int SWM = (P4 - 1) < 7 ? (P4 - 1) : 7;  // xor-swizzle mask on 16 B groups
  static constexpr int BUF = S * HP * P;   // floats of one activation buffer
  static constexpr int KC = (HP >= 256) ? 8 : 16;  // weight rows per staged chunk
  static constexpr int WBUF = 2 * KC * HP; // floats of the double-buffered weight stage
  static_assert(NT % NJ == 0 && P % 4 == 0, "tile shape");
};

// activation buffer: element (stream s, neuron i, point p) of a tile.  Rows of P points;
// the 16-byte group index is xor-swizzled with the neuron group so that both "same
// neuron, consecutive points" and "consecutive neuron groups, same point group" accesses
// are bank-conflict free.
template <class C>
__device__ __forceinline__ int aidx(int s, int i, int p) {
  return (s * C::HP + i) * C::P + ((((p >> 2) ^ ((i >> 2) & C::SWM))) << 2) + (p & 3);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// s, s', s'' (and s''') of the activation from the quantity kept for the reverse sweep:
// tanh keeps t = tanh(z) (all derivatives are polynomials in t), sin keeps z.
__device__ __forceinline__ void act_eval(int act, float z, float& keep, float& s0, float& s1, float& s2) {
  if (act == TPX_ACT_TANH) {
    float t = tanhf(z);
    keep = t;
    s0 = t;
    s1 = fmaf(-t, t, 1.0f);
    s2 = -2.0f * t * s1;
  } else {
    float sn, cs;
    sincosf(z, &sn, &cs);
    keep = z;
    s0 = sn;
    s1 = cs;
    s2 = -sn;
  }
}
__device__ __forceinline__ void act_from_keep(int act, float keep, float& s0, float& s1, float& s2, float& s3) {
  if (act == TPX_ACT_TANH) {
    float t = keep;
    s0 = t;
    s1 = fmaf(-t, t, 1.0f);
    s2 = -2.0f * t * s1;
    s3 = -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f);
  } else {
    float sn, cs;
    sincosf(keep, &sn, &cs);
    s0 = sn;
    s1 = cs;
    s2 = -sn;
    s3 = -cs;
  }
}

// acc[s][p][j] += sum_i act[s][i][p] * Wg[i][j]  for this thread's TP points x 4 neurons.
// Wg: global, HP rows (reduction index) x HP columns, streamed through smem in chunks of
// KC rows with cp.async double buffering (weights stay L2-resident: every CTA reads them).
template <class C>
__device__ __forceinline__ void gemm_layer(float (&acc)[C::S][C::TP][4], const float* __restrict__ Wg,
                                           const float* act, float* wbuf, int tx, int ty, int tid) {
  constexpr int HP = C::HP, KC = C::KC, TP = C::TP;
  constexpr int CHUNK4 = KC * HP / 4;  // float4 per chunk
  constexpr int NCH = HP / KC;
  auto issue = [&](int c) {
    float* dst = wbuf + (c & 1) * (KC * HP);
    const float* src = Wg + (size_t)c * KC * HP;
#pragma unroll
    for (int v = tid; v < CHUNK4; v += NT) cp_async16(dst + v * 4, src + v * 4);
    cp_async_commit();
  };
  issue(0);
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * (KC * HP);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const int i = c * KC + kk;
      const float4 w = *reinterpret_cast<const float4*>(wb + kk * HP + tx * 4);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int s = 0; s < C::S; ++s) {
        float av[TP];
        if constexpr (TP == 4) {
          const float4 a = *reinterpret_cast<const float4*>(act + (s * HP + i) * C::P + ((ty ^ ((i >> 2) & C::SWM)) << 2));
          av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
        } else {
          const float2 a = *reinterpret_cast<const float2*>(act + (s * HP + i) * C::P + (((ty >> 1) ^ ((i >> 2) & C::SWM)) << 2) + (ty & 1) * 2);
          av[0] = a.x; av[1] = a.y;
        }
#pragma unroll
        for (int p = 0; p < TP; ++p)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[s][p][j] = fmaf(av[p], wv[j], acc[s][p][j]);
      }
    }
    __syncthreads();
  }
}

// store this thread's TP points of (stream s, neuron j) into an activation-layout buffer
template <class C>
__device__ __forceinline__ void store_row(float* buf, int s, int j, int ty, const float (&v)[C::TP]) {
  if constexpr (C::TP == 4) {
    *reinterpret_cast<float4*>(buf + (s * C::HP + j) * C::P + ((ty ^ ((j >> 2) & C::SWM)) << 2)) =
        make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *reinterpret_cast<float2*>(buf + (s * C::HP + j) * C::P + (((ty >> 1) ^ ((j >> 2) & C::SWM)) << 2) + (ty & 1) * 2) =
        make_float2(v[0], v[1]);
  }
}
template <class C>
__device__ __forceinline__ void load_row(const float* buf, int s, int j, int ty, float (&v)[C::TP]) {
  if constexpr (C::TP == 4) {
    const float4 a = *reinterpret_cast<const float4*>(buf + (s * C::HP + j) * C::P + ((ty ^ ((j >> 2) & C::SWM)) << 2));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    const float2 a = *reinterpret_cast<const float2*>(buf + (s * C::HP + j) * C::P + (((ty >> 1) ^ ((j >> 2) & C::SWM)) << 2) + (ty & 1) * 2);
    v[0] = a.x; v[1] = a.y;
  }
}

// Forward through all hidden layers for one tile.  On return `act` holds the streams of
// the last hidden layer's output.  If SAVE, the quantities the reverse sweep needs
// (keep = tanh(z) or z, d_k z, L z per hidden layer) go to `ckpt` (global, per CTA,
// L layers x BUF floats, same layout as the smem buffer).
template <class C, bool SAVE>
__device__ __forceinline__ void forward_hidden(const Layout& lay, int activation, const JetArgs& jet,
                                               const float* __restrict__ packed, const float* xs /*smem P x d_in*/,
                                               float* act, float* wbuf, float* ckpt, int tx, int ty, int tid) {
  constexpr int HP = C::HP, TP = C::TP, ND = C::ND, LAP = C::LAP, S = C::S;
  float acc[S][TP][4];
  // ---- layer 0: z = W0 x + b0, d_k z = W0[:, dcol k], L z = 0
  {
    const float4 bb = *reinterpret_cast<const float4*>(packed + lay.b0() + tx * 4);
    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int p = 0; p < TP; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[0][p][j] = bv[j];
    for (int c = 0; c < lay.d_in; ++c) {
      const float4 w = *reinterpret_cast<const float4*>(packed + lay.w0() + (size_t)c * HP + tx * 4);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int p = 0; p < TP; ++p) {
        const float xv = xs[(ty * TP + p) * lay.d_in + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][p][j] = fmaf(xv, wv[j], acc[0][p][j]);
      }
    }
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(packed + lay.w0() + (size_t)jet.dcol[k] * HP + tx * 4);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int p = 0; p < TP; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[1 + k][p][j] = wv[j];
    }
    if constexpr (LAP) {
#pragma unroll
      for (int p = 0; p < TP; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[1 + ND][p][j] = 0.0f;
    }
  }
#pragma unroll 1
  for (int l = 0; l < lay.L; ++l) {
    if (l > 0) {
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int p = 0; p < TP; ++p)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[s][p][j] = 0.0f;
      gemm_layer<C>(acc, packed + lay.wt(l), act, wbuf, tx, ty, tid);
      const float4 bb = *reinterpret_cast<const float4*>(packed + lay.b(l) + tx * 4);
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int p = 0; p < TP; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][p][j] += bv[j];
    }
    // ---- activation jets, written in place (all reads of `act` for this layer are done:
    //      gemm_layer ends with __syncthreads)
    float* ck = SAVE ? ckpt + (size_t)l * C::BUF : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jn = tx * 4 + j;
      float hv[TP], kv[TP], s1v[TP], s2v[TP];
#pragma unroll
      for (int p = 0; p < TP; ++p) {
        float s0;
        act_eval(activation, acc[0][p][j], kv[p], s0, s1v[p], s2v[p]);
        hv[p] = s0;
      }
      store_row<C>(act, 0, jn, ty, hv);
      if constexpr (SAVE) store_row<C>(ck, 0, jn, ty, kv);
      float q[TP];
#pragma unroll
      for (int p = 0; p < TP; ++p) q[p] = 0.0f;
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        float dv[TP], zv[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) {
          zv[p] = acc[1 + k][p][j];
          dv[p] = s1v[p] * zv[p];
          if constexpr (LAP) q[p] = fmaf(jet.lap_w[k] * zv[p], zv[p], q[p]);
        }
        store_row<C>(act, 1 + k, jn, ty, dv);
        if constexpr (SAVE) store_row<C>(ck, 1 + k, jn, ty, zv);
      }
      if constexpr (LAP) {
        float lv[TP], zv[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) {
          zv[p] = acc[1 + ND][p][j];
          lv[p] = fmaf(s2v[p], q[p], s1v[p] * zv[p]);
        }
        store_row<C>(act, 1 + ND, jn, ty, lv);
        if constexpr (SAVE) store_row<C>(ck, 1 + ND, jn, ty, zv);
      }
    }
    __syncthreads();
  }
}

}  // namespace tpx
