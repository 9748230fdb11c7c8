// FP32 (CUDA-core) forward-jet and reverse kernels, one persistent CTA set striding over
// tiles of P collocation points.  All layer activations of a tile live in shared memory;
// weights are streamed from L2 (fcn_common.cuh).  HBM traffic per point is the input row,
// the requested output streams and (reverse) the output adjoints.
#pragma once
#include "fcn_common.cuh"

namespace tpx {

template <class C>
constexpr size_t fwd_smem_bytes() {
  return sizeof(float) * (size_t)(C::BUF + C::WBUF + C::P * TPX_MAX_DIN);
}

template <class C>
__global__ void __launch_bounds__(NT) fcn_fwd_kernel(Layout lay, int activation, JetArgs jet,
                                                     const float* __restrict__ packed,
                                                     const float* __restrict__ x, int64_t n,
                                                     float* __restrict__ u, float* __restrict__ du,
                                                     float* __restrict__ lap) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;
  float* wbuf = act + C::BUF;
  float* xs = wbuf + C::WBUF;
  const int tid = threadIdx.x, tx = tid % C::NJ, ty = tid / C::NJ;
  const int64_t ntiles = (n + C::P - 1) / C::P;
  const int d_in = lay.d_in, d_out = lay.d_out;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * C::P;
    for (int v = tid; v < C::P * d_in; v += NT) {
      const int64_t g = base * d_in + v;
      xs[v] = (g < n * d_in) ? x[g] : 0.0f;
    }
    __syncthreads();
    forward_hidden<C, false>(lay, activation, jet, packed, xs, act, wbuf, nullptr, tx, ty, tid);
    // ---- output layer (linear in every stream): one dot product per (point, stream, output)
    for (int idx = tid; idx < C::P * C::S * d_out; idx += NT) {
      const int p = idx % C::P;
      const int s = (idx / C::P) % C::S;
      const int o = idx / (C::P * C::S);
      const int64_t gp = base + p;
      if (gp >= n) continue;
      const float* wl = packed + lay.wl() + (size_t)o * C::HP;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int i = 0; i < C::HP; i += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wl + i));
        a0 = fmaf(w.x, act[aidx<C>(s, i, p)], a0);
        a1 = fmaf(w.y, act[aidx<C>(s, i + 1, p)], a1);
        a2 = fmaf(w.z, act[aidx<C>(s, i + 2, p)], a2);
        a3 = fmaf(w.w, act[aidx<C>(s, i + 3, p)], a3);
      }
      const float a = (a0 + a1) + (a2 + a3);
      if (s == 0) {
        u[gp * d_out + o] = a + __ldg(packed + lay.bl() + o);
      } else if (s <= C::ND) {
        du[(gp * d_out + o) * C::ND + (s - 1)] = a;
      } else {
        lap[gp * d_out + o] = a;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// reverse kernel
// ------------------------------------------------------------------------------------
template <class C>
constexpr size_t bwd_smem_bytes() {
  // bufH, bufA, bufZ, weight stage, x tile, output adjoints [S][d_out<=?][P] (dynamic part
  // for the adjoints is added by the launcher: S * d_out * P floats)
  return sizeof(float) * (size_t)(3 * C::BUF + C::WBUF + C::P * TPX_MAX_DIN);
}

__device__ __forceinline__ void red_add(double* addr, float v) { atomicAdd(addr, (double)v); }

template <class C>
__global__ void __launch_bounds__(NT) fcn_bwd_kernel(Layout lay, int activation, JetArgs jet,
                                                     const float* __restrict__ packed,
                                                     const float* __restrict__ x, int64_t n,
                                                     const float* __restrict__ gu,
                                                     const float* __restrict__ gdu,
                                                     const float* __restrict__ glap,
                                                     double* __restrict__ gacc, float* __restrict__ ckpt_all) {
  constexpr int HP = C::HP, TP = C::TP, ND = C::ND, LAP = C::LAP, S = C::S, P = C::P;
  extern __shared__ __align__(16) float smem[];
  float* bufH = smem;
  float* bufA = bufH + C::BUF;
  float* bufZ = bufA + C::BUF;
  float* wbuf = bufZ + C::BUF;
  float* xs = wbuf + C::WBUF;
  float* og = xs + P * TPX_MAX_DIN;  // [S][d_out][P]
  const int tid = threadIdx.x, tx = tid % C::NJ, ty = tid / C::NJ;
  const int d_in = lay.d_in, d_out = lay.d_out, L = lay.L;
  float* ckpt = ckpt_all + (size_t)blockIdx.x * (size_t)L * C::BUF;
  const int64_t ntiles = (n + P - 1) / P;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * P;
    for (int v = tid; v < P * d_in; v += NT) {
      const int64_t g = base * d_in + v;
      xs[v] = (g < n * d_in) ? x[g] : 0.0f;
    }
    // output adjoints of this tile, zero for rows beyond n or absent streams
    for (int idx = tid; idx < S * d_out * P; idx += NT) {
      const int p = idx % P;
      const int o = (idx / P) % d_out;
      const int s = idx / (P * d_out);
      const int64_t gp = base + p;
      float v = 0.0f;
      if (gp < n) {
        if (s == 0) {
          if (gu) v = gu[gp * d_out + o];
        } else if (s <= ND) {
          if (gdu) v = gdu[(gp * d_out + o) * ND + (s - 1)];
        } else {
          if (glap) v = glap[gp * d_out + o];
        }
      }
      og[idx] = v;
    }
    __syncthreads();
    // ---- 1. forward recompute with checkpoints (per-CTA scratch, L2 resident)
    forward_hidden<C, true>(lay, activation, jet, packed, xs, bufH, wbuf, ckpt, tx, ty, tid);
    // bufH = streams of a_{L-1}
    // ---- 2. output layer: grads of WL, bL and adjoint of a_{L-1}
    for (int idx = tid; idx < d_out * HP; idx += NT) {
      const int i = idx % HP, o = idx / HP;
      float a = 0.0f;
      for (int s = 0; s < S; ++s) {
        const float* ogr = og + (s * d_out + o) * P;
        for (int p = 0; p < P; ++p) a = fmaf(ogr[p], bufH[aidx<C>(s, i, p)], a);
      }
      red_add(gacc + lay.wl() + (size_t)o * HP + i, a);
    }
    if (tid < d_out) {
      float a = 0.0f;
      for (int p = 0; p < P; ++p) a += og[(0 * d_out + tid) * P + p];
      red_add(gacc + lay.bl() + tid, a);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jn = tx * 4 + j;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float v[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) v[p] = 0.0f;
        for (int o = 0; o < d_out; ++o) {
          const float w = __ldg(packed + lay.wl() + (size_t)o * HP + jn);
#pragma unroll
          for (int p = 0; p < TP; ++p) v[p] = fmaf(w, og[(s * d_out + o) * P + ty * TP + p], v[p]);
        }
        store_row<C>(bufA, s, jn, ty, v);
      }
    }
    __syncthreads();
    // ---- 3. hidden layers, last to first
    for (int l = L - 1; l >= 0; --l) {
      // stage checkpoint of layer l (-> bufZ) and, for l > 0, of layer l-1 (-> bufH)
      {
        const float* src = ckpt + (size_t)l * C::BUF;
        for (int v = tid; v < C::BUF / 4; v += NT) cp_async16(bufZ + v * 4, src + v * 4);
        if (l > 0) {
          const float* src2 = ckpt + (size_t)(l - 1) * C::BUF;
          for (int v = tid; v < C::BUF / 4; v += NT) cp_async16(bufH + v * 4, src2 + v * 4);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
      // elementwise on this thread's own elements:
      //   bufA: adjoint of a_l  ->  adjoint of the pre-activation streams of layer l
      //   bufH: checkpoint of layer l-1  ->  a_{l-1} streams (input of layer l)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jn = tx * 4 + j;
        float keep[TP], hb[TP], lhb[TP], lz[TP];
        float s0[TP], s1[TP], s2[TP], s3[TP];
        load_row<C>(bufZ, 0, jn, ty, keep);
        load_row<C>(bufA, 0, jn, ty, hb);
#pragma unroll
        for (int p = 0; p < TP; ++p) act_from_keep(activation, keep[p], s0[p], s1[p], s2[p], s3[p]);
        float zb[TP], q[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) {
          zb[p] = hb[p] * s1[p];
          q[p] = 0.0f;
          lhb[p] = 0.0f;
        }
        if constexpr (LAP) {
          load_row<C>(bufA, 1 + ND, jn, ty, lhb);
          load_row<C>(bufZ, 1 + ND, jn, ty, lz);
        }
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          float dz[TP], dhb[TP], dzb[TP];
          load_row<C>(bufZ, 1 + k, jn, ty, dz);
          load_row<C>(bufA, 1 + k, jn, ty, dhb);
#pragma unroll
          for (int p = 0; p < TP; ++p) {
            zb[p] = fmaf(dhb[p] * s2[p], dz[p], zb[p]);
            dzb[p] = dhb[p] * s1[p];
            if constexpr (LAP) {
              q[p] = fmaf(jet.lap_w[k] * dz[p], dz[p], q[p]);
              dzb[p] = fmaf(2.0f * jet.lap_w[k] * lhb[p] * s2[p], dz[p], dzb[p]);
            }
          }
          store_row<C>(bufA, 1 + k, jn, ty, dzb);
        }
        if constexpr (LAP) {
          float lzb[TP];
#pragma unroll
          for (int p = 0; p < TP; ++p) {
            zb[p] = fmaf(lhb[p], fmaf(s3[p], q[p], s2[p] * lz[p]), zb[p]);
            lzb[p] = lhb[p] * s1[p];
          }
          store_row<C>(bufA, 1 + ND, jn, ty, lzb);
        }
        store_row<C>(bufA, 0, jn, ty, zb);
        if (l > 0) {
          // a_{l-1} from the checkpoint of layer l-1
          float k1[TP], t0[TP], t1[TP], t2[TP], t3[TP], qq[TP];
          load_row<C>(bufH, 0, jn, ty, k1);
#pragma unroll
          for (int p = 0; p < TP; ++p) {
            act_from_keep(activation, k1[p], t0[p], t1[p], t2[p], t3[p]);
            qq[p] = 0.0f;
          }
          store_row<C>(bufH, 0, jn, ty, t0);
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            float dz[TP], dh[TP];
            load_row<C>(bufH, 1 + k, jn, ty, dz);
#pragma unroll
            for (int p = 0; p < TP; ++p) {
              dh[p] = t1[p] * dz[p];
              if constexpr (LAP) qq[p] = fmaf(jet.lap_w[k] * dz[p], dz[p], qq[p]);
            }
            store_row<C>(bufH, 1 + k, jn, ty, dh);
          }
          if constexpr (LAP) {
            float lz1[TP], lh[TP];
            load_row<C>(bufH, 1 + ND, jn, ty, lz1);
#pragma unroll
            for (int p = 0; p < TP; ++p) lh[p] = fmaf(t2[p], qq[p], t1[p] * lz1[p]);
            store_row<C>(bufH, 1 + ND, jn, ty, lh);
          }
        }
      }
      __syncthreads();
      if (l > 0) {
        // bias gradient: sum over the tile's points of the value-stream adjoint
        if (tid < HP) {
          float a = 0.0f;
          for (int p = 0; p < P; ++p) a += bufA[aidx<C>(0, tid, p)];
          red_add(gacc + lay.b(l) + tid, a);
        }
        // weight gradient  dW_l[j][i] += sum_{s,p} zbar_s[j][p] * a_{l-1,s}[i][p]
        for (int blk = tid; blk < C::NJ * C::NJ; blk += NT) {
          const int ig = blk % C::NJ, jg = blk / C::NJ;
          float wacc[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) wacc[a][b] = 0.0f;
#pragma unroll 1
          for (int s = 0; s < S; ++s) {
#pragma unroll 2
            for (int g = 0; g < C::P4; ++g) {
              float4 zf[4], hf[4];
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                zf[a] = *reinterpret_cast<const float4*>(bufA + (s * HP + jg * 4 + a) * P + ((g ^ (jg & C::SWM)) << 2));
                hf[a] = *reinterpret_cast<const float4*>(bufH + (s * HP + ig * 4 + a) * P + ((g ^ (ig & C::SWM)) << 2));
              }
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  wacc[a][b] = fmaf(zf[a].x, hf[b].x, wacc[a][b]);
                  wacc[a][b] = fmaf(zf[a].y, hf[b].y, wacc[a][b]);
                  wacc[a][b] = fmaf(zf[a].z, hf[b].z, wacc[a][b]);
                  wacc[a][b] = fmaf(zf[a].w, hf[b].w, wacc[a][b]);
                }
            }
          }
          double* dst = gacc + lay.w(l) + (size_t)(jg * 4) * HP + ig * 4;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) red_add(dst + (size_t)a * HP + b, wacc[a][b]);
        }
        // adjoint of a_{l-1}:  abar_s[i][p] = sum_j W_l[j][i] zbar_s[j][p]
        float acc[S][TP][4];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int p = 0; p < TP; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][p][j] = 0.0f;
        gemm_layer<C>(acc, packed + lay.w(l), bufA, wbuf, tx, ty, tid);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float v[TP];
#pragma unroll
            for (int p = 0; p < TP; ++p) v[p] = acc[s][p][j];
            store_row<C>(bufA, s, tx * 4 + j, ty, v);
          }
        __syncthreads();
      } else {
        // layer 0: z = W0 x + b0, d_k z = W0[:, dcol k]
        for (int idx = tid; idx < HP * (d_in + 1); idx += NT) {
          const int jn = idx % HP, c = idx / HP;
          float a = 0.0f;
          if (c < d_in) {
            for (int p = 0; p < P; ++p) a = fmaf(bufA[aidx<C>(0, jn, p)], xs[p * d_in + c], a);
#pragma unroll
            for (int k = 0; k < ND; ++k)
              if (jet.dcol[k] == c)
                for (int p = 0; p < P; ++p) a += bufA[aidx<C>(1 + k, jn, p)];
            red_add(gacc + lay.w0() + (size_t)c * HP + jn, a);
          } else {
            for (int p = 0; p < P; ++p) a += bufA[aidx<C>(0, jn, p)];
            red_add(gacc + lay.b0() + jn, a);
          }
        }
        __syncthreads();
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// launchers (one translation unit per HP instantiates these)
// ------------------------------------------------------------------------------------
struct LaunchArgs {
  Layout lay;
  int activation;
  JetArgs jet;
  int lap;
  const float* packed;
  const float* x;
  int64_t n;
  float *u, *du, *lapo;
  const float *gu, *gdu, *glap;
  double* gacc;
  float* ckpt;
  size_t ckpt_bytes;
  int sm_count;
  cudaStream_t stream;
};

template <int HP>
int launch_fwd(const LaunchArgs& a);
template <int HP>
int launch_bwd(const LaunchArgs& a);
template <int HP>
int bwd_workspace(int nd, int lap, int L, int d_out, int sm_count, size_t* bytes);

template <int HP>
constexpr int fwd_tp() { return 4; }
template <int HP>
constexpr int bwd_tp() { return HP >= 256 ? 4 : 2; }

template <class C>
int do_launch_fwd(const LaunchArgs& a) {
  const size_t smem = fwd_smem_bytes<C>();
  if (smem > 227 * 1024) return TPX_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(fcn_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return TPX_E_CUDA;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fcn_fwd_kernel<C>, NT, smem);
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = (a.n + C::P - 1) / C::P;
  int64_t grid = (int64_t)a.sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
  fcn_fwd_kernel<C><<<(unsigned)grid, NT, smem, a.stream>>>(a.lay, a.activation, a.jet, a.packed, a.x, a.n, a.u, a.du, a.lapo);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

template <class C>
size_t bwd_total_smem(int d_out) {
  return bwd_smem_bytes<C>() + sizeof(float) * (size_t)C::S * d_out * C::P;
}

template <class C>
int bwd_grid(int d_out, int sm_count, size_t* smem_out) {
  const size_t smem = bwd_total_smem<C>(d_out);
  *smem_out = smem;
  if (smem > 227 * 1024) return -1;
  if (cudaFuncSetAttribute(fcn_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fcn_bwd_kernel<C>, NT, smem);
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

template <class C>
int do_launch_bwd(const LaunchArgs& a) {
  size_t smem;
  int grid = bwd_grid<C>(a.lay.d_out, a.sm_count, &smem);
  if (grid < 0) return TPX_E_UNSUPPORTED;
  const int64_t ntiles = (a.n + C::P - 1) / C::P;
  if (grid > ntiles) grid = (int)ntiles;
  if (grid < 1) return 0;
  const size_t need = (size_t)grid * a.lay.L * C::BUF * sizeof(float);
  if (a.ckpt_bytes < need) return TPX_E_WORKSPACE;
  fcn_bwd_kernel<C><<<(unsigned)grid, NT, smem, a.stream>>>(a.lay, a.activation, a.jet, a.packed, a.x, a.n, a.gu, a.gdu, a.glap, a.gacc, a.ckpt);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

template <class C>
int do_bwd_workspace(int L, int d_out, int sm_count, size_t* bytes) {
  size_t smem;
  int grid = bwd_grid<C>(d_out, sm_count, &smem);
  if (grid < 0) return TPX_E_UNSUPPORTED;
  *bytes = (size_t)grid * L * C::BUF * sizeof(float);
  return 0;
}

#define TPX_DISPATCH_JET(HPV, TPV, FN, ...)                                         \
  switch (nd * 2 + lap) {                                                           \
    case 0: return FN<Cfg<HPV, 0, 0, TPV>>(__VA_ARGS__);                             \
    case 2: return FN<Cfg<HPV, 1, 0, TPV>>(__VA_ARGS__);                             \
    case 3: return FN<Cfg<HPV, 1, 1, TPV>>(__VA_ARGS__);                             \
    case 4: return FN<Cfg<HPV, 2, 0, TPV>>(__VA_ARGS__);                             \
    case 5: return FN<Cfg<HPV, 2, 1, TPV>>(__VA_ARGS__);                             \
    case 6: return FN<Cfg<HPV, 3, 0, TPV>>(__VA_ARGS__);                             \
    case 7: return FN<Cfg<HPV, 3, 1, TPV>>(__VA_ARGS__);                             \
    default: return TPX_E_UNSUPPORTED;                                              \
  }

#define TPX_INSTANTIATE_HP(HPV)                                                      \
  template <>                                                                        \
  int launch_fwd<HPV>(const LaunchArgs& a) {                                         \
    const int nd = a.jet.nd, lap = a.lap;                                            \
    TPX_DISPATCH_JET(HPV, fwd_tp<HPV>(), do_launch_fwd, a)                           \
  }                                                                                  \
  template <>                                                                        \
  int launch_bwd<HPV>(const LaunchArgs& a) {                                         \
    const int nd = a.jet.nd, lap = a.lap;                                            \
    TPX_DISPATCH_JET(HPV, bwd_tp<HPV>(), do_launch_bwd, a)                           \
  }                                                                                  \
  template <>                                                                        \
  int bwd_workspace<HPV>(int nd, int lap, int L, int d_out, int sm_count, size_t* bytes) { \
    TPX_DISPATCH_JET(HPV, bwd_tp<HPV>(), do_bwd_workspace, L, d_out, sm_count, bytes) \
  }

}  // namespace tpx
