// Points-as-lanes tensor-core forward-jet kernel (hidden width <= 64, tanh / sin, <= 4 streams); see tcp_common.cuh.
//
// One persistent CTA per SM, 16 activation warps + 1 MMA warp, one 128-point tile in flight.  TMEM holds two regions
// X | Y of S x 64 columns.  Per hidden matrix l:
//     D_l[p][j] (region R')  =  sum_i A_{l-1}[p][i] W_l[j][i]        per stream: M = 128 points, N = 64, K = 64
// with A_{l-1} = (hi | lo) fp16 pairs written IN PLACE over the columns of D_{l-1} the same thread has just consumed
// (chunk c of stream s: 16 fp32 columns -> 8 columns hi + 8 columns lo), so a layer step is
//     tcgen05.ld  ->  registers: scale, bias, tanh, jets, checkpoint stores, scale, split  ->  tcgen05.st
// and the only shared-memory traffic of the step is the tensor core reading W.  The MMA warp starts the products of
// neuron chunk c (one K = 16 step of every stream) as soon as the four quadrant warps of that chunk have arrived, so
// most of the tensor-core time hides behind the activation work of the other chunks.
//
// Scales: the operand of stream s of point p is scaled by 2^e with e from a bound of the row's magnitude:
//   value stream |tanh|, |sin| <= 1; d_k stream  <= ||W_l||_inf * (measured row maximum of the previous layer's stream),
//   Laplacian stream <= max|s''| sum |c_k| (bound of d_k z)^2 + ||W_l||_inf * (previous row maximum);  layer 0: static
//   bounds from W0.  The four warps of a quadrant exchange their partial row maxima through shared memory one layer
//   ahead of their use (no barrier: the mbarrier chain of the MMA orders them).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_kernels.h"
#include "tcp_common.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tcp {

struct PackArgs {
  const float* W[TPX_MAX_HIDDEN + 1];
  const float* b[TPX_MAX_HIDDEN + 1];
  int in_w[TPX_MAX_HIDDEN + 1];
  int out_w[TPX_MAX_HIDDEN + 1];
};

// block l (0 <= l <= L): norms of W_l
__global__ void tcp_norm_kernel(Layout lay, PackArgs pa, float* __restrict__ img) {
  const int l = blockIdx.x, L = lay.L;
  const int ow = pa.out_w[l], iw = pa.in_w[l];
  __shared__ float s_max, s_inf, s_one, s_col[8];
  if (threadIdx.x == 0) { s_max = 0.f; s_inf = 0.f; s_one = 0.f; }
  if (threadIdx.x < 8) s_col[threadIdx.x] = 0.f;
  __syncthreads();
  const float* W = pa.W[l];
  // rows (output neurons): max row sum, max element; columns: max column sum
  for (int j = threadIdx.x; j < ow; j += blockDim.x) {
    float rs = 0.f, mx = 0.f;
    for (int i = 0; i < iw; ++i) { const float w = fabsf(W[(size_t)j * iw + i]); rs += w; mx = fmaxf(mx, w); }
    atomicMax(reinterpret_cast<int*>(&s_inf), __float_as_int(rs));
    atomicMax(reinterpret_cast<int*>(&s_max), __float_as_int(mx));
  }
  for (int i = threadIdx.x; i < iw; i += blockDim.x) {
    float cs = 0.f, mx = 0.f;
    for (int j = 0; j < ow; ++j) { const float w = fabsf(W[(size_t)j * iw + i]); cs += w; mx = fmaxf(mx, w); }
    atomicMax(reinterpret_cast<int*>(&s_one), __float_as_int(cs));
    if (l == 0 && i < 8) s_col[i] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (l >= 1 && l < L) {
      // tau: max|W| * tau in [2^10, 2^11)
      const int E = (int)(__float_as_uint(s_max) >> 23) & 0xFF;
      int es = 127 + 10 - (E - 127);
      es = es < 1 ? 1 : (es > 253 ? 253 : es);
      if (s_max == 0.f) es = 127;
      img[lay.norm(l) + 0] = __uint_as_float((uint32_t)(254 - es) << 23);   // 1 / tau
      img[lay.norm(l) + 1] = s_inf * 1.0001f;
      img[lay.norm(l) + 2] = s_one * 1.0001f;
      img[lay.norm(l) + 3] = __uint_as_float((uint32_t)es << 23);           // tau
    }
    if (l == 0)
      for (int c = 0; c < 8; ++c) img[lay.w0max() + c] = s_col[c];
    if (l == L) img[lay.wlmax()] = s_max;
  }
}

// one thread per logical element; padding stays zero (image zeroed by the launcher)
__global__ void tcp_fill_kernel(Layout lay, PackArgs pa, float* __restrict__ img) {
  const int L = lay.L;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < lay.d_in * HP; e += gridDim.x * blockDim.x) {
    const int c = e / HP, j = e % HP;
    if (j < pa.out_w[0]) img[lay.w0() + e] = pa.W[0][(size_t)j * pa.in_w[0] + c];
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L * HP; e += gridDim.x * blockDim.x) {
    const int l = e / HP, j = e % HP;
    if (j < pa.out_w[l]) img[lay.bias(0) + e] = pa.b[l][j];
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < lay.d_out * HP; e += gridDim.x * blockDim.x) {
    const int o = e / HP, j = e % HP;
    if (j < pa.in_w[L]) img[lay.wl() + e] = pa.W[L][(size_t)o * pa.in_w[L] + j];
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < lay.d_out && e < 8; e += gridDim.x * blockDim.x) img[lay.bl() + e] = pa.b[L][e];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < (L - 1) * HP * HP; e += gridDim.x * blockDim.x) {
    const int l = 1 + e / (HP * HP), j = (e / HP) % HP, i = e % HP;
    float w = 0.0f;
    if (j < pa.out_w[l] && i < pa.in_w[l]) w = pa.W[l][(size_t)j * pa.in_w[l] + i];
    const float ws = w * img[lay.norm(l) + 3];
    const __half h = __float2half_rn(ws);
    const __half lo = __float2half_rn(ws - __half2float(h));
    const int off = j * 64 + ((((i >> 3) ^ (j & 7))) << 3) + (i & 7);       // fp16 elements
    reinterpret_cast<__half*>(img + lay.mat(l, 0))[off] = h;
    reinterpret_cast<__half*>(img + lay.mat(l, 1))[off] = lo;
  }
}

// ------------------------------------------------------------------------------------------
struct FwdSmem {
  int wimg, small, mx, part, st, total;   // byte offsets from the 1 KB aligned base
};
constexpr int OC = 2;   // output columns per pass of the output layer
__host__ __device__ inline FwdSmem fwd_smem(const Layout& lay) {
  FwdSmem s;
  s.wimg = 0;
  s.small = (lay.L - 1) * 16384;
  s.mx = s.small + (int)lay.small() * 4;
  s.part = s.mx + 2 * ACT_WARPS * 3 * 32 * 4;
  s.st = s.part + ACT_WARPS * 4 * OC * 32 * 4;
  s.total = s.st + MAXL * STATS * 4;
  return s;
}

struct FwdArgs {
  Layout lay;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  float *u, *du, *lap;
  float* ckpt;     // optional: [tile][l = 1..L-1][s][64 neurons][128 points]: tanh z (sin: z), d_k z, L z
  float* stats;    // with ckpt: [tile][L][STATS] tile maxima for the scales of the reverse sweep
};

template <int ND, int LAP, int ACT>
__global__ void __launch_bounds__(THREADS, 1) tcp_fwd_kernel(FwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  constexpr uint32_t RY = 256;               // TMEM column of region Y (X = 0)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a[NCH], bar_d;
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Layout lay = a.lay;
  const FwdSmem so = fwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* small = reinterpret_cast<float*>(sm + so.small);

  {
    const int nv = (L - 1) * 1024;            // 16-byte vectors of the weight images
    const float4* src = reinterpret_cast<const float4*>(a.img + lay.mat(1, 0));
    float4* dst = reinterpret_cast<float4*>(sm + so.wimg);
    for (int v = tid; v < nv; v += THREADS) dst[v] = __ldg(src + v);
    for (int v = tid; v < (int)lay.small(); v += THREADS) small[v] = __ldg(a.img + v);
    for (int v = tid; v < MAXL * STATS; v += THREADS) reinterpret_cast<uint32_t*>(sm + so.st)[v] = 0u;
  }
  if (warp == ACT_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int c = 0; c < NCH; ++c) mbar_init(&bar_a[c], 4);
    mbar_init(&bar_d, 1);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TP - 1) / TP;

  if (warp == ACT_WARPS) {
    // =========================== MMA issuer ===========================
    const uint32_t wbase = smem_u32(sm + so.wimg);
    constexpr uint32_t idesc = idesc_f16(128, HP, 0, 0);
    uint32_t pa = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int l = 1; l < L; ++l) {
        const uint32_t src = tmem + (((l - 1) & 1) ? RY : 0u), dst = tmem + (((l - 1) & 1) ? 0u : RY);
        const uint32_t wh = wbase + (uint32_t)(l - 1) * 16384u, wl = wh + 8192u;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&bar_a[c], pa);
          fence_after();
          if (elect_one()) {
            const uint64_t dh = desc_sw128(wh + c * 32, 16, 1024), dl = desc_sw128(wl + c * 32, 16, 1024);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const uint32_t ah = src + s * 64 + c * 16, al = ah + 8, d = dst + s * 64;
              mma_f16_ts(d, al, dh, idesc, c > 0);
              mma_f16_ts(d, ah, dl, idesc, 1);
              mma_f16_ts(d, ah, dh, idesc, 1);
            }
            if (c == NCH - 1) mma_commit(&bar_d);
          }
          __syncwarp();
        }
        pa ^= 1;
      }
    }
  } else {
    // =========================== activation warps ===========================
    const int quad = warp & 3, chunk = warp >> 2;
    const int j0 = chunk * CHN;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float* W0 = small + lay.w0();
    const float* WL = small + lay.wl();
    const float* bL = small + lay.bl();
    float* mx = reinterpret_cast<float*>(sm + so.mx);            // [buf][warp][3][32]
    float* part = reinterpret_cast<float*>(sm + so.part);        // [warp][4][OC][32]
    uint32_t* st = reinterpret_cast<uint32_t*>(sm + so.st);      // [L][STATS]
    uint32_t ph_d = 0;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    const bool save = a.ckpt != nullptr;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gp = tile * TP + quad * 32 + lane;
      const bool valid = gp < a.n;
      float xr[TPX_MAX_DIN];
#pragma unroll
      for (int cc = 0; cc < TPX_MAX_DIN; ++cc) xr[cc] = (valid && cc < d_in) ? __ldg(a.x + gp * d_in + cc) : 0.0f;
      float inv[S];                              // 1 / (operand scale * tau) of the product about to be read
      float* ck_tile = save ? a.ckpt + (size_t)tile * ((size_t)(L - 1) * S * HP * TP) + (size_t)j0 * TP + quad * 32 + lane : nullptr;

      for (int l = 0; l < L; ++l) {
        const uint32_t R = tmem + ((l & 1) ? RY : 0u) + lane_base;   // D_l arrives here and A_l is written here
        const float* bias = small + lay.bias(l) + j0;
        float z[S][CHN];
        if (l == 0) {
#pragma unroll
          for (int i = 0; i < CHN; ++i) z[0][i] = bias[i];
#pragma unroll
          for (int cc = 0; cc < TPX_MAX_DIN; ++cc)
            if (cc < d_in) {
#pragma unroll
              for (int i = 0; i < CHN; ++i) z[0][i] = fmaf(W0[cc * HP + j0 + i], xr[cc], z[0][i]);
            }
#pragma unroll
          for (int k = 0; k < ND; ++k)
#pragma unroll
            for (int i = 0; i < CHN; ++i) z[1 + k][i] = W0[dcol[k] * HP + j0 + i];
          if (LAP) {
#pragma unroll
            for (int i = 0; i < CHN; ++i) z[S - 1][i] = 0.0f;
          }
        } else {
          mbar_wait(&bar_d, ph_d);
          ph_d ^= 1;
          fence_after();
#pragma unroll
          for (int s = 0; s < S; ++s) tmem_ld16f_nowait(R + s * 64 + j0, z[s]);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < CHN; ++i) z[0][i] = fmaf(z[0][i], inv[0], bias[i]);
#pragma unroll
          for (int s = 1; s < S; ++s)
#pragma unroll
            for (int i = 0; i < CHN; ++i) z[s][i] *= inv[s];
        }
        // ---- bounds of this layer's activation streams (rows of the next operand), from the previous layer's maxima
        float bd[S];
        bd[0] = 1.0f;
        {
          float bz[S];                           // bounds of |d_k z|, |L z| over the whole row
          if (l == 0) {
#pragma unroll
            for (int k = 0; k < ND; ++k) bz[1 + k] = small[lay.w0max() + dcol[k]];
            if (LAP) bz[S - 1] = 0.0f;
          } else {
            const float winf = small[lay.norm(l) + 1];
            const float* mprev = mx + (((l - 1) & 1) * ACT_WARPS + quad) * 96 + lane;     // chunk c' at + c' * 4 * 96
#pragma unroll
            for (int s = 1; s < S; ++s) {
              float m = mprev[(s - 1) * 32];
#pragma unroll
              for (int c2 = 1; c2 < NCH; ++c2) m = fmaxf(m, mprev[c2 * 4 * 96 + (s - 1) * 32]);
              bz[s] = winf * m;
            }
          }
          float qb = 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            bd[1 + k] = bz[1 + k];
            qb = fmaf(fabsf(lapw[k]) * bz[1 + k], bz[1 + k], qb);
          }
          if (LAP) bd[S - 1] = fmaf(s2_max<ACT>(), qb, bz[S - 1]);
        }
        // ---- thread-local jets of 16 neurons (all streams of this point)
        float* ckl = (save && l >= 1) ? ck_tile + (size_t)(l - 1) * S * HP * TP : nullptr;
        float mz[S], ma[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { mz[s] = 0.0f; ma[s] = 0.0f; }
#pragma unroll
        for (int i = 0; i < CHN; ++i) {
          float t, s1, s2, keep;
          if constexpr (ACT == 0) {
            t = tanh_fast(z[0][i]);
            s1 = fmaf(-t, t, 1.0f);
            s2 = -2.0f * t * s1;
            keep = t;
          } else {
            keep = z[0][i];
            sincosf(keep, &t, &s1);
            s2 = -t;
          }
          z[0][i] = t;
          if (ckl) __stcs(ckl + (0 * HP + i) * TP, keep);
          float qq = 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const float dz = z[1 + k][i];
            if (LAP) qq = fmaf(lapw[k] * dz, dz, qq);
            const float v = s1 * dz;
            z[1 + k][i] = v;
            mz[1 + k] = fmaxf(mz[1 + k], fabsf(dz));
            ma[1 + k] = fmaxf(ma[1 + k], fabsf(v));
            if (ckl) __stcs(ckl + ((1 + k) * HP + i) * TP, dz);
          }
          if (LAP) {
            const float lz = z[S - 1][i];
            const float v = fmaf(s2, qq, s1 * lz);
            z[S - 1][i] = v;
            mz[S - 1] = fmaxf(mz[S - 1], fabsf(lz));
            ma[S - 1] = fmaxf(ma[S - 1], fabsf(v));
            if (ckl) __stcs(ckl + ((S - 1) * HP + i) * TP, lz);
          }
        }
        if (a.stats) {                           // tile maxima for the reverse sweep
#pragma unroll
          for (int s = 1; s < S; ++s) {
            const float wa = warp_max_nonneg(ma[s]), wz = warp_max_nonneg(mz[s]);
            if (lane == 0) {
              atomicMax(st + l * STATS + s, __float_as_uint(wa));
              atomicMax(st + l * STATS + 4 + (s - 1), __float_as_uint(wz));
            }
          }
        }
        if (l < L - 1) {
          // ---- next operand: scale, split, store in place; publish this chunk's row maxima for the next layer's bounds
          float* mown = mx + ((l & 1) * ACT_WARPS + warp) * 96 + lane;
#pragma unroll
          for (int s = 1; s < S; ++s) mown[(s - 1) * 32] = ma[s];
          const float itau = small[lay.norm(l + 1) + 0];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int es = scale_exp_for(bd[s]);
            const float sc = exp_to_float(es);
            inv[s] = inv_of_exp(es) * itau;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) split2(z[s][2 * v] * sc, z[s][2 * v + 1] * sc, hi[v], lo[v]);
            tmem_st8u(R + s * 64 + j0, hi);
            tmem_st8u(R + s * 64 + j0 + 8, lo);
          }
          tmem_wait_st();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a[chunk]);
        } else if (a.u != nullptr) {
          // ---- output layer: partial dot products over this warp's 16 neurons, met in shared memory by the 4 chunk warps
          // of the quadrant; warp (quad, chunk = s) finishes stream s
          for (int o0 = 0; o0 < d_out; o0 += OC) {
#pragma unroll
            for (int oo = 0; oo < OC; ++oo) {
              const int o = (o0 + oo < d_out) ? o0 + oo : o0;
#pragma unroll
              for (int s = 0; s < S; ++s) {
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < CHN; ++i) acc = fmaf(WL[o * HP + j0 + i], z[s][i], acc);
                part[((warp * 4 + s) * OC + oo) * 32 + lane] = acc;
              }
            }
            named_bar_sync(1 + quad, 128);
            if (chunk < S && valid) {
              const int s = chunk;
#pragma unroll
              for (int oo = 0; oo < OC; ++oo) {
                const int o = o0 + oo;
                if (o < d_out) {
                  float r = 0.0f;
#pragma unroll
                  for (int c2 = 0; c2 < NCH; ++c2) r += part[(((c2 * 4 + quad) * 4 + s) * OC + oo) * 32 + lane];
                  if (s == 0) a.u[gp * d_out + o] = r + bL[o];
                  else if (s <= ND) a.du[(gp * d_out + o) * ND + (s - 1)] = r;
                  else a.lap[gp * d_out + o] = r;
                }
              }
            }
            if (o0 + OC < d_out) named_bar_sync(1 + quad, 128);
          }
        }
      }
      named_bar_sync(1 + quad, 128);          // the quadrant's row-maximum exchange buffers are reused by the next tile
      if (a.stats) {
        // every warp has contributed to st (the last layer's maxima precede this barrier); one warp writes them out and
        // clears the slots for the next tile
        named_bar_sync(5, 32 * ACT_WARPS);
        if (warp == 0) {
          for (int v = lane; v < L * STATS; v += 32) {
            a.stats[(size_t)tile * (L * STATS) + v] = __uint_as_float(st[v]);
            st[v] = 0u;
          }
        }
        named_bar_sync(5, 32 * ACT_WARPS);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == ACT_WARPS) tmem_dealloc(tmem, 512);
}

static int env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

}  // namespace tcp

using namespace tcp;

bool tcp_net_supported(const tpx_fcn_desc* net) {
  // A/B switch (development): the points-as-lanes forward kernel and the split-fp16 reverse kernel are correct
  // (tests/test_gpu_fcn_jet.py runs them) but do not beat the round-1 kernels yet (profiles/README.md r02a-r02c)
  static const int enabled = tcp::env_flag("TPX_TCP");
  if (!enabled) return false;
  if (net->activation != TPX_ACT_TANH && net->activation != TPX_ACT_SIN) return false;
  if (net->n_hidden < 2 || net->n_hidden > tcp::MAXL) return false;
  for (int l = 0; l < net->n_hidden; ++l)
    if (net->widths[l] > tcp::HP) return false;
  if (net->d_in > tcp::MAXIO || net->d_out > tcp::MAXIO) return false;
  const tcp::Layout lay{net->n_hidden, net->d_in, net->d_out};
  if ((size_t)fwd_smem(lay).total + 1024 > 227 * 1024) return false;
  return true;
}

bool tcp_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tcp_net_supported(net)) return false;
  if (1 + nd + lap > 4) return false;
  if (lap && nd == 0) return false;
  return true;
}

size_t tcp_packed_floats(const tpx_fcn_desc* net) {
  tcp::Layout lay{net->n_hidden, net->d_in, net->d_out};
  return lay.total();
}

int tcp_pack(const tpx_fcn_desc* net, const void* const* params_host, float* img, cudaStream_t stream) {
  tcp::Layout lay{net->n_hidden, net->d_in, net->d_out};
  tcp::PackArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int l = 0; l <= net->n_hidden; ++l) {
    pa.in_w[l] = (l == 0) ? net->d_in : net->widths[l - 1];
    pa.out_w[l] = (l == net->n_hidden) ? net->d_out : net->widths[l];
    pa.W[l] = (const float*)params_host[2 * l];
    pa.b[l] = (const float*)params_host[2 * l + 1];
  }
  if (cudaMemsetAsync(img, 0, lay.total() * sizeof(float), stream) != cudaSuccess) return TPX_E_CUDA;
  tcp_norm_kernel<<<net->n_hidden + 1, 64, 0, stream>>>(lay, pa, img);
  tcp_fill_kernel<<<64, 256, 0, stream>>>(lay, pa, img);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

size_t tcp_ckpt_floats_per_tile(int L, int S) { return (size_t)(L - 1) * S * tcp::HP * tcp::TP; }

template <int ND, int LAP, int ACT>
static int launch_tcp_fwd_act(const FwdArgs& a, int sm_count, cudaStream_t stream) {
  const size_t smem = (size_t)fwd_smem(a.lay).total + 1024;
  if (smem > 227 * 1024) return TPX_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(tcp_fwd_kernel<ND, LAP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tcp_fwd_kernel: shared memory attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
  const int64_t ntiles = (a.n + TP - 1) / TP;
  int64_t grid = sm_count < ntiles ? sm_count : ntiles;
  if (grid < 1) return 0;
  tcp_fwd_kernel<ND, LAP, ACT><<<(unsigned)grid, THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tcp_fwd_kernel launch (%lld CTAs, %zu bytes smem): %s", (long long)grid, smem, cudaGetErrorString(e));
  return 0;
}
template <int ND, int LAP>
static int launch_tcp_fwd(const FwdArgs& a, int act, int sm_count, cudaStream_t stream) {
  return act == TPX_ACT_SIN ? launch_tcp_fwd_act<ND, LAP, 1>(a, sm_count, stream) : launch_tcp_fwd_act<ND, LAP, 0>(a, sm_count, stream);
}

int tcp_forward(const tpx_fcn_desc* net, const JetArgs& jet, int lap, const float* img, const float* x, int64_t n,
                float* u, float* du, float* lapo, float* ckpt, float* stats, int sm_count, cudaStream_t stream) {
  FwdArgs a;
  a.lay = tcp::Layout{net->n_hidden, net->d_in, net->d_out};
  a.jet = jet;
  a.img = img; a.x = x; a.n = n; a.u = u; a.du = du; a.lap = lapo;
  a.ckpt = ckpt; a.stats = stats;
  switch (jet.nd * 2 + lap) {
    case 0: return launch_tcp_fwd<0, 0>(a, net->activation, sm_count, stream);
    case 2: return launch_tcp_fwd<1, 0>(a, net->activation, sm_count, stream);
    case 3: return launch_tcp_fwd<1, 1>(a, net->activation, sm_count, stream);
    case 4: return launch_tcp_fwd<2, 0>(a, net->activation, sm_count, stream);
    case 5: return launch_tcp_fwd<2, 1>(a, net->activation, sm_count, stream);
    case 6: return launch_tcp_fwd<3, 0>(a, net->activation, sm_count, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

}  // namespace tpx
