// Tensor-core (tcgen05 / TMEM) forward-jet kernel for hidden width <= 64, tanh, <= 4 streams.
//
// Work decomposition ("R32"): a tile is 32 collocation points.  The 128 rows of one UMMA are
// (stream q, point p) = TMEM lane 32 q + p: every stream (value, d_k, Laplacian) of every point
// is one row of the SAME weight matrix product, so one M=128 N=64 K=64 product per layer serves a
// whole tile.  Per layer:   D[128 x 64] = A[128 x 64] W^T   with fp32-grade accuracy from three
// tf32 products  A_hi W_hi + A_hi W_lo + A_lo W_hi  (hi = round-to-nearest tf32, lo = exact rest).
//   * A_hi / A_lo live in TMEM (written by the activation warps with tcgen05.st), W_hi / W_lo in
//     shared memory (128-byte swizzle, K-major), D in TMEM: the MMA runs at its 32-cycle floor.
//   * warp q of a tile slot owns TMEM quadrant q = stream q.  The activation jets couple the
//     streams of a point (d_k h = s'(z) d_k z, L h = s''(z) sum c_k (d_k z)^2 + s'(z) L z), so the
//     value warp publishes tanh(z) and the derivative warps publish d_k z through shared memory.
//   * two tile slots per CTA ping-pong: while the tensor core works on one tile's layer, the
//     activation warps of the other tile run.  One persistent CTA per SM; all weights resident.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_common.cuh"
#include "tc_kernels.h"
#include "tc_layout.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tc {

constexpr int SLOTS = 2;           // tiles in flight per CTA
constexpr int HALVES = 2;          // warps per TMEM quadrant (each takes HP / HALVES neurons)
constexpr int NB = HP / HALVES;    // neurons per activation warp
constexpr int SLOT_WARPS = 4 * HALVES;
constexpr int EPI_WARPS = SLOT_WARPS * SLOTS;
constexpr int THREADS = 32 * (EPI_WARPS + 1);
constexpr int CH = 16;             // neurons per register chunk
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SLOT_COLS = 192;  // D | A_hi | A_lo

// tanh|x| = (1 - e) / (1 + e), e = e^{-2|x|}: ex2.approx (2^-22 relative) and rcp.approx refined by one Newton step;
// absolute error <= 3e-7, which is what the next layer's sums see (measured parity: tests/test_gpu_fcn_jet.py)
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -2.885390081777927f));   // e^{-2|x|} in (0, 1]
  const float d = e + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = r * fmaf(-d, r, 2.0f);
  return copysignf((1.0f - e) * r, x);
}

struct TcPackArgs {
  const float* W[TPX_MAX_HIDDEN + 1];
  const float* b[TPX_MAX_HIDDEN + 1];
  int in_w[TPX_MAX_HIDDEN + 1];
  int out_w[TPX_MAX_HIDDEN + 1];
};

// one thread per logical parameter element; padding stays zero (image zeroed by the launcher with cudaMemsetAsync)
__global__ void tc_pack_fill_kernel(TcLayout lay, TcPackArgs pa, float* __restrict__ img) {
  const int L = lay.L;
  // W0, biases, WL, bL
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
  // hidden matrices
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < (L - 1) * HP * HP; e += gridDim.x * blockDim.x) {
    const int l = 1 + e / (HP * HP), j = (e / HP) % HP, i = e % HP;
    float w = 0.0f;
    if (j < pa.out_w[l] && i < pa.in_w[l]) w = pa.W[l][(size_t)j * pa.in_w[l] + i];
    uint32_t u = __float_as_uint(w);
    // round to nearest (ties away) tf32, same as cvt.rna.tf32.f32 for finite values
    uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
    const float hi = __uint_as_float(h), lo = w - hi;
    img[lay.mat(l, 0) + swz128(j, i, HP)] = hi;
    img[lay.mat(l, 1) + swz128(j, i, HP)] = lo;
    img[lay.mat(l, 2) + swz128(i, j, HP)] = hi;
    img[lay.mat(l, 3) + swz128(i, j, HP)] = lo;
  }
}

// ------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------
struct FwdSmem {
  // offsets in floats from the 1024-byte aligned base
  int mats;      // (L-1) * 2 * 4096   W_hi, W_lo per hidden layer
  int small;     // copy of the small block
  int ex;        // SLOTS * 3 * HP * TILE exchange buffers
  int part;      // SLOTS * 4 * 8 * TILE partial output sums
  int total;
};
__host__ __device__ inline FwdSmem fwd_smem(const TcLayout& lay) {
  FwdSmem s;
  s.mats = 0;
  s.small = (lay.L - 1) * 2 * 4096;
  s.ex = s.small + (int)lay.small();
  s.part = s.ex + SLOTS * 3 * HP * TILE;
  s.total = s.part + SLOTS * 4 * 8 * TILE;
  return s;
}

struct FwdArgs {
  TcLayout lay;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  float *u, *du, *lap;
  float* ckpt;      // optional: [tile][layer][8 arrays][64][32] activation-jet checkpoints for the reverse kernel
  long long* dbg;   // optional timestamps (development)
};

template <int ND, int LAP>
__global__ void __launch_bounds__(THREADS, 1) tc_fwd_kernel(FwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a[SLOTS], bar_d[SLOTS];
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));   // 1 KB aligned, stays a shared pointer
  const TcLayout lay = a.lay;
  const FwdSmem so = fwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weights to smem, barriers, TMEM
  for (int l = 1; l < L; ++l)
    for (int part = 0; part < 2; ++part) {
      const float4* src = reinterpret_cast<const float4*>(a.img + lay.mat(l, part));
      float4* dst = reinterpret_cast<float4*>(sm + so.mats + ((l - 1) * 2 + part) * 4096);
      for (int v = tid; v < 1024; v += THREADS) dst[v] = __ldg(src + v);
    }
  for (int v = tid; v < (int)lay.small(); v += THREADS) sm[so.small + v] = __ldg(a.img + v);
  if (warp == EPI_WARPS) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&bar_a[s], SLOT_WARPS);
      mbar_init(&bar_d[s], 1);
    }
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;

  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  const int64_t npairs = (ntiles + SLOTS - 1) / SLOTS;

  if (warp == EPI_WARPS) {
    // =========================== MMA issuer ===========================
    const uint32_t wbase = smem_u32(sm + so.mats);
    constexpr uint32_t idesc = idesc_tf32(128, HP);
    uint32_t ph[SLOTS] = {0, 0};
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      for (int l = 1; l < L; ++l) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          if (pair * SLOTS + s >= ntiles) continue;
          mbar_wait(&bar_a[s], ph[s]);
          ph[s] ^= 1;
          fence_after();
          const bool dbg_on = a.dbg && blockIdx.x == 0 && pair == blockIdx.x + gridDim.x && s == 0;
          if (dbg_on && lane == 0) a.dbg[256 + l * 4 + 0] = clock64();
          if (elect_one()) {
            const uint32_t tD = tmem + s * SLOT_COLS, tAh = tD + 64, tAl = tD + 128;
            const uint32_t wh = wbase + (uint32_t)((l - 1) * 2) * 16384u, wl = wh + 16384u;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wl + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAl + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
            mma_commit(&bar_d[s]);
          }
          __syncwarp();
          if (dbg_on && lane == 0) a.dbg[256 + l * 4 + 1] = clock64();
        }
      }
    }
  } else {
    // =========================== activation warps ===========================
    // warp = slot * 8 + half * 4 + q : quadrant q (= stream q) of the slot's tile, neurons [NB half, NB half + NB)
    const int slot = warp / SLOT_WARPS, half = (warp % SLOT_WARPS) >> 2, q = warp & 3;
    const int j0 = half * NB;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tmem + slot * SLOT_COLS + lane_base, tAh = tD + 64, tAl = tD + 128;
    float* ex_z = sm + so.ex + slot * 3 * HP * TILE;       // [j][p]   tanh(z) of the value stream
    float* ex_dz = ex_z + HP * TILE;                       // [k][j][p] pre-activation of the derivative streams
    const float* W0 = sm + so.small + lay.w0();
    const float* WL = sm + so.small + lay.wl();
    const float* bL = sm + so.small + lay.bl();
    const int barid = 1 + slot * HALVES + half;
    uint32_t ph_d = 0;
    const bool is_val = (q == 0), is_dk = (q >= 1 && q <= ND), is_lap = (LAP && q == ND + 1), idle = (q >= S);
    const int dcol = is_dk ? a.jet.dcol[q - 1] : 0;
    float lapw[ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < ND; ++k) lapw[k] = a.jet.lap_w[k];

    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t tile = pair * SLOTS + slot;
      if (tile >= ntiles) continue;
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
      float outacc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) outacc[o] = 0.0f;

      for (int l = 0; l < L; ++l) {
        const float* bias = sm + so.small + lay.bias(l);
        const bool last = (l == L - 1);
        float* ckl = a.ckpt ? a.ckpt + ((size_t)tile * L + l) * (8 * HP * TILE) : nullptr;
        const bool dbg_on = a.dbg && blockIdx.x == 0 && pair == blockIdx.x + gridDim.x && lane == 0 && half == 0 && slot == 0;
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 0] = clock64();
        if (l > 0) {
          mbar_wait(&bar_d[slot], ph_d);
          ph_d ^= 1;
          fence_after();
        }
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 1] = clock64();
        float xr[TPX_MAX_DIN];
        if (l == 0 && is_val) {
#pragma unroll
          for (int cc = 0; cc < TPX_MAX_DIN; ++cc) xr[cc] = (valid && cc < d_in) ? __ldg(a.x + gp * d_in + cc) : 0.0f;
        }
        // own pre-activation values of neurons [jc, jc + CH)
        auto own = [&](int jc, float (&w)[CH]) {
          if (l == 0) {
            if (is_val) {
#pragma unroll
              for (int i = 0; i < CH; ++i) w[i] = bias[jc + i];
#pragma unroll
              for (int cc = 0; cc < TPX_MAX_DIN; ++cc) {
                if (cc < d_in) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) w[i] = fmaf(W0[cc * HP + jc + i], xr[cc], w[i]);
                }
              }
            } else if (is_dk) {
#pragma unroll
              for (int i = 0; i < CH; ++i) w[i] = W0[dcol * HP + jc + i];
            } else {
#pragma unroll
              for (int i = 0; i < CH; ++i) w[i] = 0.0f;
            }
          } else {
            tmem_ld16(tD + jc, w);
            if (is_val) {
#pragma unroll
              for (int i = 0; i < CH; ++i) w[i] += bias[jc + i];
            }
          }
        };
        // ---- phase 1: publish what the other streams of the point need: tanh(z) and d_k z
        // (one straight-line loop per role so that the 16 independent chains of a chunk interleave)
        if (is_val) {
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            float w[CH];
            own(j0 + c * CH, w);
#pragma unroll
            for (int i = 0; i < CH; ++i) w[i] = tanh_fast(w[i]);
#pragma unroll
            for (int i = 0; i < CH; ++i) ex_z[(j0 + c * CH + i) * TILE + lane] = w[i];
            if (ckl) {
#pragma unroll
              for (int i = 0; i < CH; ++i) ckl[(j0 + c * CH + i) * TILE + lane] = w[i];
            }
          }
        } else if (is_dk && LAP) {
          float* dst = ex_dz + (q - 1) * HP * TILE;
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            float w[CH];
            own(j0 + c * CH, w);
#pragma unroll
            for (int i = 0; i < CH; ++i) dst[(j0 + c * CH + i) * TILE + lane] = w[i];
          }
        }
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 2] = clock64();
        named_bar_sync(barid, 128);
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 3] = clock64();
        // ---- phase 2: this stream's activation output
        auto emit = [&](int jc, float (&w)[CH]) {
          if (!last) {
            float lo[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float hi = tf32_hi(w[i]);
              lo[i] = w[i] - hi;
              w[i] = hi;
            }
            tmem_st16(tAh + jc, w);
            tmem_st16(tAl + jc, lo);
          } else {
#pragma unroll
            for (int o = 0; o < 8; ++o) {
              if (o < d_out) {
                float acc = outacc[o];
#pragma unroll
                for (int i = 0; i < CH; ++i) acc = fmaf(WL[o * HP + jc + i], w[i], acc);
                outacc[o] = acc;
              }
            }
          }
        };
#pragma unroll
        for (int o = 0; o < 8; ++o) outacc[o] = 0.0f;
        if (is_val) {
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            const int jc = j0 + c * CH;
            float w[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) w[i] = ex_z[(jc + i) * TILE + lane];
            emit(jc, w);
          }
        } else if (is_dk) {
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            const int jc = j0 + c * CH;
            float w[CH];
            own(jc, w);
            if (ckl) {
#pragma unroll
              for (int i = 0; i < CH; ++i) ckl[((q * 2 + 0) * HP + jc + i) * TILE + lane] = w[i];
            }
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float t = ex_z[(jc + i) * TILE + lane];
              w[i] *= fmaf(-t, t, 1.0f);
            }
            if (ckl) {
#pragma unroll
              for (int i = 0; i < CH; ++i) ckl[((q * 2 + 1) * HP + jc + i) * TILE + lane] = w[i];
            }
            emit(jc, w);
          }
        } else if (is_lap) {
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            const int jc = j0 + c * CH;
            float w[CH];
            own(jc, w);
            if (ckl) {
#pragma unroll
              for (int i = 0; i < CH; ++i) ckl[((q * 2 + 0) * HP + jc + i) * TILE + lane] = w[i];
            }
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float t = ex_z[(jc + i) * TILE + lane];
              const float s1 = fmaf(-t, t, 1.0f);
              float qq = 0.0f;
#pragma unroll
              for (int k = 0; k < ND; ++k) {
                const float dz = ex_dz[(k * HP + jc + i) * TILE + lane];
                qq = fmaf(lapw[k] * dz, dz, qq);
              }
              w[i] = fmaf(-2.0f * t * s1, qq, s1 * w[i]);
            }
            if (ckl) {
#pragma unroll
              for (int i = 0; i < CH; ++i) ckl[((q * 2 + 1) * HP + jc + i) * TILE + lane] = w[i];
            }
            emit(jc, w);
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < NB / CH; ++c) {
            float w[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) w[i] = 0.0f;
            emit(j0 + c * CH, w);
          }
        }
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 4] = clock64();
        if (!last) {
          tmem_wait_st();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a[slot]);
        }
        if (dbg_on) a.dbg[(q * 8 + l) * 8 + 5] = clock64();
      }
      // ---- outputs of this stream: the two halves of a row meet through shared memory
      float* part = sm + so.part + (slot * 4 + q) * (8 * TILE);          // [o][p] partial sums of half 1
      if (half == 1) {
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (o < d_out) part[o * TILE + lane] = outacc[o];
      }
      named_bar_sync(1 + SLOTS * HALVES + slot, 32 * SLOT_WARPS);
      if (half == 0 && valid && !idle) {
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (o < d_out) {
            const float r = outacc[o] + part[o * TILE + lane];
            if (is_val) a.u[gp * d_out + o] = r + bL[o];
            else if (is_dk) a.du[(gp * d_out + o) * ND + (q - 1)] = r;
            else a.lap[gp * d_out + o] = r;
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem, TMEM_COLS);
}

static int env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

}  // namespace tc

using namespace tc;

long long* g_tc_debug = nullptr;

bool tc_net_supported(const tpx_fcn_desc* net) {
  static const int disabled = env_flag("TPX_DISABLE_TC");
  if (disabled) return false;
  if (net->activation != TPX_ACT_TANH) return false;
  if (net->n_hidden < 2) return false;                 // needs at least one hidden x hidden product
  for (int l = 0; l < net->n_hidden; ++l)
    if (net->widths[l] > HP) return false;
  if (net->d_out > 8) return false;
  return true;
}

bool tc_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tc_net_supported(net)) return false;
  if (1 + nd + lap > 4) return false;
  if (lap && nd == 0) return false;
  return true;
}

size_t tc_packed_floats(const tpx_fcn_desc* net) {
  TcLayout lay{net->n_hidden, net->d_in, net->d_out};
  return lay.total();
}

int tc_pack(const tpx_fcn_desc* net, const void* const* params_host, float* img, cudaStream_t stream) {
  TcLayout lay{net->n_hidden, net->d_in, net->d_out};
  TcPackArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int l = 0; l <= net->n_hidden; ++l) {
    pa.in_w[l] = (l == 0) ? net->d_in : net->widths[l - 1];
    pa.out_w[l] = (l == net->n_hidden) ? net->d_out : net->widths[l];
    pa.W[l] = (const float*)params_host[2 * l];
    pa.b[l] = (const float*)params_host[2 * l + 1];
  }
  if (cudaMemsetAsync(img, 0, lay.total() * sizeof(float), stream) != cudaSuccess) return TPX_E_CUDA;
  tc_pack_fill_kernel<<<64, 256, 0, stream>>>(lay, pa, img);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

template <int ND, int LAP>
static int launch_tc_fwd(const FwdArgs& a, int sm_count, cudaStream_t stream) {
  const FwdSmem so = fwd_smem(a.lay);
  const size_t smem = (size_t)so.total * sizeof(float) + 1024;
  if (smem > 227 * 1024) return TPX_E_UNSUPPORTED;
  if (cudaFuncSetAttribute(tc_fwd_kernel<ND, LAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPX_E_CUDA;
  const int64_t ntiles = (a.n + TILE - 1) / TILE, npairs = (ntiles + SLOTS - 1) / SLOTS;
  int64_t grid = sm_count;
  if (grid > npairs) grid = npairs;
  if (grid < 1) return 0;
  tc_fwd_kernel<ND, LAP><<<(unsigned)grid, THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

int tc_forward(const tpx_fcn_desc* net, const JetArgs& jet, int lap, const float* img, const float* x, int64_t n,
               float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream) {
  FwdArgs a;
  a.lay = TcLayout{net->n_hidden, net->d_in, net->d_out};
  a.jet = jet;
  a.img = img; a.x = x; a.n = n; a.u = u; a.du = du; a.lap = lapo;
  a.ckpt = saved;
  a.dbg = g_tc_debug;
  switch (jet.nd * 2 + lap) {
    case 0: return launch_tc_fwd<0, 0>(a, sm_count, stream);
    case 2: return launch_tc_fwd<1, 0>(a, sm_count, stream);
    case 3: return launch_tc_fwd<1, 1>(a, sm_count, stream);
    case 4: return launch_tc_fwd<2, 0>(a, sm_count, stream);
    case 5: return launch_tc_fwd<2, 1>(a, sm_count, stream);
    case 6: return launch_tc_fwd<3, 0>(a, sm_count, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

}  // namespace tpx
