// Tensor-core (tcgen05 / TMEM) reverse kernel for hidden width <= 64, tanh, <= 4 streams:
// per tile of 32 points it re-runs the forward jet (checkpointing the activation jets in a
// per-CTA, L2-resident scratch), then sweeps the layers backwards.  Rows of every UMMA are
// (stream q, point p) = TMEM lane 32 q + p, as in the forward kernel (tc_kernels.cu).
//
// Per hidden matrix l (1 <= l < L) the reverse step issues two products:
//   adjoint   D[128 x 64] = Zbar_l[128 x 64] W_l            A = Zbar (hi/lo) in TMEM, B = W_l^T image (K-major)
//   gradient  dW_l[j, i] += sum_r Zbar_l[r, j] a_{l-1}[r, i] r = all 128 rows (4 streams x 32 points) is the K
//             dimension: both operands are written TRANSPOSED into shared memory (row = neuron, K = row
//             index, 128-byte swizzle; one conflict-free scalar store per element), Zbar stacked as
//             [hi ; lo] on M = 128 and a_{l-1} as two N = 64 operands hi / lo, so two chains give
//             hi*hi + lo*hi + hi*lo + lo*lo.  dW_l accumulates in TMEM over ALL tiles of the CTA and is
//             read out once at the end of the kernel.
// Everything with a tiny contraction (layer 0, output layer, biases) is reduced over the 32 lanes with a
// transposing shuffle reduction and accumulated in registers across tiles.
// Weight images (32 KB per product) stream through a 2-deep ring with cp.async.bulk + mbarrier.
#include <cstdio>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_common.cuh"
#include "tc_kernels.h"
#include "tc_layout.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tc {

constexpr int B_NB = 32;                 // neurons per activation warp (two warps per quadrant)
constexpr int B_CH = 16;                 // neurons per exchange chunk
constexpr int B_EPI_WARPS = 8;
constexpr int B_THREADS = 32 * (B_EPI_WARPS + 1);
constexpr uint32_t B_TMEM_COLS = 512;
constexpr uint32_t T_D = 0, T_AH = 64, T_AL = 128, T_DW = 192;   // TMEM columns
constexpr int MAXL = 6;                  // hidden layers supported (dW accumulators: 64 columns each)
constexpr int MAXIO = 4;                 // d_in, d_out supported by the register accumulators
constexpr int EX_ARRAYS = 7;             // exchange arrays of 16 x 32 floats per half group

struct BwdSmem {
  int wring, zt, at, ex, small, total;   // float offsets from the 1 KB aligned base
};
__host__ __device__ inline BwdSmem bwd_smem(const TcLayout& lay) {
  BwdSmem s;
  s.wring = 0;                            // 2 x (hi, lo) x 4096
  s.zt = 2 * 8192;                        // 128 rows x 128 K
  s.at = s.zt + 16384;                    // hi: 64 rows x 128 K, lo: 64 rows x 128 K
  s.ex = s.at + 16384;
  s.small = s.ex + 2 * EX_ARRAYS * B_CH * TILE;
  s.total = s.small + (int)lay.small();
  return s;
}

struct BwdArgs {
  TcLayout lay;
  Layout lay32;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  const float *gu, *gdu, *glap;
  double* gacc;
  float* ckpt;       // per CTA: L x 8 arrays x 64 x 32 floats
};

__host__ __device__ inline size_t ckpt_floats_per_cta(int L) { return (size_t)L * 8 * HP * TILE; }

// lane L receives the sum over all 32 lanes of v[L & 15]  (v is destroyed)
__device__ __forceinline__ float reduce16(float (&v)[B_CH], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const bool up = (lane & s) != 0;
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__device__ __forceinline__ float tanh_fast_b(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -2.885390081777927f));
  const float d = e + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = r * fmaf(-d, r, 2.0f);
  return copysignf((1.0f - e) * r, x);
}

__device__ __forceinline__ void bulk_load(float* dst_smem, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int ND, int LAP>
__global__ void __launch_bounds__(B_THREADS, 1) tc_bwd_kernel(BwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a, bar_d, bar_w[2], bar_m[2];
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const TcLayout lay = a.lay;
  const Layout lay32 = a.lay32;
  const BwdSmem so = bwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nphase = 2 * (L - 1);            // MMA phases per tile: F1..F(L-1), R(L-1)..R1

  for (int v = tid; v < (int)lay.small(); v += B_THREADS) sm[so.small + v] = __ldg(a.img + v);
  if (warp == B_EPI_WARPS) tmem_alloc(&tmem_slot, B_TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar_a, B_EPI_WARPS);
    mbar_init(&bar_d, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_w[b], 1);
      mbar_init(&bar_m[b], 1);
    }
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == B_EPI_WARPS) {
    // =========================== weight producer + MMA issuer ===========================
    const uint32_t ring = smem_u32(sm + so.wring), zt = smem_u32(sm + so.zt), at = smem_u32(sm + so.at);
    constexpr uint32_t idesc = idesc_tf32(128, HP);
    const int64_t total_phases = my_tiles * nphase;
    uint32_t ph_a = 0, ph_w[2] = {0, 0}, ph_m[2] = {0, 0};
    auto phase_src = [&](int64_t g) -> const float* {      // image of global phase g
      const int n = (int)(g % nphase);
      const bool fwd = n < L - 1;
      const int l = fwd ? n + 1 : (L - 1) - (n - (L - 1));
      return a.img + lay.mat(l, fwd ? 0 : 2);
    };
    if (total_phases > 0 && elect_one()) bulk_load(sm + so.wring, phase_src(0), 32768u, &bar_w[0]);
    __syncwarp();
    for (int64_t g = 0; g < total_phases; ++g) {
      const int buf = (int)(g & 1);
      const int n = (int)(g % nphase);
      const bool fwd = n < L - 1;
      const int l = fwd ? n + 1 : (L - 1) - (n - (L - 1));
      const bool first_tile = g < nphase;
      // prefetch the next phase's image into the other buffer once its previous user has finished
      if (g + 1 < total_phases) {
        if (g >= 1) {
          mbar_wait(&bar_m[buf ^ 1], ph_m[buf ^ 1]);
          ph_m[buf ^ 1] ^= 1;
        }
        if (elect_one()) bulk_load(sm + so.wring + (buf ^ 1) * 8192, phase_src(g + 1), 32768u, &bar_w[buf ^ 1]);
        __syncwarp();
      }
      mbar_wait(&bar_a, ph_a);
      ph_a ^= 1;
      mbar_wait(&bar_w[buf], ph_w[buf]);
      ph_w[buf] ^= 1;
      fence_after();
      if (elect_one()) {
        const uint32_t wh = ring + buf * 32768u, wl = wh + 16384u;
        const uint32_t tD = tmem + T_D, tAh = tmem + T_AH, tAl = tmem + T_AL;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wl + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_tf32_ts(tD, tAl + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
        if (!fwd) {
          const uint32_t tW = tmem + T_DW + (uint32_t)(l - 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ss(tW, desc_k_sw128(zt + (ks >> 2) * 16384 + (ks & 3) * 32), desc_k_sw128(at + (ks >> 2) * 8192 + (ks & 3) * 32),
                        idesc, (ks > 0 || !first_tile) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ss(tW, desc_k_sw128(zt + (ks >> 2) * 16384 + (ks & 3) * 32),
                        desc_k_sw128(at + 32768u + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
        }
        mma_commit(&bar_d);
        mma_commit(&bar_m[buf]);
      }
      __syncwarp();
    }
  } else {
    // =========================== activation warps ===========================
    const int half = warp >> 2, q = warp & 3;
    const int j0 = half * B_NB;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tmem + T_D + lane_base, tAh = tmem + T_AH + lane_base, tAl = tmem + T_AL + lane_base;
    float* ex = sm + so.ex + half * (EX_ARRAYS * B_CH * TILE);
    float* X_T = ex;                                  // [16][32] tanh(z)
    float* X_DZ = ex + 1 * B_CH * TILE;               // [k][16][32]  d_k z           (k < 3)
    float* X_G = ex + 4 * B_CH * TILE;                // [k][16][32]  adjoint of d_k a (k < 2 with LAP, < 3 without: shares below)
    float* X_LZ = ex + (LAP ? 3 : 6) * B_CH * TILE;   // with LAP: ND <= 2, so slot 3 is free for L z
    float* X_GL = ex + 6 * B_CH * TILE;               // adjoint of L a
    float* Zt = sm + so.zt;
    float* At = sm + so.at;
    const float* W0 = sm + so.small + lay.w0();
    const float* WL = sm + so.small + lay.wl();
    const int barid = 1 + half;
    const bool is_val = (q == 0), is_dk = (q >= 1 && q <= ND), is_lap = (LAP && q == ND + 1), idle = (q >= S);
    const int kq = is_dk ? q - 1 : 0;
    const int dcol = is_dk ? a.jet.dcol[kq] : 0;
    float lapw[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
    float* ck = a.ckpt + (size_t)blockIdx.x * ckpt_floats_per_cta(L);
    uint32_t ph_d = 0;
    const int kcol = q * 32 + lane;                 // K index of this row in the transposed operands

    // register accumulators across tiles; lane i (and i + 16) <-> neuron j0 + 16 c + (i & 15)
    float acc_b[MAXL][2], acc_w0[MAXIO][2], acc_wl[MAXIO][2], acc_bl[MAXIO];
#pragma unroll
    for (int l = 0; l < MAXL; ++l) acc_b[l][0] = acc_b[l][1] = 0.0f;
#pragma unroll
    for (int c = 0; c < MAXIO; ++c) acc_w0[c][0] = acc_w0[c][1] = acc_wl[c][0] = acc_wl[c][1] = acc_bl[c] = 0.0f;

    auto add_b = [&](int l, int c, float v) {
#pragma unroll
      for (int ll = 0; ll < MAXL; ++ll)
        if (ll == l) {
          if (c == 0) acc_b[ll][0] += v; else acc_b[ll][1] += v;
        }
    };
    auto add_w0 = [&](int cc, int c, float v) {
#pragma unroll
      for (int k = 0; k < MAXIO; ++k)
        if (k == cc) {
          if (c == 0) acc_w0[k][0] += v; else acc_w0[k][1] += v;
        }
    };
    auto add_wl = [&](int o, int c, float v) {
#pragma unroll
      for (int k = 0; k < MAXIO; ++k)
        if (k == o) {
          if (c == 0) acc_wl[k][0] += v; else acc_wl[k][1] += v;
        }
    };
    auto wait_d = [&]() {
      mbar_wait(&bar_d, ph_d);
      ph_d ^= 1;
      fence_after();
    };
    auto signal_a = [&]() {
      tmem_wait_st();
      fence_async_smem();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a);
    };
    // hi/lo split of 16 values of this row -> TMEM A operand columns [jc, jc+16)
    auto put_A = [&](int jc, const float (&w)[B_CH]) {
      float hi[B_CH], lo[B_CH];
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        hi[i] = tf32_hi(w[i]);
        lo[i] = w[i] - hi[i];
      }
      tmem_st16(tAh + jc, hi);
      tmem_st16(tAl + jc, lo);
    };
    // transposed operand stores: element (row, K = kcol)
    auto put_Zt = [&](int jc, const float (&w)[B_CH]) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const float hi = tf32_hi(w[i]);
        Zt[swz128(jc + i, kcol, 128)] = hi;
        Zt[swz128(64 + jc + i, kcol, 128)] = w[i] - hi;
      }
    };
    auto put_At = [&](int jc, const float (&w)[B_CH]) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const float hi = tf32_hi(w[i]);
        At[swz128(jc + i, kcol, 64)] = hi;
        At[8192 + swz128(jc + i, kcol, 64)] = w[i] - hi;
      }
    };

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
      float xr[MAXIO];
#pragma unroll
      for (int c = 0; c < MAXIO; ++c) xr[c] = (valid && c < d_in) ? __ldg(a.x + gp * d_in + c) : 0.0f;
      // adjoint of this row's output stream
      float g[MAXIO];
#pragma unroll
      for (int o = 0; o < MAXIO; ++o) {
        float v = 0.0f;
        if (valid && o < d_out) {
          if (is_val) { if (a.gu) v = __ldg(a.gu + gp * d_out + o); }
          else if (is_dk) { if (a.gdu) v = __ldg(a.gdu + (gp * d_out + o) * ND + kq); }
          else if (is_lap) { if (a.glap) v = __ldg(a.glap + gp * d_out + o); }
        }
        g[o] = v;
      }
      if (is_val) {
#pragma unroll
        for (int o = 0; o < MAXIO; ++o) {
          float s = g[o];
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
          acc_bl[o] += s;
        }
      }

      // ------------------------------------------------------------------ forward with checkpoints
      for (int l = 0; l < L; ++l) {
        const float* bias = sm + so.small + lay.bias(l);
        const bool last = (l == L - 1);
        if (l > 0) wait_d();
        float* ckl = ck + (size_t)l * 8 * HP * TILE;
#pragma unroll 1
        for (int c = 0; c < B_NB / B_CH; ++c) {
          const int jc = j0 + c * B_CH;
          float w[B_CH];
          // own pre-activation values
          if (l == 0) {
            if (is_val) {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) w[i] = bias[jc + i];
#pragma unroll
              for (int cc = 0; cc < MAXIO; ++cc)
                if (cc < d_in) {
#pragma unroll
                  for (int i = 0; i < B_CH; ++i) w[i] = fmaf(W0[cc * HP + jc + i], xr[cc], w[i]);
                }
            } else if (is_dk) {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) w[i] = W0[dcol * HP + jc + i];
            } else {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) w[i] = 0.0f;
            }
          } else {
            tmem_ld16(tD + jc, w);
            if (is_val) {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) w[i] += bias[jc + i];
            }
          }
          // stage A: publish + checkpoint the pre-activation jets
          if (is_val) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) w[i] = tanh_fast_b(w[i]);
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              X_T[(c * 0 + i) * TILE + lane] = w[i];
              ckl[((0 * 2 + 0) * HP + jc + i) * TILE + lane] = w[i];
            }
          } else if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              if (LAP) X_DZ[(kq * B_CH + i) * TILE + lane] = w[i];
              ckl[((q * 2 + 0) * HP + jc + i) * TILE + lane] = w[i];
            }
          } else if (is_lap) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) ckl[((q * 2 + 0) * HP + jc + i) * TILE + lane] = w[i];
          }
          named_bar_sync(barid, 128);
          // stage B: this stream's activation output
          if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = X_T[i * TILE + lane];
              w[i] *= fmaf(-t, t, 1.0f);
              ckl[((q * 2 + 1) * HP + jc + i) * TILE + lane] = w[i];
            }
          } else if (is_lap) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = X_T[i * TILE + lane];
              const float s1 = fmaf(-t, t, 1.0f);
              float qq = 0.0f;
#pragma unroll
              for (int k = 0; k < ND; ++k) {
                const float dz = X_DZ[(k * B_CH + i) * TILE + lane];
                qq = fmaf(lapw[k] * dz, dz, qq);
              }
              w[i] = fmaf(-2.0f * t * s1, qq, s1 * w[i]);
              ckl[((q * 2 + 1) * HP + jc + i) * TILE + lane] = w[i];
            }
          } else if (idle) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) w[i] = 0.0f;
          }
          if (!last) {
            put_A(jc, w);
          } else if (!idle) {
            // output layer gradient: dWL[o][j] += sum_rows g_row[o] a_{L-1,row}[j]
#pragma unroll
            for (int o = 0; o < MAXIO; ++o)
              if (o < d_out) {
                float v[B_CH];
#pragma unroll
                for (int i = 0; i < B_CH; ++i) v[i] = g[o] * w[i];
                add_wl(o, c, reduce16(v, lane));
              }
          }
          named_bar_sync(barid, 128);      // the exchange arrays are reused by the next chunk
        }
        if (!last) signal_a();
      }

      // ------------------------------------------------------------------ reverse sweep
      for (int l = L - 1; l >= 0; --l) {
        const float* ckl = ck + (size_t)l * 8 * HP * TILE;
        const float* ckp = ck + (size_t)(l > 0 ? l - 1 : 0) * 8 * HP * TILE;
        if (l < L - 1) wait_d();
#pragma unroll 1
        for (int c = 0; c < B_NB / B_CH; ++c) {
          const int jc = j0 + c * B_CH;
          float ga[B_CH], cz[B_CH];
          // adjoint of this row's stream of a_l
          if (l == L - 1) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              float s = 0.0f;
#pragma unroll
              for (int o = 0; o < MAXIO; ++o)
                if (o < d_out) s = fmaf(WL[o * HP + jc + i], g[o], s);
              ga[i] = s;
            }
          } else {
            tmem_ld16(tD + jc, ga);
          }
          // checkpointed pre-activation jet of this row: t | d_k z | L z
          if (!idle) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) cz[i] = ckl[((q * 2 + 0) * HP + jc + i) * TILE + lane];
          }
          // stage A: publish
          if (is_val) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) X_T[i * TILE + lane] = cz[i];
          } else if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              X_DZ[(kq * B_CH + i) * TILE + lane] = cz[i];
              X_G[(kq * B_CH + i) * TILE + lane] = ga[i];
            }
          } else if (is_lap) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              X_LZ[i * TILE + lane] = cz[i];
              X_GL[i * TILE + lane] = ga[i];
            }
          }
          named_bar_sync(barid, 128);
          // stage B: adjoint of this row's pre-activation stream
          float zb[B_CH];
          if (is_val) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = cz[i];
              const float s1 = fmaf(-t, t, 1.0f), s2 = -2.0f * t * s1;
              float acc = ga[i] * s1, qq = 0.0f;
#pragma unroll
              for (int k = 0; k < ND; ++k) {
                const float dz = X_DZ[(k * B_CH + i) * TILE + lane];
                acc = fmaf(X_G[(k * B_CH + i) * TILE + lane] * s2, dz, acc);
                if (LAP) qq = fmaf(lapw[k] * dz, dz, qq);
              }
              if (LAP) {
                const float s3 = -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f);
                acc = fmaf(X_GL[i * TILE + lane], fmaf(s3, qq, s2 * X_LZ[i * TILE + lane]), acc);
              }
              zb[i] = acc;
            }
          } else if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = X_T[i * TILE + lane];
              const float s1 = fmaf(-t, t, 1.0f);
              float acc = ga[i] * s1;
              if (LAP) acc = fmaf(2.0f * lapw[kq] * X_GL[i * TILE + lane] * (-2.0f * t * s1), cz[i], acc);
              zb[i] = acc;
            }
          } else if (is_lap) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = X_T[i * TILE + lane];
              zb[i] = ga[i] * fmaf(-t, t, 1.0f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) zb[i] = 0.0f;
          }
          if (l > 0) {
            put_A(jc, zb);
            put_Zt(jc, zb);
            // activation jet of layer l-1 of this row (the other gradient operand)
            float ap[B_CH];
            if (idle) {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) ap[i] = 0.0f;
            } else {
#pragma unroll
              for (int i = 0; i < B_CH; ++i) ap[i] = ckp[((q * 2 + (is_val ? 0 : 1)) * HP + jc + i) * TILE + lane];
            }
            put_At(jc, ap);
            if (is_val) add_b(l, c, reduce16(zb, lane));
          } else {
            // layer 0: z = W0 x + b0, d_k z = W0[:, dcol_k]
            if (is_val) {
#pragma unroll
              for (int cc = 0; cc < MAXIO; ++cc)
                if (cc < d_in) {
                  float v[B_CH];
#pragma unroll
                  for (int i = 0; i < B_CH; ++i) v[i] = zb[i] * xr[cc];
                  add_w0(cc, c, reduce16(v, lane));
                }
              add_b(0, c, reduce16(zb, lane));
            } else if (is_dk) {
              add_w0(0, c, reduce16(zb, lane));
            }
          }
          named_bar_sync(barid, 128);
        }
        if (l > 0) signal_a();
      }
    }

    // ------------------------------------------------------------------ flush the accumulators
    // all MMAs of this CTA have completed: the last reverse step waited for the last commit
    const int HP32 = lay32.HP;
    if (lane < 16) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = j0 + c * B_CH + lane;
        if (j >= HP32) continue;
        if (is_val) {
#pragma unroll
          for (int l = 0; l < MAXL; ++l)
            if (l < L) atomicAdd(a.gacc + (l == 0 ? lay32.b0() : lay32.b(l)) + j, (double)acc_b[l][c]);
#pragma unroll
          for (int cc = 0; cc < MAXIO; ++cc)
            if (cc < d_in) atomicAdd(a.gacc + lay32.w0() + (size_t)cc * HP32 + j, (double)acc_w0[cc][c]);
        } else if (is_dk) {
          atomicAdd(a.gacc + lay32.w0() + (size_t)dcol * HP32 + j, (double)acc_w0[0][c]);
        }
        if (!idle) {
#pragma unroll
          for (int o = 0; o < MAXIO; ++o)
            if (o < d_out) atomicAdd(a.gacc + lay32.wl() + (size_t)o * HP32 + j, (double)acc_wl[o][c]);
        }
      }
    }
    if (is_val && half == 0 && lane == 0) {
#pragma unroll
      for (int o = 0; o < MAXIO; ++o)
        if (o < d_out) atomicAdd(a.gacc + lay32.bl() + o, (double)acc_bl[o]);
    }
    // hidden matrices: rows j (hi, quadrants 0-1) and 64 + j (lo, quadrants 2-3) both add into dW[j][i]
    if (my_tiles > 0) {
      const int j = (q & 1) * 32 + lane;
      for (int l = 1; l < L; ++l) {
        const uint32_t tW = tmem + T_DW + (uint32_t)(l - 1) * 64 + lane_base;
        float w[B_CH];
#pragma unroll 1
        for (int c = 0; c < B_NB / B_CH; ++c) {
          tmem_ld16(tW + j0 + c * B_CH, w);
#pragma unroll
          for (int i = 0; i < B_CH; ++i) {
            const int col = j0 + c * B_CH + i;
            if (j < HP32 && col < HP32) atomicAdd(a.gacc + lay32.w(l) + (size_t)j * HP32 + col, (double)w[i]);
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == B_EPI_WARPS) tmem_dealloc(tmem, B_TMEM_COLS);
}

}  // namespace tc

using namespace tc;

bool tc_bwd_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tc_supported(net, nd, lap)) return false;
  if (net->n_hidden > MAXL || net->d_in > MAXIO || net->d_out > MAXIO) return false;
  return true;
}

template <int ND, int LAP>
static int bwd_grid_smem(const TcLayout& lay, int sm_count, int64_t n, int64_t* grid, size_t* smem) {
  *smem = (size_t)bwd_smem(lay).total * sizeof(float) + 1024;
  if (*smem > 227 * 1024) return TPX_E_UNSUPPORTED;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  *grid = sm_count < ntiles ? sm_count : ntiles;
  return 0;
}

size_t tc_bwd_workspace_bytes(const tpx_fcn_desc* net, int sm_count) {
  return (size_t)sm_count * ckpt_floats_per_cta(net->n_hidden) * sizeof(float);
}

template <int ND, int LAP>
static int launch_tc_bwd(const BwdArgs& a, int sm_count, size_t ws_bytes, cudaStream_t stream) {
  int64_t grid;
  size_t smem;
  int rc = bwd_grid_smem<ND, LAP>(a.lay, sm_count, a.n, &grid, &smem);
  if (rc) return rc;
  if (grid < 1) return 0;
  if (ws_bytes < (size_t)grid * ckpt_floats_per_cta(a.lay.L) * sizeof(float)) return TPX_E_WORKSPACE;
  if (cudaFuncSetAttribute(tc_bwd_kernel<ND, LAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPX_E_CUDA;
  tc_bwd_kernel<ND, LAP><<<(unsigned)grid, B_THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}

int tc_backward(const tpx_fcn_desc* net, const Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* ckpt, size_t ckpt_bytes,
                int sm_count, cudaStream_t stream) {
  BwdArgs a;
  a.lay = TcLayout{net->n_hidden, net->d_in, net->d_out};
  a.lay32 = lay32;
  a.jet = jet;
  a.img = img; a.x = x; a.n = n; a.gu = gu; a.gdu = gdu; a.glap = glap; a.gacc = gacc; a.ckpt = ckpt;
  switch (jet.nd * 2 + lap) {
    case 0: return launch_tc_bwd<0, 0>(a, sm_count, ckpt_bytes, stream);
    case 2: return launch_tc_bwd<1, 0>(a, sm_count, ckpt_bytes, stream);
    case 3: return launch_tc_bwd<1, 1>(a, sm_count, ckpt_bytes, stream);
    case 4: return launch_tc_bwd<2, 0>(a, sm_count, ckpt_bytes, stream);
    case 5: return launch_tc_bwd<2, 1>(a, sm_count, ckpt_bytes, stream);
    case 6: return launch_tc_bwd<3, 0>(a, sm_count, ckpt_bytes, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

}  // namespace tpx
