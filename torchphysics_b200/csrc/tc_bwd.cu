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

constexpr int B_CH = 16;                 // neurons per activation warp
constexpr int B_GROUPS = HP / B_CH;      // warp groups: the 4 quadrant warps that share one chunk of neurons
constexpr int B_EPI_WARPS = 4 * B_GROUPS;
constexpr int B_THREADS = 32 * (B_EPI_WARPS + 1);
constexpr uint32_t B_TMEM_COLS = 512;
constexpr uint32_t T_D = 0, T_AH = 64, T_AL = 128, T_DW = 192;   // TMEM columns
constexpr int MAXL = 6;                  // hidden layers supported (dW accumulators: 64 columns each)
constexpr int MAXIO = 4;                 // d_in, d_out supported by the register accumulators
constexpr int EX_ARRAYS = 7;             // exchange arrays of 16 x 32 floats per warp group

struct BwdSmem {
  int wring, zt, at, ex, small, total;   // float offsets from the 1 KB aligned base
};
__host__ __device__ inline BwdSmem bwd_smem(const TcLayout& lay) {
  BwdSmem s;
  s.wring = 0;                            // (hi, lo) x 4096: one product image, refetched per phase
  s.zt = 8192;                            // 128 rows x 128 K
  s.at = s.zt + 16384;                    // hi: 64 rows x 128 K, lo: 64 rows x 128 K
  s.ex = s.at + 16384;
  s.small = s.ex + B_GROUPS * EX_ARRAYS * B_CH * TILE;
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
  float* ckpt;       // RECOMP: per CTA, else per tile: L x 8 arrays x 64 x 32 floats
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

template <int ND, int LAP, bool RECOMP>
__global__ void __launch_bounds__(B_THREADS, 1) tc_bwd_kernel(BwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a, bar_d, bar_f, bar_w;
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const TcLayout lay = a.lay;
  const Layout lay32 = a.lay32;
  const BwdSmem so = bwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nphase = RECOMP ? 2 * (L - 1) : (L - 1);   // MMA phases per tile: [F1..F(L-1),] R(L-1)..R1

  for (int v = tid; v < (int)lay.small(); v += B_THREADS) sm[so.small + v] = __ldg(a.img + v);
  if (warp == B_EPI_WARPS) tmem_alloc(&tmem_slot, B_TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar_a, B_EPI_WARPS);
    mbar_init(&bar_d, 1);
    mbar_init(&bar_f, 1);
    mbar_init(&bar_w, 1);
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
    uint32_t ph_a = 0, ph_w = 0, ph_d = 0;
    auto phase_layer = [&](int64_t g, bool& fwd) -> int {
      const int n = (int)(g % nphase);
      fwd = RECOMP && n < L - 1;
      return fwd ? n + 1 : (L - 1) - (RECOMP ? n - (L - 1) : n);
    };
    auto phase_src = [&](int64_t g) -> const float* {
      bool fwd;
      const int l = phase_layer(g, fwd);
      return a.img + lay.mat(l, fwd ? 0 : 2);
    };
    if (total_phases > 0 && elect_one()) bulk_load(sm + so.wring, phase_src(0), 32768u, &bar_w);
    __syncwarp();
    for (int64_t g = 0; g < total_phases; ++g) {
      bool fwd;
      const int l = phase_layer(g, fwd);
      const bool first_tile = g < nphase;
      mbar_wait(&bar_a, ph_a);
      ph_a ^= 1;
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after();
      if (elect_one()) {
        const uint32_t wh = ring, wl = wh + 16384u;
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
        mma_commit(&bar_d);                  // the adjoint (or forward) product is all the next epilogue needs ...
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
          mma_commit(&bar_f);                // ... the gradient product only gates the reuse of its operands
        }
      }
      __syncwarp();
      // the weight image is free once this phase's TMEM-operand products are done: fetch the next image
      // behind the activation warps' work
      mbar_wait(&bar_d, ph_d);
      ph_d ^= 1;
      if (g + 1 < total_phases) {
        if (elect_one()) bulk_load(sm + so.wring, phase_src(g + 1), 32768u, &bar_w);
        __syncwarp();
      }
    }
  } else {
    // =========================== activation warps ===========================
    // warp = grp * 4 + q: TMEM quadrant q (= stream q of the tile), neurons [16 grp, 16 grp + 16)
    const int grp = warp >> 2, q = warp & 3;
    const int jc = grp * B_CH;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tmem + T_D + lane_base, tAh = tmem + T_AH + lane_base, tAl = tmem + T_AL + lane_base;
    float* ex = sm + so.ex + grp * (EX_ARRAYS * B_CH * TILE);
    float* X_T = ex;                                  // [16][32] tanh(z)
    float* X_DZ = ex + 1 * B_CH * TILE;               // [k][16][32]  d_k z           (k < 3)
    float* X_G = ex + 4 * B_CH * TILE;                // [k][16][32]  adjoint of d_k a (k < 2 with LAP, < 3 without)
    float* X_LZ = ex + (LAP ? 3 : 6) * B_CH * TILE;   // with LAP: ND <= 2, so array 3 is free for L z
    float* X_GL = ex + 6 * B_CH * TILE;               // adjoint of L a
    const float* W0 = sm + so.small + lay.w0();
    const float* WL = sm + so.small + lay.wl();
    const int barid = 1 + grp;
    const bool is_val = (q == 0), is_dk = (q >= 1 && q <= ND), is_lap = (LAP && q == ND + 1), idle = (q >= S);
    const int kq = is_dk ? q - 1 : 0;
    const int dcol = is_dk ? a.jet.dcol[kq] : 0;
    float lapw[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
    uint32_t ph_d = 0, ph_f = 0;
    bool dw_pending = false;             // a gradient product may still be reading Zt / At
    // transposed operands: element (row = neuron, K = this thread's row index 32 q + lane); 128-byte swizzle
    //   offset = (K / 32) * atom + row * 32 + ((((K % 32) / 4) ^ (row % 8)) * 4) + K % 4,   row % 8 == i % 8 (jc % 16 == 0)
    float* Zt = sm + so.zt + q * 4096 + jc * 32 + (lane & 3);
    float* At = sm + so.at + q * 2048 + jc * 32 + (lane & 3);
    int xo[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) xo[m] = ((lane >> 2) ^ m) << 2;

    // register accumulators across tiles; lane i (and i + 16) <-> neuron jc + (i & 15)
    float acc_b[MAXL], acc_w0[MAXIO], acc_wl[MAXIO], acc_bl[MAXIO];
#pragma unroll
    for (int l = 0; l < MAXL; ++l) acc_b[l] = 0.0f;
#pragma unroll
    for (int c = 0; c < MAXIO; ++c) acc_w0[c] = acc_wl[c] = acc_bl[c] = 0.0f;
    auto add_b = [&](int l, float v) {
#pragma unroll
      for (int ll = 0; ll < MAXL; ++ll)
        if (ll == l) acc_b[ll] += v;
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
    auto put_A = [&](const float (&w)[B_CH]) {
      float hi[B_CH], lo[B_CH];
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        hi[i] = tf32_hi(w[i]);
        lo[i] = w[i] - hi[i];
      }
      tmem_st16(tAh + jc, hi);
      tmem_st16(tAl + jc, lo);
    };
    auto put_Zt = [&](const float (&w)[B_CH]) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const float hi = tf32_hi(w[i]);
        Zt[i * 32 + xo[i & 7]] = hi;
        Zt[(64 + i) * 32 + xo[i & 7]] = w[i] - hi;
      }
    };
    auto put_At = [&](const float (&w)[B_CH]) {
#pragma unroll
      for (int i = 0; i < B_CH; ++i) {
        const float hi = tf32_hi(w[i]);
        At[i * 32 + xo[i & 7]] = hi;
        At[8192 + i * 32 + xo[i & 7]] = w[i] - hi;
      }
    };
    // output layer gradient: dWL[o][j] += sum_rows g_row[o] a_{L-1,row}[j]
    auto grad_out = [&](const float (&al)[B_CH], const float (&g)[MAXIO]) {
#pragma unroll
      for (int o = 0; o < MAXIO; ++o)
        if (o < d_out) {
          float v[B_CH];
#pragma unroll
          for (int i = 0; i < B_CH; ++i) v[i] = g[o] * al[i];
          acc_wl[o] += reduce16(v, lane);
        }
    };

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
      float xr[MAXIO];
#pragma unroll
      for (int c = 0; c < MAXIO; ++c) xr[c] = (valid && c < d_in) ? __ldg(a.x + gp * d_in + c) : 0.0f;
      float g[MAXIO];                      // adjoint of this row's output stream
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
      if (is_val && grp == 0) {
#pragma unroll
        for (int o = 0; o < MAXIO; ++o) {
          float s = g[o];
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
          acc_bl[o] += s;
        }
      }
      // checkpoints: recomputed into the per-CTA scratch, or the ones the forward kernel saved for this tile
      float* ck = a.ckpt + (size_t)(RECOMP ? (int64_t)blockIdx.x : tile) * ckpt_floats_per_cta(L);

      // ------------------------------------------------------------------ forward with checkpoints
      if constexpr (RECOMP) {
        for (int l = 0; l < L; ++l) {
          const float* bias = sm + so.small + lay.bias(l);
          const bool last = (l == L - 1);
          if (l > 0) wait_d();
          float* ckl = ck + (size_t)l * 8 * HP * TILE + jc * TILE + lane;
          float w[B_CH];
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
              X_T[i * TILE + lane] = w[i];
              ckl[(0 * HP + i) * TILE] = w[i];
            }
          } else if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              if (LAP) X_DZ[(kq * B_CH + i) * TILE + lane] = w[i];
              ckl[((q * 2 + 0) * HP + i) * TILE] = w[i];
            }
          } else if (is_lap) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) ckl[((q * 2 + 0) * HP + i) * TILE] = w[i];
          }
          named_bar_sync(barid, 128);
          // stage B: this stream's activation output
          if (is_dk) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) {
              const float t = X_T[i * TILE + lane];
              w[i] *= fmaf(-t, t, 1.0f);
              ckl[((q * 2 + 1) * HP + i) * TILE] = w[i];
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
              ckl[((q * 2 + 1) * HP + i) * TILE] = w[i];
            }
          } else if (idle) {
#pragma unroll
            for (int i = 0; i < B_CH; ++i) w[i] = 0.0f;
          }
          if (!last) {
            put_A(w);
            signal_a();
          } else if (!idle) {
            grad_out(w, g);
          }
          named_bar_sync(barid, 128);        // the exchange arrays are reused by the next layer
        }
      }

      // ------------------------------------------------------------------ reverse sweep
      for (int l = L - 1; l >= 0; --l) {
        const float* ckl = ck + (size_t)l * 8 * HP * TILE + jc * TILE + lane;
        const float* ckp = ck + (size_t)(l > 0 ? l - 1 : 0) * 8 * HP * TILE + jc * TILE + lane;
        // this row's checkpoints (global): issued before waiting for the adjoint product
        float cz[B_CH], ap[B_CH];
#pragma unroll
        for (int i = 0; i < B_CH; ++i) {
          cz[i] = idle ? 0.0f : ckl[((q * 2 + 0) * HP + i) * TILE];
          ap[i] = (idle || l == 0) ? 0.0f : ckp[((q * 2 + (is_val ? 0 : 1)) * HP + i) * TILE];
        }
        float ga[B_CH];                    // adjoint of this row's stream of a_l
        if (l == L - 1) {
          if constexpr (!RECOMP) {
            if (!idle) {                   // output layer gradient from the saved jet of the last hidden layer
              float al[B_CH];
#pragma unroll
              for (int i = 0; i < B_CH; ++i) al[i] = ckl[((q * 2 + (is_val ? 0 : 1)) * HP + i) * TILE];
              grad_out(al, g);
            }
          }
#pragma unroll
          for (int i = 0; i < B_CH; ++i) {
            float s = 0.0f;
#pragma unroll
            for (int o = 0; o < MAXIO; ++o)
              if (o < d_out) s = fmaf(WL[o * HP + jc + i], g[o], s);
            ga[i] = s;
          }
        } else {
          wait_d();
          tmem_ld16(tD + jc, ga);
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
          put_A(zb);
          if (dw_pending) {                // the previous gradient product must have finished reading Zt / At
            mbar_wait(&bar_f, ph_f);
            ph_f ^= 1;
          }
          put_Zt(zb);
          put_At(ap);
          signal_a();
          dw_pending = true;
          if (is_val) add_b(l, reduce16(zb, lane));
        } else {
          // layer 0: z = W0 x + b0, d_k z = W0[:, dcol_k]
          if (is_val) {
#pragma unroll
            for (int cc = 0; cc < MAXIO; ++cc)
              if (cc < d_in) {
                float v[B_CH];
#pragma unroll
                for (int i = 0; i < B_CH; ++i) v[i] = zb[i] * xr[cc];
                acc_w0[cc] += reduce16(v, lane);
              }
            acc_b[0] += reduce16(zb, lane);
          } else if (is_dk) {
            acc_w0[0] += reduce16(zb, lane);
          }
        }
        named_bar_sync(barid, 128);          // the exchange arrays are reused by the next step
      }
    }
    if (dw_pending) {                        // the last gradient product of this CTA
      mbar_wait(&bar_f, ph_f);
      fence_after();
    }

    // ------------------------------------------------------------------ flush the accumulators
    const int HP32 = lay32.HP;
    if (lane < 16) {
      const int j = jc + lane;
      if (j < HP32) {
        if (is_val) {
#pragma unroll
          for (int l = 0; l < MAXL; ++l)
            if (l < L) atomicAdd(a.gacc + (l == 0 ? lay32.b0() : lay32.b(l)) + j, (double)acc_b[l]);
#pragma unroll
          for (int cc = 0; cc < MAXIO; ++cc)
            if (cc < d_in) atomicAdd(a.gacc + lay32.w0() + (size_t)cc * HP32 + j, (double)acc_w0[cc]);
        } else if (is_dk) {
          atomicAdd(a.gacc + lay32.w0() + (size_t)dcol * HP32 + j, (double)acc_w0[0]);
        }
        if (!idle) {
#pragma unroll
          for (int o = 0; o < MAXIO; ++o)
            if (o < d_out) atomicAdd(a.gacc + lay32.wl() + (size_t)o * HP32 + j, (double)acc_wl[o]);
        }
      }
    }
    if (is_val && grp == 0 && lane == 0) {
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
        tmem_ld16(tW + jc, w);
#pragma unroll
        for (int i = 0; i < B_CH; ++i) {
          const int col = jc + i;
          if (j < HP32 && col < HP32) atomicAdd(a.gacc + lay32.w(l) + (size_t)j * HP32 + col, (double)w[i]);
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

size_t tc_bwd_workspace_bytes(const tpx_fcn_desc* net, int sm_count) {
  return (size_t)sm_count * ckpt_floats_per_cta(net->n_hidden) * sizeof(float);
}

size_t tc_saved_bytes(const tpx_fcn_desc* net, int64_t n) {
  return (size_t)((n + TILE - 1) / TILE) * ckpt_floats_per_cta(net->n_hidden) * sizeof(float);
}

template <int ND, int LAP, bool RECOMP>
static int launch_tc_bwd(const BwdArgs& a, int sm_count, size_t ws_bytes, cudaStream_t stream) {
  const size_t smem = (size_t)bwd_smem(a.lay).total * sizeof(float) + 1024;
  if (smem > 227 * 1024) return fail(TPX_E_UNSUPPORTED, "tc_bwd_kernel needs %zu bytes of shared memory", smem);
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  const int64_t grid = sm_count < ntiles ? sm_count : ntiles;
  if (grid < 1) return 0;
  const int64_t units = RECOMP ? grid : ntiles;
  if (ws_bytes < (size_t)units * ckpt_floats_per_cta(a.lay.L) * sizeof(float))
    return fail(TPX_E_WORKSPACE, "checkpoint buffer of %zu bytes is too small (%zu needed)", ws_bytes, (size_t)units * ckpt_floats_per_cta(a.lay.L) * sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(tc_bwd_kernel<ND, LAP, RECOMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tc_bwd_kernel: shared memory attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
  tc_bwd_kernel<ND, LAP, RECOMP><<<(unsigned)grid, B_THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tc_bwd_kernel launch (%lld CTAs x %d threads, %zu bytes smem): %s", (long long)grid, B_THREADS, smem, cudaGetErrorString(e));
  return 0;
}

int tc_backward(const tpx_fcn_desc* net, const Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* ckpt, size_t ckpt_bytes,
                int saved, int sm_count, cudaStream_t stream) {
  BwdArgs a;
  a.lay = TcLayout{net->n_hidden, net->d_in, net->d_out};
  a.lay32 = lay32;
  a.jet = jet;
  a.img = img; a.x = x; a.n = n; a.gu = gu; a.gdu = gdu; a.glap = glap; a.gacc = gacc; a.ckpt = ckpt;
#define TPX_TC_BWD_CASE(NDV, LAPV) \
  return saved ? launch_tc_bwd<NDV, LAPV, false>(a, sm_count, ckpt_bytes, stream) : launch_tc_bwd<NDV, LAPV, true>(a, sm_count, ckpt_bytes, stream)
  switch (jet.nd * 2 + lap) {
    case 0: TPX_TC_BWD_CASE(0, 0);
    case 2: TPX_TC_BWD_CASE(1, 0);
    case 3: TPX_TC_BWD_CASE(1, 1);
    case 4: TPX_TC_BWD_CASE(2, 0);
    case 5: TPX_TC_BWD_CASE(2, 1);
    case 6: TPX_TC_BWD_CASE(3, 0);
    default: return TPX_E_UNSUPPORTED;
  }
#undef TPX_TC_BWD_CASE
}

}  // namespace tpx
