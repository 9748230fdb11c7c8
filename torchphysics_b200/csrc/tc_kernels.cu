// Tensor-core (tcgen05 / TMEM) forward-jet kernel for hidden width <= 64, tanh or sin, <= 4 streams.
//
// Work decomposition ("R32"): a tile is 32 collocation points.  The 128 rows of one UMMA are
// (stream q, point p) = TMEM lane 32 q + p: every stream (value, d_k, Laplacian) of every point
// is one row of the SAME weight matrix product, so one M=128 N=64 K=64 product per layer serves a
// whole tile.  Per layer:   D[128 x 64] = A[128 x 64] W^T   with fp32-grade accuracy from three
// tf32 products  A_hi W_hi + A_hi W_lo + A_lo W_hi  (hi = round-to-nearest tf32, lo = exact rest).
//   * A_hi / A_lo live in TMEM (written by the activation warps with tcgen05.st), W_hi / W_lo in
//     shared memory (128-byte swizzle, K-major), D in TMEM: the MMA runs at its 32-cycle floor.
//   * warp q of a tile slot owns TMEM quadrant q = the rows of stream q.  The activation jets couple the
//     streams of a point (d_k h = s'(z) d_k z, L h = s''(z) sum c_k (d_k z)^2 + s'(z) L z), so each layer step
//     moves D through a transposing exchange buffer in shared memory: rows of a stream in, all streams of a
//     point out (see the comment above FwdSmem); tanh and the jets are then thread-local on every warp.
//   * two tile slots per CTA ping-pong: while the tensor core works on one tile's layer, the
//     activation warps of the other tile run.  One persistent CTA per SM; all weights resident.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fcn_common.cuh"
#include "f32x2.cuh"
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
// Activation step of one layer for a group of 4 warps (one per TMEM quadrant = stream) and 32 neurons:
//   in      the OWNER warp of stream s reads its rows of D (tcgen05.ld, lane = point) and writes them to the
//           group's exchange buffer  ex[s][p][neuron]  (row padded to 36 floats: 128-bit accesses, no conflicts)
//   compute warp w takes neurons [8 w, 8 w + 8) of the group for ALL streams of its 32 points: tanh and the jet
//           formulas are thread-local (no role divergence, every scheduler does a quarter of the tanh work),
//           checkpoints go to HBM from here, results overwrite ex in place
//   out     the owner reads its stream's 32 neurons back, splits hi / lo and writes the next A operand to TMEM
constexpr int EXP = 36;                       // padded neuron extent of one (stream, point) row
constexpr int EX_GROUP = 4 * TILE * EXP;      // floats of one group's exchange buffer
constexpr int OC = 2;                         // output columns per reduction pass of the output layer

struct FwdSmem {
  // offsets in floats from the 1024-byte aligned base
  int mats;      // (L-1) * 2 * 4096   W_hi, W_lo per hidden layer
  int small;     // copy of the small block
  int ex;        // SLOTS * HALVES exchange buffers
  int part;      // SLOTS * SLOT_WARPS * 4 * OC * TILE partial output sums
  int total;
};
__host__ __device__ inline FwdSmem fwd_smem(const TcLayout& lay) {
  FwdSmem s;
  s.mats = 0;
  s.small = (lay.L - 1) * 2 * 4096;
  s.ex = s.small + (int)lay.small();
  s.part = s.ex + SLOTS * HALVES * EX_GROUP;
  s.total = s.part + SLOTS * SLOT_WARPS * 4 * OC * TILE;
  return s;
}

struct FwdArgs {
  TcLayout lay;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  float *u, *du, *lap;
  float* ckpt;      // optional: [tile][layer][4 arrays][64][32] checkpoints for the reverse kernel: tanh(z), d_k z, L z
  long long* dbg;   // development: clock64 stamps of one tile (scripts/tc_probe.py N dbgf)
};

template <int ND, int LAP, int ACT>       // ACT: 0 = tanh, 1 = sin (tp.models.Sinus; the checkpoint then keeps z instead of tanh z)
__global__ void __launch_bounds__(THREADS, 1) tc_fwd_kernel(FwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
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
#ifdef TPX_TIMELINE
          const bool mdbg = a.dbg && blockIdx.x == 0 && pair == blockIdx.x + 2 * (int64_t)gridDim.x && lane == 0;
          if (mdbg) a.dbg[512 + (l * 2 + s) * 2] = clock64();
#endif
          if (elect_one()) {
            const uint32_t tD = tmem + s * SLOT_COLS, tAh = tD + 64, tAl = tD + 128;
            const uint32_t wh = wbase + (uint32_t)((l - 1) * 2) * 16384u, wl = wh + 16384u;
            // the two small correction products first, the main product last: the TMEM accumulator rounds toward zero on
            // every add (one ulp of the RUNNING sum), so adds onto a still-small sum cost nothing (measured: 3x smaller error)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wl + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAl + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_tf32_ts(tD, tAh + ks * 8, desc_k_sw128(wh + (ks >> 2) * 8192 + (ks & 3) * 32), idesc, 1);
            mma_commit(&bar_d[s]);
          }
          __syncwarp();
#ifdef TPX_TIMELINE
          if (mdbg) a.dbg[512 + (l * 2 + s) * 2 + 1] = clock64();
#endif
        }
      }
    }
  } else {
    // =========================== activation warps ===========================
    // warp = slot * 8 + half * 4 + quad: TMEM quadrant quad (= stream quad) of the slot's tile; the group
    // (slot, half) shares neurons [NB half, NB half + NB)
    const int slot = warp / SLOT_WARPS, half = (warp % SLOT_WARPS) >> 2, quad = warp & 3;
    const int j0 = half * NB;                 // first neuron of the group
    const int jw = j0 + quad * 8;             // compute phase: neurons [jw, jw + 8), all streams
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tD = tmem + slot * SLOT_COLS + lane_base, tAh = tD + 64, tAl = tD + 128;
    float* ex = sm + so.ex + (slot * HALVES + half) * EX_GROUP;
    float* ex_own = ex + (quad * TILE + lane) * EXP;            // owner view: stream quad, the group's 32 neurons
    float* ex_cmp = ex + lane * EXP + quad * 8;                 // compute view of stream s: + s * TILE * EXP
    float* part = sm + so.part + slot * (SLOT_WARPS * 4 * OC * TILE);
    const float* W0 = sm + so.small + lay.w0();
    const float* WL = sm + so.small + lay.wl();
    const float* bL = sm + so.small + lay.bl();
    const int barid = 1 + slot * HALVES + half;
    const int slot_bar = 1 + SLOTS * HALVES + slot;
    const bool owner = quad < S;              // this warp's TMEM rows carry a stream
    uint32_t ph_d = 0;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    if (!owner) {                              // rows of an unused quadrant: a zero A operand, written once
      float zero[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) zero[i] = 0.0f;
#pragma unroll
      for (int c = 0; c < NB / CH; ++c) {
        tmem_st16(tAh + j0 + c * CH, zero);
        tmem_st16(tAl + j0 + c * CH, zero);
      }
      tmem_wait_st();
    }

    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t tile = pair * SLOTS + slot;
      if (tile >= ntiles) continue;
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
      float xr[TPX_MAX_DIN];
#pragma unroll
      for (int cc = 0; cc < TPX_MAX_DIN; ++cc) xr[cc] = (valid && cc < d_in) ? __ldg(a.x + gp * d_in + cc) : 0.0f;

#ifdef TPX_TIMELINE
      long long* dbgp = (a.dbg && blockIdx.x == 0 && pair == blockIdx.x + 2 * (int64_t)gridDim.x && lane == 0 && (warp == 0 || warp == 11))
                            ? a.dbg + (warp == 0 ? 0 : 256) : nullptr;
#endif
#ifdef TPX_TIMELINE
#define TPX_FSTAMP(e) do { if (dbgp) dbgp[l * 8 + (e)] = clock64(); } while (0)
#else
#define TPX_FSTAMP(e) do { } while (0)
#endif
      for (int l = 0; l < L; ++l) {
        TPX_FSTAMP(0);
        const float* bias = sm + so.small + lay.bias(l) + jw;
        const bool last = (l == L - 1);
        float z[S][8];                         // pre-activation streams, then activation streams, of 8 neurons
        if (l == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) z[0][i] = bias[i];
#pragma unroll
          for (int cc = 0; cc < TPX_MAX_DIN; ++cc)
            if (cc < d_in) {
#pragma unroll
              for (int i = 0; i < 8; ++i) z[0][i] = fmaf(W0[cc * HP + jw + i], xr[cc], z[0][i]);
            }
#pragma unroll
          for (int k = 0; k < ND; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) z[1 + k][i] = W0[dcol[k] * HP + jw + i];
          if (LAP) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[S - 1][i] = 0.0f;
          }
        } else {
          // ---- in: D rows of this warp's stream -> exchange buffer
          mbar_wait(&bar_d[slot], ph_d);
          ph_d ^= 1;
          fence_after();
          TPX_FSTAMP(1);
          if (owner) {
#pragma unroll
            for (int c = 0; c < NB / CH; ++c) {
              float w[CH];
              tmem_ld16(tD + j0 + c * CH, w);
#pragma unroll
              for (int v = 0; v < CH / 4; ++v)
                *reinterpret_cast<float4*>(ex_own + c * CH + 4 * v) = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
            }
          }
          named_bar_sync(barid, 128);
          TPX_FSTAMP(2);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float4 lo4 = *reinterpret_cast<const float4*>(ex_cmp + s * TILE * EXP);
            const float4 hi4 = *reinterpret_cast<const float4*>(ex_cmp + s * TILE * EXP + 4);
            z[s][0] = lo4.x; z[s][1] = lo4.y; z[s][2] = lo4.z; z[s][3] = lo4.w;
            z[s][4] = hi4.x; z[s][5] = hi4.y; z[s][6] = hi4.z; z[s][7] = hi4.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) z[0][i] += bias[i];
        }
        // ---- compute: thread-local jets of 8 neurons (all streams of this point)
        float* ckl = a.ckpt ? a.ckpt + ((size_t)tile * L + l) * (4 * HP * TILE) + jw * TILE + lane : nullptr;
        if constexpr (ACT == 0) {
          // tanh: neuron pairs in packed fp32 (f32x2.cuh): the same operations as the scalar path, two per instruction
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float x0 = z[0][i], x1 = z[0][i + 1];
            float e0, e1, r0, r1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fabsf(x0) * -2.885390081777927f));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fabsf(x1) * -2.885390081777927f));
            const F2 e = f2(e0, e1);
            const F2 d = e + f2(1.0f);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.v.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.v.y));
            F2 r = f2(r0, r1);
            r = r * fma2(neg(d), r, f2(2.0f));
            const F2 ta = (f2(1.0f) + neg(e)) * r;
            const F2 t = f2(copysignf(ta.v.x, x0), copysignf(ta.v.y, x1));
            const F2 s1 = fma2(neg(t), t, f2(1.0f));
            const F2 s2 = (f2(-2.0f) * t) * s1;
            z[0][i] = t.v.x; z[0][i + 1] = t.v.y;
            if (ckl) { __stcs(ckl + (0 * HP + i) * TILE, t.v.x); __stcs(ckl + (0 * HP + i + 1) * TILE, t.v.y); }
            F2 qq = f2(0.0f);
#pragma unroll
            for (int k = 0; k < ND; ++k) {
              const F2 dz = f2(z[1 + k][i], z[1 + k][i + 1]);
              if (LAP) qq = fma2(f2(lapw[k]) * dz, dz, qq);
              const F2 a = s1 * dz;
              z[1 + k][i] = a.v.x; z[1 + k][i + 1] = a.v.y;
              if (ckl) { __stcs(ckl + ((1 + k) * HP + i) * TILE, dz.v.x); __stcs(ckl + ((1 + k) * HP + i + 1) * TILE, dz.v.y); }
            }
            if (LAP) {
              const F2 lz = f2(z[S - 1][i], z[S - 1][i + 1]);
              const F2 a = fma2(s2, qq, s1 * lz);
              z[S - 1][i] = a.v.x; z[S - 1][i + 1] = a.v.y;
              if (ckl) { __stcs(ckl + ((S - 1) * HP + i) * TILE, lz.v.x); __stcs(ckl + ((S - 1) * HP + i + 1) * TILE, lz.v.y); }
            }
          }
        } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t, s1, s2, keep;           // activation, its first two derivatives, what the reverse sweep keeps
          keep = z[0][i];
          sincosf(keep, &t, &s1);
          s2 = -t;
          z[0][i] = t;
          if (ckl) __stcs(ckl + (0 * HP + i) * TILE, keep);
          float qq = 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const float dz = z[1 + k][i];
            if (LAP) qq = fmaf(lapw[k] * dz, dz, qq);
            z[1 + k][i] = s1 * dz;
            if (ckl) __stcs(ckl + ((1 + k) * HP + i) * TILE, dz);
          }
          if (LAP) {
            const float lz = z[S - 1][i];
            z[S - 1][i] = fmaf(s2, qq, s1 * lz);
            if (ckl) __stcs(ckl + ((S - 1) * HP + i) * TILE, lz);
          }
        }
        }
        TPX_FSTAMP(3);
        if (!last) {
          // ---- out: activation streams -> exchange buffer -> owners split hi / lo into the next A operand
#pragma unroll
          for (int s = 0; s < S; ++s) {
            *reinterpret_cast<float4*>(ex_cmp + s * TILE * EXP) = make_float4(z[s][0], z[s][1], z[s][2], z[s][3]);
            *reinterpret_cast<float4*>(ex_cmp + s * TILE * EXP + 4) = make_float4(z[s][4], z[s][5], z[s][6], z[s][7]);
          }
          named_bar_sync(barid, 128);
          TPX_FSTAMP(4);
          if (owner) {
#pragma unroll
            for (int c = 0; c < NB / CH; ++c) {
              float hi[CH], lo[CH];
#pragma unroll
              for (int v = 0; v < CH / 4; ++v) {
                const float4 w4 = *reinterpret_cast<const float4*>(ex_own + c * CH + 4 * v);
                hi[4 * v] = w4.x; hi[4 * v + 1] = w4.y; hi[4 * v + 2] = w4.z; hi[4 * v + 3] = w4.w;
              }
#pragma unroll
              for (int i = 0; i < CH; i += 2) {
                const F2 l = fma2(f2(-1.0f), f2(tf32_trunc(hi[i]), tf32_trunc(hi[i + 1])), f2(hi[i], hi[i + 1]));      // x - trunc(x), exact
                lo[i] = l.v.x; lo[i + 1] = l.v.y;        // hi stays the raw word: the MMA truncates it (tc_common.cuh)
              }
              tmem_st16(tAh + j0 + c * CH, hi);
              tmem_st16(tAl + j0 + c * CH, lo);
            }
            tmem_wait_st();
          }
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_a[slot]);
          TPX_FSTAMP(5);
        } else if (a.u != nullptr) {
          // ---- output layer (skipped by the checkpoint-only pass of the reverse sweep): every warp holds partial dot products over its 8 neurons for all streams; the 8
          // warps of the slot meet in shared memory, OC output columns per pass
          const int w8 = warp % SLOT_WARPS;
          for (int o0 = 0; o0 < d_out; o0 += OC) {
#pragma unroll
            for (int oo = 0; oo < OC; ++oo) {
              const int o = (o0 + oo < d_out) ? o0 + oo : o0;
#pragma unroll
              for (int s = 0; s < S; ++s) {
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(WL[o * HP + jw + i], z[s][i], acc);
                part[((w8 * 4 + s) * OC + oo) * TILE + lane] = acc;
              }
            }
            named_bar_sync(slot_bar, 32 * SLOT_WARPS);
            if (half == 0 && owner && valid) {
#pragma unroll
              for (int oo = 0; oo < OC; ++oo) {
                const int o = o0 + oo;
                if (o < d_out) {
                  float r = 0.0f;
#pragma unroll
                  for (int w = 0; w < SLOT_WARPS; ++w) r += part[((w * 4 + quad) * OC + oo) * TILE + lane];
                  if (quad == 0) a.u[gp * d_out + o] = r + bL[o];
                  else if (quad <= ND) a.du[(gp * d_out + o) * ND + (quad - 1)] = r;
                  else a.lap[gp * d_out + o] = r;
                }
              }
            }
            if (o0 + OC < d_out) named_bar_sync(slot_bar, 32 * SLOT_WARPS);   // part is reused by the next pass
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

// shape test only (no debug switch): the packed layout depends on it
bool tc_net_fits(const tpx_fcn_desc* net) {
  if (net->activation != TPX_ACT_TANH && net->activation != TPX_ACT_SIN) return false;
  if (net->n_hidden < 2) return false;                 // needs at least one hidden x hidden product
  for (int l = 0; l < net->n_hidden; ++l)
    if (net->widths[l] > HP) return false;
  if (net->d_out > 8 || net->d_in > 4) return false;   // (the reverse kernel takes 4 inputs; such nets run on the layer-major kernels)
  // all weight images stay resident in shared memory next to the exchange buffers: up to 4 hidden matrices
  const TcLayout lay{net->n_hidden, net->d_in, net->d_out};
  if ((size_t)fwd_smem(lay).total * sizeof(float) + 1024 > 227 * 1024) return false;
  return true;
}

bool tc_net_supported(const tpx_fcn_desc* net) {
  static const int disabled = env_flag("TPX_DISABLE_TC");
  return !disabled && tc_net_fits(net);
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

template <int ND, int LAP, int ACT>
static int launch_tc_fwd_act(const FwdArgs& a, int sm_count, cudaStream_t stream) {
  const FwdSmem so = fwd_smem(a.lay);
  const size_t smem = (size_t)so.total * sizeof(float) + 1024;
  if (smem > 227 * 1024) return TPX_E_UNSUPPORTED;
  if (cudaFuncSetAttribute(tc_fwd_kernel<ND, LAP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TPX_E_CUDA;
  const int64_t ntiles = (a.n + TILE - 1) / TILE, npairs = (ntiles + SLOTS - 1) / SLOTS;
  int64_t grid = sm_count;
  if (grid > npairs) grid = npairs;
  if (grid < 1) return 0;
  tc_fwd_kernel<ND, LAP, ACT><<<(unsigned)grid, THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : TPX_E_CUDA;
}
static thread_local int g_fwd_act = TPX_ACT_TANH;      // activation of the network being launched (set by tc_forward)
template <int ND, int LAP>
static int launch_tc_fwd(const FwdArgs& a, int sm_count, cudaStream_t stream) {
  return g_fwd_act == TPX_ACT_SIN ? launch_tc_fwd_act<ND, LAP, 1>(a, sm_count, stream) : launch_tc_fwd_act<ND, LAP, 0>(a, sm_count, stream);
}

int tc_forward(const tpx_fcn_desc* net, const JetArgs& jet, int lap, const float* img, const float* x, int64_t n,
               float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream) {
  FwdArgs a;
  a.lay = TcLayout{net->n_hidden, net->d_in, net->d_out};
  a.jet = jet;
  a.img = img; a.x = x; a.n = n; a.u = u; a.du = du; a.lap = lapo;
  a.ckpt = saved;
  a.dbg = g_tc_debug;
  g_fwd_act = net->activation;
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
