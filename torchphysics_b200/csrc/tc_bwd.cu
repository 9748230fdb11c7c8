// Tensor-core (tcgen05 / TMEM) reverse kernel for hidden width <= 64, tanh or sin, <= 4 streams.
//
// It sweeps the layers of one 32-point tile backwards, reading the activation-jet checkpoints the forward
// kernel saved (tc_kernels.cu: per tile and layer S arrays [64 neurons][32 points]: tanh(z), d_k z, L z).
// Rows of every UMMA are (stream q, point p) = TMEM lane 32 q + p, as in the forward kernel.
//
// Per hidden matrix l (1 <= l < L) the reverse step issues two products:
//   adjoint   D[128 x 64] = Zbar_l[128 x 64] W_l            A = Zbar (hi/lo) in TMEM, B = W_l^T image (K-major)
//   gradient  dW_l[j, i] += sum_r Zbar_l[r, j] a_{l-1}[r, i] r = all 128 rows (4 streams x 32 points) is the K
//             dimension: both operands are written TRANSPOSED into shared memory (row = neuron, K = row
//             index, 128-byte swizzle; one conflict-free scalar store per element), Zbar stacked as
//             [hi ; lo] on M = 128 and a_{l-1} as two N = 64 operands hi / lo, so two chains give
//             hi*hi + lo*hi + hi*lo + lo*lo.  dW_l accumulates in TMEM over ALL tiles of the CTA and is
//             read out once at the end of the kernel.
//
// Activation step (16 warps = 4 neuron groups x 4 TMEM quadrants, lane = point).  A TMEM quadrant holds ONE
// stream, but the adjoint formulas couple all streams of a point, so the step is split like the forward's:
//   in      the OWNER warp of stream s reads its 16 adjoint columns of D and writes them to the group's
//           exchange buffer ex[s][p][neuron] (rows padded to 20 floats: conflict-free 128-bit accesses)
//   compute warp w of the group takes neurons [4 w, 4 w + 4) for ALL streams of its 32 points: adjoints of the
//           pre-activation streams, the activation jets a_{l-1} of the previous layer (from its checkpoints,
//           kept in registers for the next step) -- all thread-local, identical work on every scheduler;
//           it writes the transposed gradient-product operands (Zt, At) itself and hands Zbar back through ex
//   out     the owner splits its stream's Zbar rows hi / lo into the TMEM A operand of the adjoint product
// Bias, first-layer and output-layer gradients are reduced over the 32 lanes with a short transposing
// shuffle reduction and accumulated in shared memory (each entry is owned by one warp).
// Weight images (32 KB per product) are fetched with cp.async.bulk + mbarrier behind the activation work.
#include <cstdio>
#include <cstring>

#include "fcn_common.cuh"
#include "f32x2.cuh"
#include "tc_common.cuh"
#include "tcp_common.cuh"
#include "tc_kernels.h"
#include "tc_layout.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tc {

constexpr int B_CH = 16;                 // neurons per warp group (= columns one owner moves per step)
constexpr int B_GROUPS = HP / B_CH;      // warp groups: the 4 quadrant warps that share one chunk of neurons
constexpr int B_NW = B_CH / 4;           // neurons per warp in the compute phase
constexpr int B_EPI_WARPS = 4 * B_GROUPS;
// 16 activation warps, the MMA warp and 3 idle warps that complete its warpgroup: setmaxnreg moves registers between whole
// warpgroups, and a 17-warp CTA is otherwise held to 96 registers per thread (5 warps on one scheduler's register file)
constexpr int B_THREADS = 32 * 20;
constexpr int B_ACT_REGS = 104, B_MMA_REGS = 64;     // 4 x 128 x (104 - 96) taken = 128 x (96 - 64) released
constexpr uint32_t B_TMEM_COLS = 512;
constexpr uint32_t T_D = 0, T_AH = 64, T_AL = 128, T_DW = 192;   // TMEM columns
// The TMEM accumulators of dW round toward zero on every add (one ulp of the running sum): over the ~450 tiles of a CTA
// at 2^21 points the sums came out ~1e-4 short (scripts/dw_scale_check.py).  Every FLUSH_TILES tiles the activation warps
// move them into per-CTA fp32 partial sums in global memory (RED.ADD.F32, round to nearest) and the products restart from 0.
constexpr int FLUSH_TILES = 8;
constexpr int MAX_BWD_CTAS = 160;        // the partial sums are sized for this many CTAs (B200: 148 SMs)
constexpr int MAXL = 6;                  // hidden layers supported (dW accumulators: 64 columns each, L - 1 <= 5)
constexpr int MAXIO = 4;                 // d_in, d_out supported
// dW accumulators: the hidden matrices processed first get N = 128 accumulators ([a_hi | a_lo] as ONE B operand, so the
// stacked Zt operand is read once per K step: 16 MMAs of 64 cycles instead of 2 x 16 of 48) as far as the 320 TMEM
// columns behind D | A_hi | A_lo allow; the rest keep N = 64 and two chains.
__host__ __device__ inline int dw_wide_count(int L) { const int n = L - 1, spare = (320 - 64 * n) / 64; return spare < 0 ? 0 : (spare < n ? spare : n); }
__host__ __device__ inline bool dw_is_wide(int L, int l) { return (L - 1 - l) < dw_wide_count(L); }            // l = L-1 first
__host__ __device__ inline uint32_t dw_col(int L, int l) {                                                      // columns from T_DW
  uint32_t c = 0;
  for (int m = L - 1; m > l; --m) c += dw_is_wide(L, m) ? 128u : 64u;
  return c;
}
constexpr int BXP = 20;                  // padded neuron extent of one (stream, point) row of the exchange buffer
constexpr int BX_GROUP = 4 * TILE * BXP;

struct BwdSmem {
  int wring, zt, at, ex, acc, small, total;   // float offsets from the 1 KB aligned base
};
// shared accumulators: [db: MAXL x 64][dW0: MAXIO x 64][dWL: MAXIO x 64][dbL: 64 (MAXIO used)]
constexpr int ACC_DB = 0, ACC_W0 = MAXL * HP, ACC_WL = ACC_W0 + MAXIO * HP, ACC_BL = ACC_WL + MAXIO * HP, ACC_TOTAL = ACC_BL + HP;
__host__ __device__ inline BwdSmem bwd_smem(const TcLayout& lay) {
  BwdSmem s;
  s.wring = 0;                            // (hi, lo) x 4096: one product image, refetched per phase
  s.zt = 8192;                            // 128 rows x 128 K
  s.at = s.zt + 16384;                    // 128 rows (a_hi 0-63, a_lo 64-127) x 128 K
  s.ex = s.at + 16384;
  s.acc = s.ex + B_GROUPS * BX_GROUP;
  s.small = s.acc + ACC_TOTAL;
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
  const float* ckpt;   // per tile: L x 4 arrays x 64 x 32 floats
  float* dwpart;       // per CTA: (MAXL - 1) x 64 x 64 fp32 partial sums of the hidden dW, [matrix][input neuron i][output neuron j]
  long long* dbg;      // development: timestamps of one tile
};

// lane receives the sum over all 32 lanes of v[lane & 3]
__device__ __forceinline__ float reduce4(float v0, float v1, float v2, float v3, int lane) {
  const bool up2 = (lane & 2) != 0;
  const float a0 = (up2 ? v2 : v0) + __shfl_xor_sync(0xffffffffu, up2 ? v0 : v2, 2);
  const float a1 = (up2 ? v3 : v1) + __shfl_xor_sync(0xffffffffu, up2 ? v1 : v3, 2);
  const bool up1 = (lane & 1) != 0;
  float b = (up1 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up1 ? a0 : a1, 1);
  b += __shfl_xor_sync(0xffffffffu, b, 4);
  b += __shfl_xor_sync(0xffffffffu, b, 8);
  b += __shfl_xor_sync(0xffffffffu, b, 16);
  return b;
}

__device__ __forceinline__ void bulk_load(float* dst_smem, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int ND, int LAP, int ACT>       // ACT: 0 = tanh (checkpoint keeps t = tanh z), 1 = sin (keeps z)
__global__ void __launch_bounds__(B_THREADS, 1) tc_bwd_kernel(BwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a, bar_z, bar_d, bar_f, bar_w;
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const TcLayout lay = a.lay;
  const Layout lay32 = a.lay32;
  const BwdSmem so = bwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nphase = L - 1;                 // MMA phases per tile: R(L-1) .. R1

  for (int v = tid; v < (int)lay.small(); v += B_THREADS) sm[so.small + v] = __ldg(a.img + v);
  for (int v = tid; v < ACC_TOTAL; v += B_THREADS) sm[so.acc + v] = 0.0f;
  for (int v = tid; v < 2 * 16384; v += B_THREADS) sm[so.zt + v] = 0.0f;     // K atoms of unused streams stay zero
  float* dwp = a.dwpart + (size_t)blockIdx.x * ((MAXL - 1) * HP * HP);
  for (int v = tid; v < (L - 1) * HP * HP; v += B_THREADS) dwp[v] = 0.0f;
  if (warp == B_EPI_WARPS) tmem_alloc(&tmem_slot, B_TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar_a, B_EPI_WARPS);
    mbar_init(&bar_z, B_EPI_WARPS);
    mbar_init(&bar_d, 1);
    mbar_init(&bar_f, 1);
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp > B_EPI_WARPS) {
    tcp::reg_dec<B_MMA_REGS>();          // idle warps of the MMA warpgroup
  } else if (warp == B_EPI_WARPS) {
    tcp::reg_dec<B_MMA_REGS>();
    // =========================== weight producer + MMA issuer ===========================
    const uint32_t ring = smem_u32(sm + so.wring), zt = smem_u32(sm + so.zt), at = smem_u32(sm + so.at);
    constexpr uint32_t idesc = idesc_tf32(128, HP);
    const int64_t total_phases = my_tiles * nphase;
    uint32_t ph_a = 0, ph_z = 0, ph_w = 0, ph_d = 0;
    auto phase_src = [&](int64_t g) -> const float* {
      const int l = (L - 1) - (int)(g % nphase);
      return a.img + lay.mat(l, 2);
    };
    if (total_phases > 0 && elect_one()) bulk_load(sm + so.wring, phase_src(0), 32768u, &bar_w);
    __syncwarp();
    for (int64_t g = 0; g < total_phases; ++g) {
      const int l = (L - 1) - (int)(g % nphase);
      const bool first_tile = ((g / nphase) % FLUSH_TILES) == 0;   // the accumulators were flushed before this tile
      mbar_wait(&bar_a, ph_a);
      ph_a ^= 1;
#ifdef TPX_TIMELINE
      if (a.dbg && blockIdx.x == 0 && g >= 2 * nphase && g < 3 * nphase && lane == 0) a.dbg[512 + l * 4 + 3] = clock64();
#endif
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after();
#ifdef TPX_TIMELINE
      const bool mdbg = a.dbg && blockIdx.x == 0 && g >= 2 * nphase && g < 3 * nphase && lane == 0;
      if (mdbg) a.dbg[512 + l * 4 + 0] = clock64();
#endif
      if (elect_one()) {
        const uint32_t wh = ring, wl = wh + 16384u;
        const uint32_t tD = tmem + T_D, tAh = tmem + T_AH, tAl = tmem + T_AL;
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
        mma_commit(&bar_d);                  // the adjoint product is all the next activation step needs ...
      }
      __syncwarp();
      // the weight image is free once the adjoint product is done: fetch the next image behind the activation warps' work
      mbar_wait(&bar_d, ph_d);
      ph_d ^= 1;
#ifdef TPX_TIMELINE
      if (mdbg) a.dbg[512 + l * 4 + 1] = clock64();
#endif
      if (g + 1 < total_phases) {
        if (elect_one()) bulk_load(sm + so.wring, phase_src(g + 1), 32768u, &bar_w);
        __syncwarp();
      }
      // ... and it only needs the TMEM A operand: the activation warps write the transposed operands of the gradient
      // product (the longest phase of a step) while it runs, and signal them separately
      mbar_wait(&bar_z, ph_z);
      ph_z ^= 1;
      fence_after();
      if (elect_one()) {
        const uint32_t tW = tmem + T_DW + dw_col(L, l);
        if (dw_is_wide(L, l)) {
          constexpr uint32_t idesc_w = idesc_tf32(128, 2 * HP);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ss(tW, desc_k_sw128(zt + (ks >> 2) * 16384 + (ks & 3) * 32), desc_k_sw128(at + (ks >> 2) * 16384 + (ks & 3) * 32),
                        idesc_w, (ks > 0 || !first_tile) ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ss(tW, desc_k_sw128(zt + (ks >> 2) * 16384 + (ks & 3) * 32), desc_k_sw128(at + (ks >> 2) * 16384 + (ks & 3) * 32),
                        idesc, (ks > 0 || !first_tile) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ss(tW, desc_k_sw128(zt + (ks >> 2) * 16384 + (ks & 3) * 32),
                        desc_k_sw128(at + 8192u + (ks >> 2) * 16384 + (ks & 3) * 32), idesc, 1);
        }
        mma_commit(&bar_f);                  // the gradient product only gates the reuse of its operands
      }
      __syncwarp();
#ifdef TPX_TIMELINE
      if (mdbg) a.dbg[512 + l * 4 + 2] = clock64();
#endif
    }
  } else {
    tcp::reg_inc<B_ACT_REGS>();
    // =========================== activation warps ===========================
    // warp = grp * 4 + quad: TMEM quadrant quad (= stream quad), group neurons [16 grp, 16 grp + 16);
    // compute phase: neurons jn .. jn + 3, all streams
    const int grp = warp >> 2, quad = warp & 3;
    const int jc = grp * B_CH, jn = jc + quad * B_NW;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tD = tmem + T_D + lane_base, tAh = tmem + T_AH + lane_base, tAl = tmem + T_AL + lane_base;
    float* ex = sm + so.ex + grp * BX_GROUP;
    float* ex_own = ex + (quad * TILE + lane) * BXP;           // owner view: stream quad, the group's 16 neurons
    float* ex_cmp = ex + lane * BXP + quad * B_NW;             // compute view of stream s: + s * TILE * BXP
    float* acc = sm + so.acc;
    const float* W0 = sm + so.small + lay.w0();
    const float* WL = sm + so.small + lay.wl();
    const int barid = 1 + grp;
    const bool owner = quad < S;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    uint32_t ph_d = 0, ph_f = 0;
    bool dw_pending = false;             // a gradient product may still be reading Zt / At
    bool flush_now = false;              // first step of a tile that follows FLUSH_TILES tiles of accumulation
    int tcount = 0;
    // this thread's slice of the hidden dW accumulators -> the CTA's fp32 partial sums (hi rows j and lo rows 64 + j of a
    // matrix, and the a_lo columns 64 + i of a wide accumulator, all add into element (j, i))
    auto flush_dw = [&]() {
      const int j = (quad & 1) * 32 + lane;
      for (int l = 1; l < L; ++l) {
        const uint32_t tW = tmem + T_DW + dw_col(L, l) + lane_base;
        float w[B_CH];
        tmem_ld16(tW + jc, w);
        if (dw_is_wide(L, l)) {
          float w2[B_CH];
          tmem_ld16(tW + HP + jc, w2);
#pragma unroll
          for (int i = 0; i < B_CH; ++i) w[i] += w2[i];
        }
#pragma unroll
        for (int i = 0; i < B_CH; ++i) atomicAdd(dwp + ((size_t)(l - 1) * HP + jc + i) * HP + j, w[i]);
      }
    };
    // transposed operands: element (row = neuron, K = 32 s + lane); 128-byte swizzle:
    //   offset = s * atom + row * 32 + ((lane / 4) ^ (row % 8)) * 4 + lane % 4
    float* Zt = sm + so.zt + jn * 32 + (lane & 3);
    float* At = sm + so.at + jn * 32 + (lane & 3);
    int xo[B_NW];
#pragma unroll
    for (int i = 0; i < B_NW; ++i) xo[i] = ((lane >> 2) ^ ((jn + i) & 7)) << 2;
    if (!owner) {                        // rows of an unused quadrant: a zero A operand, written once
      float zero[B_CH];
#pragma unroll
      for (int i = 0; i < B_CH; ++i) zero[i] = 0.0f;
      tmem_st16(tAh + jc, zero);
      tmem_st16(tAl + jc, zero);
      tmem_wait_st();
    }

    // activation jets of a layer from its checkpoints (t, d_k z, L z) -> (a, d_k a, L a)
    auto jets = [&](const float (&c)[S][B_NW], float (&h)[S][B_NW]) {
      if constexpr (ACT == 0) {              // tanh: neuron pairs in packed fp32 (f32x2.cuh), bit-identical to the scalar formulas
#pragma unroll
        for (int i = 0; i < B_NW; i += 2) {
          const F2 t = f2(c[0][i], c[0][i + 1]);
          const F2 s1 = fma2(neg(t), t, f2(1.0f));
          const F2 s2 = (f2(-2.0f) * t) * s1;
          h[0][i] = t.v.x; h[0][i + 1] = t.v.y;
          F2 qq = f2(0.0f);
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const F2 dz = f2(c[1 + k][i], c[1 + k][i + 1]);
            if (LAP) qq = fma2(f2(lapw[k]) * dz, dz, qq);
            const F2 hk = s1 * dz;
            h[1 + k][i] = hk.v.x; h[1 + k][i + 1] = hk.v.y;
          }
          if (LAP) {
            const F2 hl = fma2(s2, qq, s1 * f2(c[S - 1][i], c[S - 1][i + 1]));
            h[S - 1][i] = hl.v.x; h[S - 1][i + 1] = hl.v.y;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < B_NW; ++i) {
          float t, s1, s2;
          sincosf(c[0][i], &t, &s1);
          s2 = -t;
          h[0][i] = t;
          float qq = 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const float dz = c[1 + k][i];
            if (LAP) qq = fmaf(lapw[k] * dz, dz, qq);
            h[1 + k][i] = s1 * dz;
          }
          if (LAP) h[S - 1][i] = fmaf(s2, qq, s1 * c[S - 1][i]);
        }
      }
    };
    auto load_ck = [&](const float* ck_tile, int l, float (&c)[S][B_NW]) {
      const float* base = ck_tile + (size_t)l * 4 * HP * TILE + jn * TILE + lane;
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < B_NW; ++i) c[s][i] = __ldcs(base + (s * HP + i) * TILE);
    };

    // adjoints of the pre-activation streams of a layer from the adjoints of its activation streams (thread-local)
    auto zbar = [&](const float (&c)[S][B_NW], const float (&ga)[S][B_NW], float (&zb)[S][B_NW]) {
      if constexpr (ACT == 0) {              // tanh: neuron pairs in packed fp32, bit-identical to the scalar formulas below
#pragma unroll
        for (int i = 0; i < B_NW; i += 2) {
          const F2 t = f2(c[0][i], c[0][i + 1]);
          const F2 s1 = fma2(neg(t), t, f2(1.0f));
          const F2 s2 = (f2(-2.0f) * t) * s1;
          F2 v0 = f2(ga[0][i], ga[0][i + 1]) * s1, qq = f2(0.0f);
          const F2 gl = LAP ? f2(ga[S - 1][i], ga[S - 1][i + 1]) : f2(0.0f);
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const F2 dz = f2(c[1 + k][i], c[1 + k][i + 1]);
            const F2 gk = f2(ga[1 + k][i], ga[1 + k][i + 1]);
            v0 = fma2(gk * s2, dz, v0);
            F2 vk = gk * s1;
            if (LAP) {
              qq = fma2(f2(lapw[k]) * dz, dz, qq);
              vk = fma2((f2(2.0f * lapw[k]) * gl) * s2, dz, vk);
            }
            zb[1 + k][i] = vk.v.x; zb[1 + k][i + 1] = vk.v.y;
          }
          if (LAP) {
            const F2 s3 = (f2(-2.0f) * s1) * fma2(f2(-3.0f) * t, t, f2(1.0f));
            v0 = fma2(gl, fma2(s3, qq, s2 * f2(c[S - 1][i], c[S - 1][i + 1])), v0);
            const F2 zl = gl * s1;
            zb[S - 1][i] = zl.v.x; zb[S - 1][i + 1] = zl.v.y;
          }
          zb[0][i] = v0.v.x; zb[0][i + 1] = v0.v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < B_NW; ++i) {
          float t, s1, s2;
          sincosf(c[0][i], &t, &s1);
          s2 = -t;
          float v0 = ga[0][i] * s1, qq = 0.0f;
          const float gl = LAP ? ga[S - 1][i] : 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            const float dz = c[1 + k][i];
            v0 = fmaf(ga[1 + k][i] * s2, dz, v0);
            float vk = ga[1 + k][i] * s1;
            if (LAP) {
              qq = fmaf(lapw[k] * dz, dz, qq);
              vk = fmaf(2.0f * lapw[k] * gl * s2, dz, vk);
            }
            zb[1 + k][i] = vk;
          }
          if (LAP) {
            const float s3 = -s1;
            v0 = fmaf(gl, fmaf(s3, qq, s2 * c[S - 1][i]), v0);
            zb[S - 1][i] = gl * s1;
          }
          zb[0][i] = v0;
        }
      }
    };
    // in: adjoint product of the previous step, rows of this warp's stream -> exchange buffer -> all streams of
    // this thread's neurons
#ifdef TPX_TIMELINE
    long long* dbgp = nullptr;
#endif
#ifdef TPX_TIMELINE
#define TPX_STAMP(e) do { if (dbgp) dbgp[e] = clock64(); } while (0)
#else
#define TPX_STAMP(e) do { } while (0)
#endif
    auto fetch_adjoint = [&](float (&ga)[S][B_NW]) {
      TPX_STAMP(0);
      mbar_wait(&bar_d, ph_d);
      ph_d ^= 1;
      fence_after();
      TPX_STAMP(1);
      if (owner) {
        float w[B_CH];
        tmem_ld16(tD + jc, w);
#pragma unroll
        for (int v = 0; v < B_CH / 4; ++v)
          *reinterpret_cast<float4*>(ex_own + 4 * v) = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
      }
      named_bar_sync(barid, 128);
      TPX_STAMP(2);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float4 v4 = *reinterpret_cast<const float4*>(ex_cmp + s * TILE * BXP);
        ga[s][0] = v4.x; ga[s][1] = v4.y; ga[s][2] = v4.z; ga[s][3] = v4.w;
      }
    };
    // hand Zbar of layer l >= 1 to the two products: transposed operands of the gradient product (with the
    // activation jets of layer l - 1), TMEM A operand of the adjoint product
    auto emit = [&](const float (&zb)[S][B_NW], const float (&cprev)[S][B_NW]) {
      // (1) Zbar -> exchange buffer -> TMEM A operand of the adjoint product, which starts as soon as all warps arrive
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<float4*>(ex_cmp + s * TILE * BXP) = make_float4(zb[s][0], zb[s][1], zb[s][2], zb[s][3]);
      TPX_STAMP(3);
      named_bar_sync(barid, 128);
      if (owner) {
        float hi[B_CH], lo[B_CH];
#pragma unroll
        for (int v = 0; v < B_CH / 4; ++v) {
          const float4 w4 = *reinterpret_cast<const float4*>(ex_own + 4 * v);
          hi[4 * v] = w4.x; hi[4 * v + 1] = w4.y; hi[4 * v + 2] = w4.z; hi[4 * v + 3] = w4.w;
        }
#pragma unroll
        for (int i = 0; i < B_CH; i += 2) {
          const F2 l = fma2(f2(-1.0f), f2(tf32_trunc(hi[i]), tf32_trunc(hi[i + 1])), f2(hi[i], hi[i + 1]));      // x - trunc(x), exact
          lo[i] = l.v.x; lo[i + 1] = l.v.y;            // hi stays the raw word: the MMA truncates it (tc_common.cuh)
        }
        tmem_st16(tAh + jc, hi);
        tmem_st16(tAl + jc, lo);
        tmem_wait_st();
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a);
      TPX_STAMP(4);
      // (2) transposed operands of the gradient product, in the shadow of the adjoint product
      float ap[S][B_NW];
      jets(cprev, ap);
      if (dw_pending) {                  // the previous gradient product must have finished reading Zt / At
        mbar_wait(&bar_f, ph_f);
        ph_f ^= 1;
      }
      if (flush_now) {                   // every gradient product so far has completed; the next ones start from zero
        fence_after();
        flush_dw();
        fence_before();
        flush_now = false;
      }
      TPX_STAMP(5);
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < B_NW; i += 2) {
          const float zh0 = zb[s][i], zh1 = zb[s][i + 1];                                // raw words: the MMA truncates them
          const F2 zl = fma2(f2(-1.0f), f2(tf32_trunc(zh0), tf32_trunc(zh1)), f2(zh0, zh1));   // x - trunc(x), exact
          Zt[s * 4096 + i * 32 + xo[i]] = zh0;
          Zt[s * 4096 + (i + 1) * 32 + xo[i + 1]] = zh1;
          Zt[s * 4096 + (64 + i) * 32 + xo[i]] = zl.v.x;
          Zt[s * 4096 + (64 + i + 1) * 32 + xo[i + 1]] = zl.v.y;
          const float ah0 = ap[s][i], ah1 = ap[s][i + 1];
          const F2 al = fma2(f2(-1.0f), f2(tf32_trunc(ah0), tf32_trunc(ah1)), f2(ah0, ah1));
          At[s * 4096 + i * 32 + xo[i]] = ah0;
          At[s * 4096 + (i + 1) * 32 + xo[i + 1]] = ah1;
          At[s * 4096 + (64 + i) * 32 + xo[i]] = al.v.x;
          At[s * 4096 + (64 + i + 1) * 32 + xo[i + 1]] = al.v.y;
        }
      dw_pending = true;
      TPX_STAMP(6);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_z);
      TPX_STAMP(7);
#ifdef TPX_TIMELINE
      if (dbgp) dbgp -= 16;
#endif
    };
    auto add_bias = [&](int l, const float (&zb)[S][B_NW]) {
      const float r = reduce4(zb[0][0], zb[0][1], zb[0][2], zb[0][3], lane);
      if (lane < B_NW) acc[ACC_DB + l * HP + jn + lane] += r;
    };

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      flush_now = tcount > 0 && tcount % FLUSH_TILES == 0;
      ++tcount;
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
#ifdef TPX_TIMELINE
      dbgp = (a.dbg && blockIdx.x == 0 && tile == blockIdx.x + 2 * (int64_t)gridDim.x && lane == 0 && (warp == 0 || warp == 7))
                 ? a.dbg + (warp == 0 ? 0 : 256) + (L - 1) * 16 : nullptr;
#endif
      const float* ck_tile = a.ckpt + (size_t)tile * ((size_t)L * 4 * HP * TILE);
      float ck[S][B_NW], ckn[S][B_NW];       // checkpoints of the current layer and of the one below (prefetched)
      load_ck(ck_tile, L - 1, ck);
      load_ck(ck_tile, L - 2, ckn);
      // ------------------------------------------------ output layer + last hidden layer (l = L - 1)
      {
        float ga[S][B_NW], zb[S][B_NW];
        {
          float g[S][MAXIO];                 // adjoints of this point's output streams
#pragma unroll
          for (int o = 0; o < MAXIO; ++o) {
            const bool on = valid && o < d_out;
            g[0][o] = (on && a.gu) ? __ldg(a.gu + gp * d_out + o) : 0.0f;
#pragma unroll
            for (int k = 0; k < ND; ++k) g[1 + k][o] = (on && a.gdu) ? __ldg(a.gdu + (gp * d_out + o) * ND + k) : 0.0f;
            if (LAP) g[S - 1][o] = (on && a.glap) ? __ldg(a.glap + gp * d_out + o) : 0.0f;
          }
          if (warp == 0) {                   // output bias: sum of the value-stream adjoints
#pragma unroll
            for (int o = 0; o < MAXIO; ++o) {
              float sum = g[0][o];
#pragma unroll
              for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
              if (lane == 0 && o < d_out) acc[ACC_BL + o] += sum;
            }
          }
          // u_s = WL a_s (+ bL): adjoint and gradient are thread-local
          float al[S][B_NW];
          jets(ck, al);
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < B_NW; ++i) ga[s][i] = 0.0f;
#pragma unroll
          for (int o = 0; o < MAXIO; ++o)
            if (o < d_out) {
              float wg[B_NW];
#pragma unroll
              for (int i = 0; i < B_NW; ++i) wg[i] = 0.0f;
#pragma unroll
              for (int s = 0; s < S; ++s)
#pragma unroll
                for (int i = 0; i < B_NW; ++i) {
                  ga[s][i] = fmaf(WL[o * HP + jn + i], g[s][o], ga[s][i]);
                  wg[i] = fmaf(g[s][o], al[s][i], wg[i]);
                }
              const float r = reduce4(wg[0], wg[1], wg[2], wg[3], lane);
              if (lane < B_NW) acc[ACC_WL + o * HP + jn + lane] += r;
            }
        }
        zbar(ck, ga, zb);
        add_bias(L - 1, zb);
        emit(zb, ckn);
      }
      // ------------------------------------------------ hidden layers l = L - 2 .. 1 (ck holds layer l)
      for (int l = L - 2; l >= 1; --l) {
        float ga[S][B_NW], zb[S][B_NW];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int i = 0; i < B_NW; ++i) ck[s][i] = ckn[s][i];
        load_ck(ck_tile, l - 1, ckn);        // behind the wait for the adjoint product
        fetch_adjoint(ga);
        zbar(ck, ga, zb);
        add_bias(l, zb);
        emit(zb, ckn);
      }
      // ------------------------------------------------ first layer: z = W0 x + b0, d_k z = W0[:, dcol_k], L z = 0
      {
        float ga[S][B_NW], zb[S][B_NW];
        fetch_adjoint(ga);
        zbar(ckn, ga, zb);
        add_bias(0, zb);
        float rk[NDX];
#pragma unroll
        for (int k = 0; k < ND; ++k) rk[k] = reduce4(zb[1 + k][0], zb[1 + k][1], zb[1 + k][2], zb[1 + k][3], lane);
#pragma unroll
        for (int cc = 0; cc < MAXIO; ++cc)
          if (cc < d_in) {
            const float xc = valid ? __ldg(a.x + gp * d_in + cc) : 0.0f;
            float r = reduce4(zb[0][0] * xc, zb[0][1] * xc, zb[0][2] * xc, zb[0][3] * xc, lane);
#pragma unroll
            for (int k = 0; k < ND; ++k)
              if (dcol[k] == cc) r += rk[k];
            if (lane < B_NW) acc[ACC_W0 + cc * HP + jn + lane] += r;
          }
      }
    }
    if (dw_pending) {                        // the last gradient product of this CTA
      mbar_wait(&bar_f, ph_f);
      fence_after();
    }

    // ------------------------------------------------------------------ flush the accumulators
    const int HP32 = lay32.HP;
    __syncwarp();
    if (lane < B_NW) {
      const int j = jn + lane;
      if (j < HP32) {
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
          if (l < L) atomicAdd(a.gacc + (l == 0 ? lay32.b0() : lay32.b(l)) + j, (double)acc[ACC_DB + l * HP + j]);
#pragma unroll
        for (int cc = 0; cc < MAXIO; ++cc)
          if (cc < d_in) atomicAdd(a.gacc + lay32.w0() + (size_t)cc * HP32 + j, (double)acc[ACC_W0 + cc * HP + j]);
#pragma unroll
        for (int o = 0; o < MAXIO; ++o)
          if (o < d_out) atomicAdd(a.gacc + lay32.wl() + (size_t)o * HP32 + j, (double)acc[ACC_WL + o * HP + j]);
      }
    }
    if (warp == 0 && lane == 0) {
#pragma unroll
      for (int o = 0; o < MAXIO; ++o)
        if (o < d_out) atomicAdd(a.gacc + lay32.bl() + o, (double)acc[ACC_BL + o]);
    }
    // hidden matrices: what is left in the TMEM accumulators joins the CTA's fp32 partial sums (a warp adds 128 consecutive bytes
    // per instruction); tc_dw_reduce_kernel sums the CTAs in float64.  (One float64 atomic per element, quadrant and CTA straight
    // into the gradient image -- 32 rows 512 bytes apart per warp instruction -- was a fixed ~60 us per launch.)
    if (my_tiles > 0) flush_dw();
  }
  fence_before();
  __syncthreads();
  if (warp == B_EPI_WARPS) tmem_dealloc(tmem, B_TMEM_COLS);
}

// hidden dW: float64 sum over CTAs of the fp32 partial sums [CTA][matrix][input neuron i][output neuron j], added to the gradient
// image (one writer per element; launches of one sweep are ordered on the stream)
__global__ void __launch_bounds__(256) tc_dw_reduce_kernel(const float* __restrict__ dwpart, int nctas, int L, Layout lay32,
                                                           double* __restrict__ gacc) {
  const int e = blockIdx.x * 256 + threadIdx.x;              // = ((l - 1) * HP + i) * HP + j
  if (e >= (L - 1) * HP * HP) return;
  const int j = e % HP, i = (e / HP) % HP, l = 1 + e / (HP * HP);
  if (j >= lay32.HP || i >= lay32.HP) return;
  double acc = 0.0;
  int c = 0;
  constexpr size_t PER = (size_t)(MAXL - 1) * HP * HP;
  for (; c + 4 <= nctas; c += 4) {
    const float v0 = __ldcg(dwpart + (size_t)c * PER + e), v1 = __ldcg(dwpart + (size_t)(c + 1) * PER + e);
    const float v2 = __ldcg(dwpart + (size_t)(c + 2) * PER + e), v3 = __ldcg(dwpart + (size_t)(c + 3) * PER + e);
    acc += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
  }
  for (; c < nctas; ++c) acc += (double)__ldcg(dwpart + (size_t)c * PER + e);
  gacc[lay32.w(l) + (size_t)j * lay32.HP + i] += acc;
}

}  // namespace tc

using namespace tc;

bool tc_bwd_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tc_supported(net, nd, lap)) return false;
  if (net->n_hidden > MAXL || net->d_in > MAXIO || net->d_out > MAXIO) return false;
  return true;
}

static size_t ckpt_floats_per_tile(int L) { return (size_t)L * 4 * HP * TILE; }

// without saved checkpoints the reverse sweep runs in chunks: forward kernel (checkpoints only) into the
// workspace, then the reverse kernel; a chunk is 16 tiles per SM
static int64_t chunk_tiles(int sm_count) { return (int64_t)sm_count * 16; }

// behind the checkpoints: the per-CTA fp32 partial sums of the hidden dW
static size_t dwpart_bytes() { return (size_t)MAX_BWD_CTAS * (MAXL - 1) * HP * HP * sizeof(float); }

size_t tc_bwd_workspace_bytes(const tpx_fcn_desc* net, int sm_count) {
  return (size_t)chunk_tiles(sm_count) * ckpt_floats_per_tile(net->n_hidden) * sizeof(float) + dwpart_bytes();
}

size_t tc_saved_bytes(const tpx_fcn_desc* net, int64_t n) {
  return (size_t)((n + TILE - 1) / TILE) * ckpt_floats_per_tile(net->n_hidden) * sizeof(float) + dwpart_bytes();
}

template <int ND, int LAP, int ACT>
static int launch_tc_bwd_act(const BwdArgs& a, int sm_count, cudaStream_t stream) {
  const size_t smem = (size_t)bwd_smem(a.lay).total * sizeof(float) + 1024;
  if (smem > 227 * 1024) return fail(TPX_E_UNSUPPORTED, "tc_bwd_kernel needs %zu bytes of shared memory", smem);
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  int64_t grid = sm_count < ntiles ? sm_count : ntiles;
  if (grid > MAX_BWD_CTAS) grid = MAX_BWD_CTAS;
  if (grid < 1) return 0;
  cudaError_t e = cudaFuncSetAttribute(tc_bwd_kernel<ND, LAP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tc_bwd_kernel: shared memory attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
  tc_bwd_kernel<ND, LAP, ACT><<<(unsigned)grid, B_THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tc_bwd_kernel launch (%lld CTAs x %d threads, %zu bytes smem): %s", (long long)grid, B_THREADS, smem, cudaGetErrorString(e));
  const int L = a.lay.L;
  tc_dw_reduce_kernel<<<((L - 1) * HP * HP + 255) / 256, 256, 0, stream>>>(a.dwpart, (int)grid, L, a.lay32, a.gacc);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tc_dw_reduce_kernel launch: %s", cudaGetErrorString(e));
  return 0;
}

static thread_local int g_bwd_act = TPX_ACT_TANH;      // activation of the network being launched (set by tc_backward)
template <int ND, int LAP>
static int launch_tc_bwd(const BwdArgs& a, int sm_count, cudaStream_t stream) {
  return g_bwd_act == TPX_ACT_SIN ? launch_tc_bwd_act<ND, LAP, 1>(a, sm_count, stream) : launch_tc_bwd_act<ND, LAP, 0>(a, sm_count, stream);
}

static int launch_tc_bwd_any(const BwdArgs& a, int nd, int lap, int sm_count, cudaStream_t stream) {
  switch (nd * 2 + lap) {
    case 0: return launch_tc_bwd<0, 0>(a, sm_count, stream);
    case 2: return launch_tc_bwd<1, 0>(a, sm_count, stream);
    case 3: return launch_tc_bwd<1, 1>(a, sm_count, stream);
    case 4: return launch_tc_bwd<2, 0>(a, sm_count, stream);
    case 5: return launch_tc_bwd<2, 1>(a, sm_count, stream);
    case 6: return launch_tc_bwd<3, 0>(a, sm_count, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

int tc_backward(const tpx_fcn_desc* net, const Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* ckpt, size_t ckpt_bytes,
                int saved, int sm_count, cudaStream_t stream) {
  BwdArgs a;
  a.lay = TcLayout{net->n_hidden, net->d_in, net->d_out};
  a.lay32 = lay32;
  a.jet = jet;
  a.img = img; a.gacc = gacc; a.ckpt = ckpt;
  a.dbg = g_tc_debug;
  g_bwd_act = net->activation;
  const size_t per_tile = ckpt_floats_per_tile(net->n_hidden) * sizeof(float);
  if (saved) {
    const size_t need = (size_t)((n + TILE - 1) / TILE) * per_tile + dwpart_bytes();
    if (ckpt_bytes < need) return fail(TPX_E_WORKSPACE, "checkpoint buffer of %zu bytes is too small (%zu needed)", ckpt_bytes, need);
    a.dwpart = (float*)((char*)ckpt + (need - dwpart_bytes()));
    a.x = x; a.n = n; a.gu = gu; a.gdu = gdu; a.glap = glap;
    return launch_tc_bwd_any(a, jet.nd, lap, sm_count, stream);
  }
  // no saved checkpoints: chunks of (forward kernel writing checkpoints only) + reverse kernel
  const int64_t ctiles = ckpt_bytes > dwpart_bytes() ? (int64_t)((ckpt_bytes - dwpart_bytes()) / per_tile) : 0;
  a.dwpart = (float*)((char*)ckpt + (size_t)ctiles * per_tile);
  if (ctiles < 1) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes holds no tile (%zu bytes each)", ckpt_bytes, per_tile);
  const int64_t cpts = ctiles * TILE;
  const int d_in = net->d_in, d_out = net->d_out, nd = jet.nd;
  for (int64_t p0 = 0; p0 < n; p0 += cpts) {
    const int64_t m = (n - p0 < cpts) ? n - p0 : cpts;
    int rc = tc_forward(net, jet, lap, img, x + p0 * d_in, m, nullptr, nullptr, nullptr, ckpt, sm_count, stream);
    if (rc) return fail(rc, "tc_fwd_kernel (checkpoint pass of the reverse sweep)");
    a.x = x + p0 * d_in; a.n = m;
    a.gu = gu ? gu + p0 * d_out : nullptr;
    a.gdu = gdu ? gdu + p0 * d_out * nd : nullptr;
    a.glap = glap ? glap + p0 * d_out : nullptr;
    rc = launch_tc_bwd_any(a, nd, lap, sm_count, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace tpx
