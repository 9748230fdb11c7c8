// Tensor-core reverse kernel for hidden width <= 64 (round 2): rows = (stream, point) as in the round-1 kernel
// (tc_bwd.cu), operands in the split-fp16 format of tcp_common.cuh, checkpoints of tcp_fwd.cu.
//
// A tile is 32 points; the 128 rows of every UMMA are (stream q, point p) = TMEM lane 32 q + p.  Per hidden matrix l the
// activation warps (16 = 4 neuron groups x 4 quadrants) run one step:
//   in       the owner warp of stream s moves its 16 adjoint columns of D to the group's exchange buffer
//   compute  warp w of the group takes 4 neurons for ALL streams of its 32 points: Zbar_l (adjoint jets with checkpoint_l)
//            and a_{l-1} (jets of checkpoint_{l-1}), scaled by powers of two, split hi / lo (fp16) and stored as 128-byte
//            rows (row = (stream, point), 64 neurons, 128-byte swizzle) into one of two image buffers
//            [Z_hi | Z_lo | A_hi | A_lo] with 8-byte vector stores
// and the MMA warp issues from that ONE set of images
//   adjoint   D[r][i] = sum_j Zbar[r][j] W_l[j][i]        A = the Z images read K-major, B = the image of row-major W_l
//                                                          read MN-major (no transposed weight copy, no TMEM A operand)
//   gradient  dW_l[j][i] += sum_r Zbar[r][j] a[r][i]      K = the 128 rows: A = [Z_hi ; Z_lo] read MN-major (M = 128),
//                                                          B = A_lo, then A_hi (MN-major, N = 64); 64 TMEM columns per matrix
// Against round 1 this drops the exchange-out, the TMEM A operand, the transposed scalar operand stores (16 per element ->
// 0.5) and half of the operand bytes: 296 KB instead of 496 KB through the shared-memory port per tile and matrix.
//
// Scales: tcp_common.cuh.  Z of stream s carries 2^ez from a bound of |Zbar_s| over the tile (||W||_1 x the exact maximum of
// the stream one step above, measured lazily; pushed through the adjoint-jet formulas with the exact maxima of the
// checkpoints recorded by the forward sweep per 128-point tile); A carries 2^(ePi - ez), ePi = the sticky scale of the
// matrix's dW accumulator, which changes only when the accumulator is flushed into per-CTA fp32 partial sums (every
// FLUSH_TILES tiles, or at once when a tile does not fit).
#include <cstdio>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_kernels.h"
#include "tcp_common.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tce {

using namespace tpx::tcp;
using Layout = tpx::tcp::Layout;   // (tpx::Layout is the FP32-kernel image, spelled out where it is meant)

constexpr int TILE = 32;                 // points per tile
constexpr int E_CH = 16;                 // neurons per warp group
constexpr int E_GROUPS = HP / E_CH;
constexpr int E_NW = E_CH / 4;           // neurons per warp in the compute phase
constexpr int E_MAXL = 4;                // hidden layers (all weight images resident next to two image buffers)
constexpr int FLUSH_TILES = 16;          // dW accumulators -> fp32 partial sums every 16 tiles (512 points)
constexpr int MAX_CTAS = 160;
constexpr uint32_t T_D = 0, T_ACC = 64;  // TMEM columns
constexpr int IMG = 16384;               // bytes of one image (128 rows x 128 B)
constexpr int IMGBUF = 4 * IMG;
constexpr int BXP = 20;                  // padded neuron extent of one (stream, point) row of the exchange buffer
constexpr int BX_GROUP = 4 * TILE * BXP;
constexpr int E_WARPS = 20, E_THREADS = 32 * E_WARPS;     // 16 activation warps, the MMA warp, 3 idle warps of its warpgroup
constexpr int ACT_REGS = 104, MMA_REGS = 64;              // 4 x 128 x (104 - 96) taken = 128 x (96 - 64) released
constexpr int BAR_TILE = 5, BAR_ALL = 6;

struct Smem {
  int img, wimg, ex, small, acc, mub, mzb, flags, total;   // byte offsets from the 1 KB aligned base
};
constexpr int ACC_DB = 0, ACC_W0 = E_MAXL * HP, ACC_WL = ACC_W0 + MAXIO * HP, ACC_BL = ACC_WL + MAXIO * HP, ACC_TOTAL = ACC_BL + HP;
__host__ __device__ inline Smem smem_of(const Layout& lay) {
  Smem s;
  s.img = 0;
  s.wimg = 2 * IMGBUF;
  s.ex = s.wimg + (lay.L - 1) * 16384;
  s.small = s.ex + E_GROUPS * BX_GROUP * 4;
  s.acc = s.small + (int)lay.small() * 4;
  s.mub = s.acc + ACC_TOTAL * 4;
  s.mzb = s.mub + 2 * 4 * 4;
  s.flags = s.mzb + 2 * E_MAXL * 4 * 4;
  s.total = s.flags + 16;
  return s;
}

struct Args {
  Layout lay;
  tpx::Layout lay32;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  const float *gu, *gdu, *glap;
  double* gacc;
  const float* ckpt;     // tcp_fwd.cu: [tile128][l = 1..L-1][s][64][128]
  const float* stats;    // [tile128][L][STATS]
  float* dwpart;         // per CTA: (E_MAXL - 1) x [64 i][128 rows (j hi, j lo)]
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
__device__ __forceinline__ void st_shared_v2(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}

template <int ND, int LAP, int ACT>
__global__ void __launch_bounds__(E_THREADS, 1) tce_bwd_kernel(Args a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_z[2], bar_f[2], bar_d;
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Layout lay = a.lay;
  const Smem so = smem_of(lay);
  const int L = lay.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* small = reinterpret_cast<float*>(sm + so.small);
  float* acc = reinterpret_cast<float*>(sm + so.acc);
  uint32_t* mub = reinterpret_cast<uint32_t*>(sm + so.mub);
  uint32_t* mzb = reinterpret_cast<uint32_t*>(sm + so.mzb);
  uint32_t* flags = reinterpret_cast<uint32_t*>(sm + so.flags);

  {
    const int nv = (L - 1) * 1024;
    const float4* src = reinterpret_cast<const float4*>(a.img + lay.mat(1, 0));
    float4* dst = reinterpret_cast<float4*>(sm + so.wimg);
    for (int v = tid; v < nv; v += E_THREADS) dst[v] = __ldg(src + v);
    for (int v = tid; v < (int)lay.small(); v += E_THREADS) small[v] = __ldg(a.img + v);
    for (int v = tid; v < ACC_TOTAL; v += E_THREADS) acc[v] = 0.0f;
    for (int v = tid; v < 2 * 4 + 2 * E_MAXL * 4 + 4; v += E_THREADS) mub[v] = 0u;     // mub, mzb, flags are contiguous
    float4* im = reinterpret_cast<float4*>(sm + so.img);                             // rows of unused streams stay zero
    for (int v = tid; v < 2 * IMGBUF / 16; v += E_THREADS) im[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float* dwp = a.dwpart + (size_t)blockIdx.x * ((E_MAXL - 1) * HP * 128);
  for (int v = tid; v < (L - 1) * HP * 128; v += E_THREADS) dwp[v] = 0.0f;
  if (warp == 16) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int r = 0; r < 2; ++r) { mbar_init(&bar_z[r], 16); mbar_init(&bar_f[r], 1); }
    mbar_init(&bar_d, 1);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  const uint32_t img0 = smem_u32(sm + so.img);

  if (warp >= 16) {
    reg_dec<MMA_REGS>();
    if (warp == 16) {
      // =========================== MMA issuer ===========================
      const uint32_t wbase = smem_u32(sm + so.wimg);
      constexpr uint32_t idesc_adj = idesc_f16(128, HP, 0, 1);
      constexpr uint32_t idesc_dw = idesc_f16(128, HP, 1, 1);
      uint32_t g = 0;                           // global step counter
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = L - 1; l >= 1; --l, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait(&bar_z[b], (g >> 1) & 1u);
          fence_after();
          const uint32_t acc_flag = flags[b];
          if (elect_one()) {
            const uint32_t wh = wbase + (uint32_t)(l - 1) * 16384u, wl = wh + 8192u;
            const uint32_t zh = img0 + b * IMGBUF, zl = zh + IMG, ah = zh + 2 * IMG, al = zh + 3 * IMG;
            const uint32_t d = tmem + T_D, tacc = tmem + T_ACC + (uint32_t)(l - 1) * 64u;
            // the two small correction products first, the main product last (the TMEM accumulator rounds toward zero)
#pragma unroll
            for (int kc = 0; kc < 4; ++kc)
              mma_f16_ss(d, desc_sw128(zl + kc * 32, 16, 1024), desc_sw128(wh + kc * 2048, 16384, 1024), idesc_adj, kc > 0);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc)
              mma_f16_ss(d, desc_sw128(zh + kc * 32, 16, 1024), desc_sw128(wl + kc * 2048, 16384, 1024), idesc_adj, 1);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc)
              mma_f16_ss(d, desc_sw128(zh + kc * 32, 16, 1024), desc_sw128(wh + kc * 2048, 16384, 1024), idesc_adj, 1);
            mma_commit(&bar_d);
#pragma unroll
            for (int kp = 0; kp < 8; ++kp)
              mma_f16_ss(tacc, desc_sw128(zh + kp * 2048, 16384, 1024), desc_sw128(al + kp * 2048, 16384, 1024), idesc_dw,
                         kp > 0 ? 1u : acc_flag);
#pragma unroll
            for (int kp = 0; kp < 8; ++kp)
              mma_f16_ss(tacc, desc_sw128(zh + kp * 2048, 16384, 1024), desc_sw128(ah + kp * 2048, 16384, 1024), idesc_dw, 1);
            mma_commit(&bar_f[b]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    reg_inc<ACT_REGS>();
    // =========================== activation warps ===========================
    const tpx::Layout lay32 = a.lay32;
    const int d_in = lay.d_in, d_out = lay.d_out;
    const int grp = warp >> 2, quad = warp & 3;
    const int jc = grp * E_CH, jn = jc + quad * E_NW;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t tD = tmem + T_D + lane_base;
    float* ex = reinterpret_cast<float*>(sm + so.ex) + grp * BX_GROUP;
    float* ex_own = ex + (quad * TILE + lane) * BXP;           // owner view: stream quad, the group's 16 neurons
    float* ex_cmp = ex + lane * BXP + quad * E_NW;             // compute view of stream s: + s * TILE * BXP
    const float* W0 = small + lay.w0() + jn;
    const float* WL = small + lay.wl() + jn;
    const float* bias0 = small + lay.bias(0) + jn;
    const int barid = 1 + grp;
    const bool owner = quad < S;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    // image store offsets of this thread: row = 32 s + lane, 8 bytes (4 neurons) at chunk (jn / 8) ^ (row & 7)
    const uint32_t store_off = (uint32_t)lane * 128u + (uint32_t)((((jn >> 3) ^ (lane & 7)) << 4) + ((jn & 7) << 1));
    uint32_t g = 0;                      // global step counter (same sequence as the MMA warp's)
    uint32_t ph_d = 0;
    int tcount = 0;
    uint32_t live = 0;
    int epi[E_MAXL], emin_prev[E_MAXL];
#pragma unroll
    for (int l = 0; l < E_MAXL; ++l) { epi[l] = 0; emin_prev[l] = 1 << 20; }

    auto drain_dw = [&]() {              // every gradient MMA issued so far has completed
      if (g >= 1) mbar_wait(&bar_f[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
      fence_after();
    };
    // this thread's slice of accumulator l -> the CTA's fp32 partial sums: TMEM lane = row (j hi | j lo), columns = i
    auto flush_dw = [&](int l, int e) {
      float w[E_CH];
      tmem_ld16(tmem + T_ACC + (uint32_t)(l - 1) * 64u + lane_base + jc, w);
      const float sc = exp2f((float)-e);
      float* dst = dwp + ((size_t)(l - 1) * HP + jc) * 128 + quad * 32 + lane;
#pragma unroll
      for (int i = 0; i < E_CH; ++i) atomicAdd(dst + i * 128, w[i] * sc);
      fence_before();
    };
    // activation jets of a layer from its checkpoints (keep, d_k z, L z) -> (a, d_k a, L a)
    auto jets = [&](const float (&c)[S][E_NW], float (&h)[S][E_NW]) {
#pragma unroll
      for (int i = 0; i < E_NW; ++i) {
        float t, s1, s2;
        act_from_keep<ACT>(c[0][i], t, s1, s2);
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
    };
    // adjoints of the pre-activation streams of a layer from the adjoints of its activation streams
    auto zbar = [&](const float (&c)[S][E_NW], const float (&ga)[S][E_NW], float (&zb)[S][E_NW]) {
#pragma unroll
      for (int i = 0; i < E_NW; ++i) {
        float t, s1, s2;
        act_from_keep<ACT>(c[0][i], t, s1, s2);
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
          const float s3 = (ACT == 0) ? -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f) : -s1;
          v0 = fmaf(gl, fmaf(s3, qq, s2 * c[S - 1][i]), v0);
          zb[S - 1][i] = gl * s1;
        }
        zb[0][i] = v0;
      }
    };

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int set = (int)((tile - blockIdx.x) / gridDim.x) & 1;
      const int64_t gp = tile * TILE + lane;
      const bool valid = gp < a.n;
      float xr[MAXIO];
#pragma unroll
      for (int cc = 0; cc < MAXIO; ++cc) xr[cc] = (valid && cc < d_in) ? __ldg(a.x + gp * d_in + cc) : 0.0f;
      // checkpoints of this tile inside its 128-point forward tile
      const float* ck_tile = a.ckpt + (size_t)(tile >> 2) * ((size_t)(L - 1) * S * HP * TP) + (size_t)jn * TP + (tile & 3) * 32 + lane;
      const float* st_tile = a.stats + (size_t)(tile >> 2) * (L * STATS);
      auto load_ck = [&](int l, float (&c)[S][E_NW]) {
        if (l >= 1) {
          const float* base = ck_tile + (size_t)(l - 1) * S * HP * TP;
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < E_NW; ++i) c[s][i] = __ldcs(base + (s * HP + i) * TP);
        } else {                               // layer 0 is recomputed: z = W0 x + b0, d_k z = W0[:, dcol_k], L z = 0
#pragma unroll
          for (int i = 0; i < E_NW; ++i) {
            float z0 = bias0[i];
#pragma unroll
            for (int cc = 0; cc < MAXIO; ++cc)
              if (cc < d_in) z0 = fmaf(W0[cc * HP + i], xr[cc], z0);
            c[0][i] = (ACT == 0) ? tanh_fast(z0) : z0;
#pragma unroll
            for (int k = 0; k < ND; ++k) c[1 + k][i] = W0[dcol[k] * HP + i];
            if (LAP) c[S - 1][i] = 0.0f;
          }
        }
      };
      // ---- adjoints of this point's output streams, their tile maxima, output bias
      float gq[S][MAXIO];
#pragma unroll
      for (int o = 0; o < MAXIO; ++o) {
        const bool on = valid && o < d_out;
        gq[0][o] = (on && a.gu) ? __ldg(a.gu + gp * d_out + o) : 0.0f;
#pragma unroll
        for (int k = 0; k < ND; ++k) gq[1 + k][o] = (on && a.gdu) ? __ldg(a.gdu + (gp * d_out + o) * ND + k) : 0.0f;
        if (LAP) gq[S - 1][o] = (on && a.glap) ? __ldg(a.glap + gp * d_out + o) : 0.0f;
      }
      if (warp == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float m = 0.0f;
#pragma unroll
          for (int o = 0; o < MAXIO; ++o) m = fmaxf(m, fabsf(gq[s][o]));
          m = warp_max_nonneg(m);
          if (lane == 0) mub[set * 4 + s] = __float_as_uint(m);
        }
#pragma unroll
        for (int o = 0; o < MAXIO; ++o) {
          float sum = gq[0][o];
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
          if (lane == 0 && o < d_out) acc[ACC_BL + o] += sum;
        }
      }
      named_bar_sync(BAR_TILE, 512);
      if (warp == 0 && lane < E_MAXL * 4) mzb[(set ^ 1) * E_MAXL * 4 + lane] = 0u;   // the other set: every warp has left the previous tile
      const bool periodic = (tcount >= FLUSH_TILES);
      if (periodic) tcount = 0;
      ++tcount;

      float ck[S][E_NW], ckn[S][E_NW];       // checkpoints of the current layer and of the one below (prefetched)
      load_ck(L - 1, ck);
      load_ck(L - 2, ckn);
      float ginv[S];
#pragma unroll
      for (int s = 0; s < S; ++s) ginv[s] = 0.0f;
      int step = 0;
      for (int l = L - 1; l >= 1; --l, ++step, ++g) {
        const bool top = (l == L - 1);
        float ga[S][E_NW], zb[S][E_NW];
        if (top) {
          // u_s = WL a_s (+ bL): adjoint and gradient are thread-local
          float al[S][E_NW];
          jets(ck, al);
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < E_NW; ++i) ga[s][i] = 0.0f;
#pragma unroll
          for (int o = 0; o < MAXIO; ++o)
            if (o < d_out) {
              float wg[E_NW];
#pragma unroll
              for (int i = 0; i < E_NW; ++i) wg[i] = 0.0f;
#pragma unroll
              for (int s = 0; s < S; ++s)
#pragma unroll
                for (int i = 0; i < E_NW; ++i) {
                  ga[s][i] = fmaf(WL[o * HP + i], gq[s][o], ga[s][i]);
                  wg[i] = fmaf(gq[s][o], al[s][i], wg[i]);
                }
              const float r = reduce4(wg[0], wg[1], wg[2], wg[3], lane);
              if (lane < E_NW) acc[ACC_WL + o * HP + jn + lane] += r;
            }
        } else {
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < E_NW; ++i) ck[s][i] = ckn[s][i];
          load_ck(l - 1, ckn);                 // behind the wait for the adjoint product
          // ---- in: adjoint product of the previous step, rows of this warp's stream -> exchange buffer -> all streams
          mbar_wait(&bar_d, ph_d);
          ph_d ^= 1;
          fence_after();
          if (owner) {
            float w[E_CH];
            tmem_ld16(tD + jc, w);
#pragma unroll
            for (int v = 0; v < E_CH / 4; ++v)
              *reinterpret_cast<float4*>(ex_own + 4 * v) = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
          }
          fence_before();
          named_bar_sync(barid, 128);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float4 v4 = *reinterpret_cast<const float4*>(ex_cmp + s * TILE * BXP);
            ga[s][0] = v4.x * ginv[s]; ga[s][1] = v4.y * ginv[s]; ga[s][2] = v4.z * ginv[s]; ga[s][3] = v4.w * ginv[s];
          }
        }
        zbar(ck, ga, zb);
        {                                      // bias gradient of layer l
          const float r = reduce4(zb[0][0], zb[0][1], zb[0][2], zb[0][3], lane);
          if (lane < E_NW) acc[ACC_DB + l * HP + jn + lane] += r;
        }
        float ap[S][E_NW];
        jets(ckn, ap);
        // ---- scales: bounds of |Zbar_s| over the tile, exact maxima of |a_s| from the forward sweep
        float zsc[S], asc[S];
        int my_epi = 0, my_emin = 1 << 20;
#pragma unroll
        for (int q = 0; q < E_MAXL; ++q) { if (q == l) { my_epi = epi[q]; my_emin = emin_prev[q]; } }
        bool my_live = (live >> l) & 1u;
        {
          float Mz[S], Ba[S], Bz[S];
          Mz[0] = 0.0f;
          const float wnorm = top ? (float)d_out * small[lay.wlmax()] : small[lay.norm(l + 1) + 2];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (s >= 1) Mz[s] = __ldg(st_tile + l * STATS + 4 + (s - 1));
            Ba[s] = wnorm * __uint_as_float(top ? mub[set * 4 + s] : mzb[(set * E_MAXL + (step - 1)) * 4 + s]);
          }
          Bz[0] = Ba[0];
          float q = 0.0f;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            Bz[0] = fmaf(s2_max<ACT>() * Mz[1 + k], Ba[1 + k], Bz[0]);
            q = fmaf(fabsf(lapw[k]) * Mz[1 + k], Mz[1 + k], q);
            Bz[1 + k] = Ba[1 + k];
            if (LAP) Bz[1 + k] = fmaf(2.0f * fabsf(lapw[k]) * s2_max<ACT>() * Mz[1 + k], Ba[S - 1], Bz[1 + k]);
          }
          if (LAP) {
            Bz[0] = fmaf(Ba[S - 1], fmaf(s3_max<ACT>(), q, s2_max<ACT>() * Mz[S - 1]), Bz[0]);
            Bz[S - 1] = Ba[S - 1];
          }
          int ezb[S], esum_min = 1 << 20;
          const float itau = small[lay.norm(l) + 0];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            ezb[s] = scale_exp_for(Bz[s]);
            const float ma = (s == 0) ? 1.0f : __ldg(st_tile + (l - 1) * STATS + s);
            const int esum = (ezb[s] - 127) + (scale_exp_for(ma) - 127);
            esum_min = esum < esum_min ? esum : esum_min;
          }
          bool need = my_live && (esum_min < my_epi || periodic);
          if (need) {
            drain_dw();
            flush_dw(l, my_epi);
            my_live = false;
          }
          if (!my_live) my_epi = (esum_min < my_emin ? esum_min : my_emin) - 4;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            zsc[s] = exp_to_float(ezb[s]);
            asc[s] = exp_to_float(my_epi - (ezb[s] - 127) + 127);
            ginv[s] = inv_of_exp(ezb[s]) * itau;
          }
#pragma unroll
          for (int q2 = 0; q2 < E_MAXL; ++q2) { if (q2 == l) { epi[q2] = my_epi; emin_prev[q2] = esum_min; } }
        }
        // ---- images of this step: the MMAs of two steps ago have finished reading the buffer
        const uint32_t b = g & 1u;
        if (g >= 2) mbar_wait(&bar_f[b], ((g >> 1) - 1) & 1u);
        const uint32_t base = img0 + b * IMGBUF + store_off;
        float mloc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          uint32_t h0, l0, h1, l1;
          split2(zb[s][0] * zsc[s], zb[s][1] * zsc[s], h0, l0);
          split2(zb[s][2] * zsc[s], zb[s][3] * zsc[s], h1, l1);
          st_shared_v2(base + s * 4096, h0, h1);
          st_shared_v2(base + s * 4096 + IMG, l0, l1);
          split2(ap[s][0] * asc[s], ap[s][1] * asc[s], h0, l0);
          split2(ap[s][2] * asc[s], ap[s][3] * asc[s], h1, l1);
          st_shared_v2(base + s * 4096 + 2 * IMG, h0, h1);
          st_shared_v2(base + s * 4096 + 3 * IMG, l0, l1);
          mloc[s] = fmaxf(fmaxf(fabsf(zb[s][0]), fabsf(zb[s][1])), fmaxf(fabsf(zb[s][2]), fabsf(zb[s][3])));
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float m = warp_max_nonneg(mloc[s]);
          if (lane == 0) atomicMax(mzb + (set * E_MAXL + step) * 4 + s, __float_as_uint(m));
        }
        if (warp == 0 && lane == 0) flags[b] = my_live ? 1u : 0u;
        live |= (1u << l);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_z[b]);
      }
      // ------------------------------------------------ first layer: z = W0 x + b0, d_k z = W0[:, dcol_k], L z = 0
      {
        float ga[S][E_NW], zb[S][E_NW];
        mbar_wait(&bar_d, ph_d);
        ph_d ^= 1;
        fence_after();
        if (owner) {
          float w[E_CH];
          tmem_ld16(tD + jc, w);
#pragma unroll
          for (int v = 0; v < E_CH / 4; ++v)
            *reinterpret_cast<float4*>(ex_own + 4 * v) = make_float4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
        }
        fence_before();
        named_bar_sync(barid, 128);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float4 v4 = *reinterpret_cast<const float4*>(ex_cmp + s * TILE * BXP);
          ga[s][0] = v4.x * ginv[s]; ga[s][1] = v4.y * ginv[s]; ga[s][2] = v4.z * ginv[s]; ga[s][3] = v4.w * ginv[s];
        }
        zbar(ckn, ga, zb);
        {
          const float r = reduce4(zb[0][0], zb[0][1], zb[0][2], zb[0][3], lane);
          if (lane < E_NW) acc[ACC_DB + 0 * HP + jn + lane] += r;
        }
        float rk[NDX];
#pragma unroll
        for (int k = 0; k < ND; ++k) rk[k] = reduce4(zb[1 + k][0], zb[1 + k][1], zb[1 + k][2], zb[1 + k][3], lane);
#pragma unroll
        for (int cc = 0; cc < MAXIO; ++cc)
          if (cc < d_in) {
            const float xc = xr[cc];
            float r = reduce4(zb[0][0] * xc, zb[0][1] * xc, zb[0][2] * xc, zb[0][3] * xc, lane);
#pragma unroll
            for (int k = 0; k < ND; ++k)
              if (dcol[k] == cc) r += rk[k];
            if (lane < E_NW) acc[ACC_W0 + cc * HP + jn + lane] += r;
          }
      }
    }

    // ------------------------------------------------------------------ flush everything
    drain_dw();
    for (int l = 1; l < L; ++l)
      if ((live >> l) & 1u) {
        int e = 0;
#pragma unroll
        for (int q = 0; q < E_MAXL; ++q) { if (q == l) e = epi[q]; }
        flush_dw(l, e);
      }
    __threadfence();
    named_bar_sync(BAR_ALL, 512);
    const int HP32 = lay32.HP;
    if (lane < E_NW) {
      const int j = jn + lane;
      if (j < HP32) {
#pragma unroll
        for (int l = 0; l < E_MAXL; ++l)
          if (l < L) atomicAdd(a.gacc + (l == 0 ? lay32.b0() : lay32.b(l)) + j, (double)acc[ACC_DB + l * HP + j]);
#pragma unroll
        for (int cc = 0; cc < MAXIO; ++cc)
          if (cc < d_in) atomicAdd(a.gacc + lay32.w0() + (size_t)cc * HP32 + j, (double)acc[ACC_W0 + cc * HP + j]);
#pragma unroll
        for (int o = 0; o < MAXIO; ++o)
          if (o < d_out) atomicAdd(a.gacc + lay32.wl() + (size_t)o * HP32 + j, (double)acc[ACC_WL + o * HP + j]);
      }
    }
    if (tid < d_out) atomicAdd(a.gacc + lay32.bl() + tid, (double)acc[ACC_BL + tid]);
    if (ntiles > blockIdx.x) {
      for (int l = 1; l < L; ++l)
        for (int e = tid; e < HP * HP; e += 512) {
          const int i = e / HP, j = e % HP;
          const float* src = dwp + ((size_t)(l - 1) * HP + i) * 128;
          const double v = (double)__ldcg(src + j) + (double)__ldcg(src + 64 + j);
          if (j < HP32 && i < HP32) atomicAdd(a.gacc + lay32.w(l) + (size_t)j * HP32 + i, v);
        }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace tce

using namespace tce;

bool tce_bwd_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tcp_supported(net, nd, lap)) return false;
  if (net->n_hidden > E_MAXL) return false;
  const tcp::Layout lay{net->n_hidden, net->d_in, net->d_out};
  if ((size_t)smem_of(lay).total + 1024 > 227 * 1024) return false;
  return true;
}

static size_t tce_dwpart_bytes() { return (size_t)MAX_CTAS * (E_MAXL - 1) * HP * 128 * sizeof(float); }
static size_t tce_tile_bytes(int L, int S) { return (tcp_ckpt_floats_per_tile(L, S) + (size_t)L * STATS) * sizeof(float); }

// saved buffer: [checkpoints of all 128-point tiles][stats of all tiles][dW partial sums]
size_t tce_saved_bytes(const tpx_fcn_desc* net, int nd, int lap, int64_t n) {
  const int64_t ntiles = (n + TP - 1) / TP;
  return (size_t)ntiles * tce_tile_bytes(net->n_hidden, 1 + nd + lap) + tce_dwpart_bytes();
}
int64_t tce_tiles_in(const tpx_fcn_desc* net, int nd, int lap, size_t bytes) {
  if (bytes <= tce_dwpart_bytes()) return 0;
  return (int64_t)((bytes - tce_dwpart_bytes()) / tce_tile_bytes(net->n_hidden, 1 + nd + lap));
}
void tce_saved_split(const tpx_fcn_desc* net, int nd, int lap, int64_t ntiles, float* base, float** ckpt, float** stats, float** dwpart) {
  const int S = 1 + nd + lap;
  *ckpt = base;
  *stats = base + (size_t)ntiles * tcp_ckpt_floats_per_tile(net->n_hidden, S);
  *dwpart = *stats + (size_t)ntiles * net->n_hidden * STATS;
}

template <int ND, int LAP, int ACT>
static int launch_act(const Args& a, int sm_count, cudaStream_t stream) {
  const size_t smem = (size_t)smem_of(a.lay).total + 1024;
  if (smem > 227 * 1024) return fail(TPX_E_UNSUPPORTED, "tce_bwd_kernel needs %zu bytes of shared memory", smem);
  const int64_t ntiles = (a.n + TILE - 1) / TILE;
  int64_t grid = sm_count < ntiles ? sm_count : ntiles;
  if (grid > MAX_CTAS) grid = MAX_CTAS;
  if (grid < 1) return 0;
  cudaError_t e = cudaFuncSetAttribute(tce_bwd_kernel<ND, LAP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tce_bwd_kernel: shared memory attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
  tce_bwd_kernel<ND, LAP, ACT><<<(unsigned)grid, E_THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tce_bwd_kernel launch (%lld CTAs, %zu bytes smem): %s", (long long)grid, smem, cudaGetErrorString(e));
  return 0;
}
template <int ND, int LAP>
static int launch_nd(const Args& a, int act, int sm_count, cudaStream_t stream) {
  return act == TPX_ACT_SIN ? launch_act<ND, LAP, 1>(a, sm_count, stream) : launch_act<ND, LAP, 0>(a, sm_count, stream);
}
static int launch_any(const Args& a, int act, int nd, int lap, int sm_count, cudaStream_t stream) {
  switch (nd * 2 + lap) {
    case 0: return launch_nd<0, 0>(a, act, sm_count, stream);
    case 2: return launch_nd<1, 0>(a, act, sm_count, stream);
    case 3: return launch_nd<1, 1>(a, act, sm_count, stream);
    case 4: return launch_nd<2, 0>(a, act, sm_count, stream);
    case 5: return launch_nd<2, 1>(a, act, sm_count, stream);
    case 6: return launch_nd<3, 0>(a, act, sm_count, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

int tce_backward(const tpx_fcn_desc* net, const tpx::Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                 int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* buf, size_t buf_bytes,
                 int saved, int sm_count, cudaStream_t stream) {
  Args a;
  a.lay = tcp::Layout{net->n_hidden, net->d_in, net->d_out};
  a.lay32 = lay32;
  a.jet = jet;
  a.img = img; a.gacc = gacc;
  const int nd = jet.nd;
  if (saved) {
    const size_t need = tce_saved_bytes(net, nd, lap, n);
    if (buf_bytes < need) return fail(TPX_E_WORKSPACE, "checkpoint buffer of %zu bytes is too small (%zu needed)", buf_bytes, need);
    float *ck, *st, *dw;
    tce_saved_split(net, nd, lap, (n + TP - 1) / TP, buf, &ck, &st, &dw);
    a.ckpt = ck; a.stats = st; a.dwpart = dw;
    a.x = x; a.n = n; a.gu = gu; a.gdu = gdu; a.glap = glap;
    return launch_any(a, net->activation, nd, lap, sm_count, stream);
  }
  // no saved checkpoints: chunks of (forward kernel writing checkpoints only) + reverse kernel through the workspace
  const int64_t ctiles = tce_tiles_in(net, nd, lap, buf_bytes);
  if (ctiles < 1) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes holds no tile", buf_bytes);
  const int64_t cpts = ctiles * TP;
  const int d_in = net->d_in, d_out = net->d_out;
  float *ck, *st, *dw;
  tce_saved_split(net, nd, lap, ctiles, buf, &ck, &st, &dw);
  a.ckpt = ck; a.stats = st; a.dwpart = dw;
  for (int64_t p0 = 0; p0 < n; p0 += cpts) {
    const int64_t m = (n - p0 < cpts) ? n - p0 : cpts;
    int rc = tcp_forward(net, jet, lap, img, x + p0 * d_in, m, nullptr, nullptr, nullptr, ck, st, sm_count, stream);
    if (rc) return fail(rc, "tcp_fwd_kernel (checkpoint pass of the reverse sweep)");
    a.x = x + p0 * d_in; a.n = m;
    a.gu = gu ? gu + p0 * d_out : nullptr;
    a.gdu = gdu ? gdu + p0 * d_out * nd : nullptr;
    a.glap = glap ? glap + p0 * d_out : nullptr;
    rc = launch_any(a, net->activation, nd, lap, sm_count, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace tpx
