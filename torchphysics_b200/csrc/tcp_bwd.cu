// Points-as-lanes tensor-core reverse kernel (hidden width <= 64, tanh / sin, <= 4 streams); see tcp_common.cuh.
//
// One persistent CTA per SM, 16 activation warps (warp = chunk * 4 + quadrant: TMEM lanes 32 q .. 32 q + 31 = points,
// neurons 16 c .. 16 c + 15) + 1 MMA warp, one 128-point tile in flight.  It sweeps the layers of the tile backwards,
// reading the checkpoints of tcp_fwd.cu ([layer][stream][64 neurons][128 points]: tanh z | z, d_k z, L z).
//
// Reverse step of hidden matrix l (L-1 >= l >= 1), one PASS per stream s (order: Laplacian, d_k ..., value):
//   activation warps   g_s = adjoint of a_{l,s} (TMEM, product of the previous step; top step: WL^T ubar, thread-local)
//                      Zbar_{l,s} = adjoint jets(g, checkpoint_l),   a_{l-1,s} = jets(checkpoint_{l-1})
//                      both scaled by powers of two, split hi / lo (fp16) and stored as 128-byte rows (row = point,
//                      64 neurons, 128-byte swizzle) into one of two 64 KB staging regions  [Z_hi | Z_lo | A_hi | A_lo]
//   MMA warp           adjoint   D_s[p][i] = sum_j Zbar_s[p][j] W_l[j][i]      A = the Z image (K-major), B = the image of
//                                row-major W_l read MN-major -> TMEM block s of X, consumed by the next step
//                      gradient  dW_l[j][i] += sum_p Zbar_s[p][j] a_s[p][i]    K = 128 points: A = [Z_hi ; Z_lo] read MN-major
//                                (M = 128), B = A_hi, then A_lo (MN-major, N = 64) -> 64 TMEM columns per hidden matrix
// The same Z image serves both products, and W needs no transposed copy.  A pass of stream s only waits for the adjoint
// blocks it reads, so the tensor core (70 % busy) runs entirely behind the activation warps; the stores of pass n + 2
// reuse the region of pass n.
//
// Scales (see tcp_common.cuh): the Z image of stream s carries 2^ez with ez from a bound of |Zbar_s| over the tile
// (||W||_1 x the exact tile maximum of the stream one step above, measured lazily, pushed through the adjoint-jet formulas
// with the exact tile maxima of the checkpoints that the forward sweep recorded); the A image carries 2^(ePi - ez) where
// ePi is the scale of the dW accumulator of the matrix, which is sticky: it only changes when the accumulator is flushed
// (every FLUSH_TILES tiles into per-CTA fp32 partial sums, or at once when a tile does not fit).
#include <cstdio>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_kernels.h"
#include "tcp_common.cuh"
#include "tpx_internal.h"

namespace tpx {
namespace tcp {

constexpr int B_MAXL = 5;                 // hidden layers: (L - 1) x 64 accumulator columns behind 4 x 64 adjoint columns
constexpr int FLUSH_TILES = 4;            // dW accumulators -> fp32 partial sums every 4 tiles (512 points)
constexpr int MAX_BWD_CTAS = 160;
constexpr uint32_t T_ACC = 256;           // TMEM column of the first dW accumulator
constexpr int REGION = 65536;             // bytes of one staging region
constexpr int IMG = 16384;                // bytes of one image (128 rows x 128 B)
constexpr int BAR_TILE = 5, BAR_ALL = 6;

struct BwdSmem {
  int wimg, stage, small, acc, mub, mzb, flags, total;   // byte offsets from the 1 KB aligned base
};
// shared accumulators per quadrant: [db: B_MAXL x 64][dW0: MAXIO x 64][dWL: MAXIO x 64]; then [dbL: 4 quads x MAXIO]
constexpr int ACC_DB = 0, ACC_W0 = B_MAXL * HP, ACC_WL = ACC_W0 + MAXIO * HP, ACC_Q = ACC_WL + MAXIO * HP;
constexpr int ACC_BL = 4 * ACC_Q, ACC_TOTAL = ACC_BL + 4 * MAXIO;
__host__ __device__ inline BwdSmem bwd_smem(const Layout& lay) {
  BwdSmem s;
  s.stage = 0;
  s.wimg = 2 * REGION;
  s.small = s.wimg + (lay.L - 1) * 16384;
  s.acc = s.small + (int)lay.small() * 4;
  s.mub = s.acc + ACC_TOTAL * 4;          // [2 sets][4] tile maxima of |ubar_s|
  s.mzb = s.mub + 2 * 4 * 4;              // [2 sets][B_MAXL steps][4] tile maxima of |Zbar_s|
  s.flags = s.mzb + 2 * B_MAXL * 4 * 4;   // [2 regions] accumulate flag of the pass's first gradient MMA
  s.total = s.flags + 16;
  return s;
}

struct BwdArgs {
  Layout lay;
  tpx::Layout lay32;
  JetArgs jet;
  const float* img;
  const float* x;
  int64_t n;
  const float *gu, *gdu, *glap;
  double* gacc;
  const float* ckpt;     // [tile][l = 1..L-1][s][64][128]
  const float* stats;    // [tile][L][STATS]
  float* dwpart;         // per CTA: (B_MAXL - 1) x [64 i][128 rows (j hi, j lo)] fp32 partial sums
};

// lane receives the sum over all 32 lanes of v[lane & 7]
__device__ __forceinline__ float reduce8(const float (&v)[8], int lane) {
  const bool u4 = (lane & 4) != 0;
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = (u4 ? v[i + 4] : v[i]) + __shfl_xor_sync(0xffffffffu, u4 ? v[i] : v[i + 4], 4);
  const bool u2 = (lane & 2) != 0;
  float b[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) b[i] = (u2 ? a[i + 2] : a[i]) + __shfl_xor_sync(0xffffffffu, u2 ? a[i] : a[i + 2], 2);
  const bool u1 = (lane & 1) != 0;
  float c = (u1 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, u1 ? b[0] : b[1], 1);
  c += __shfl_xor_sync(0xffffffffu, c, 8);
  c += __shfl_xor_sync(0xffffffffu, c, 16);
  return c;
}

__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int ND, int LAP, int ACT>
__global__ void __maxnreg__(120) tcp_bwd_kernel(BwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_z[2], bar_f[2], bar_d[4];
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Layout lay = a.lay;
  const tpx::Layout lay32 = a.lay32;
  const BwdSmem so = bwd_smem(lay);
  const int L = lay.L, d_in = lay.d_in, d_out = lay.d_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* small = reinterpret_cast<float*>(sm + so.small);
  float* accs = reinterpret_cast<float*>(sm + so.acc);
  uint32_t* mub = reinterpret_cast<uint32_t*>(sm + so.mub);
  uint32_t* mzb = reinterpret_cast<uint32_t*>(sm + so.mzb);
  uint32_t* flags = reinterpret_cast<uint32_t*>(sm + so.flags);

  {
    const int nv = (L - 1) * 1024;
    const float4* src = reinterpret_cast<const float4*>(a.img + lay.mat(1, 0));
    float4* dst = reinterpret_cast<float4*>(sm + so.wimg);
    for (int v = tid; v < nv; v += THREADS) dst[v] = __ldg(src + v);
    for (int v = tid; v < (int)lay.small(); v += THREADS) small[v] = __ldg(a.img + v);
    for (int v = tid; v < ACC_TOTAL; v += THREADS) accs[v] = 0.0f;
    for (int v = tid; v < 2 * 4 + 2 * B_MAXL * 4 + 4; v += THREADS) mub[v] = 0u;     // mub, mzb, flags are contiguous
  }
  float* dwp = a.dwpart + (size_t)blockIdx.x * ((B_MAXL - 1) * HP * 128);
  for (int v = tid; v < (L - 1) * HP * 128; v += THREADS) dwp[v] = 0.0f;
  if (warp == ACT_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int r = 0; r < 2; ++r) { mbar_init(&bar_z[r], ACT_WARPS); mbar_init(&bar_f[r], 1); }
    for (int s = 0; s < 4; ++s) mbar_init(&bar_d[s], 1);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TP - 1) / TP;
  const uint32_t stage0 = smem_u32(sm + so.stage);

  if (warp == ACT_WARPS) {
    // =========================== MMA issuer ===========================
    const uint32_t wbase = smem_u32(sm + so.wimg);
    constexpr uint32_t idesc_adj = idesc_f16(128, HP, 0, 1);
    constexpr uint32_t idesc_dw = idesc_f16(128, HP, 1, 1);
    uint32_t gp = 0;                          // global pass counter
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int l = L - 1; l >= 1; --l) {
        const uint32_t wh = wbase + (uint32_t)(l - 1) * 16384u, wl = wh + 8192u;
        const uint32_t tacc = tmem + T_ACC + (uint32_t)(l - 1) * 64u;
#pragma unroll
        for (int p = 0; p < S; ++p, ++gp) {
          const int s = S - 1 - p;
          const uint32_t r = gp & 1u;
          mbar_wait(&bar_z[r], (gp >> 1) & 1u);
          fence_after();
          const uint32_t acc_flag = flags[r];
          if (elect_one()) {
            const uint32_t zh = stage0 + r * REGION, zl = zh + IMG, ah = zh + 2 * IMG, al = zh + 3 * IMG;
            const uint32_t d = tmem + (uint32_t)s * 64u;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              const uint64_t bh = desc_sw128(wh + kc * 2048, 16384, 1024), bl = desc_sw128(wl + kc * 2048, 16384, 1024);
              const uint64_t azh = desc_sw128(zh + kc * 32, 16, 1024), azl = desc_sw128(zl + kc * 32, 16, 1024);
              mma_f16_ss(d, azl, bh, idesc_adj, kc > 0);
              mma_f16_ss(d, azh, bl, idesc_adj, 1);
              mma_f16_ss(d, azh, bh, idesc_adj, 1);
            }
            mma_commit(&bar_d[s]);
#pragma unroll
            for (int kp = 0; kp < 8; ++kp) {
              const uint64_t za = desc_sw128(zh + kp * 2048, 16384, 1024);
              mma_f16_ss(tacc, za, desc_sw128(al + kp * 2048, 16384, 1024), idesc_dw, (kp > 0) ? 1u : acc_flag);
              mma_f16_ss(tacc, za, desc_sw128(ah + kp * 2048, 16384, 1024), idesc_dw, 1);
            }
            mma_commit(&bar_f[r]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== activation warps ===========================
    const int quad = warp & 3, chunk = warp >> 2;
    const int j0 = chunk * CHN;
    const int prow = quad * 32 + lane;                       // row of this thread's point in the images
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float* W0 = small + lay.w0();
    const float* WL = small + lay.wl();
    float* acc = accs + quad * ACC_Q;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    uint32_t gp = 0;                          // global pass counter (same sequence as the MMA warp's)
    uint32_t ph_d = 0;                        // bit s: parity of the next completion of bar_d[s] to wait for
    int tcount = 0;                           // tiles accumulated since the last periodic flush
    // per hidden matrix l (bit / byte l): the accumulator holds products; real exponent of its scale; smallest exponent
    // sum a stream of the previous tile allowed (all threads keep identical copies: every input of these is CTA-uniform)
    uint32_t live = 0;
    int epi[B_MAXL], emin_prev[B_MAXL];
#pragma unroll
    for (int l = 0; l < B_MAXL; ++l) { epi[l] = 0; emin_prev[l] = 1 << 20; }

    // all gradient MMAs issued so far have completed (the commit of a pass covers every earlier MMA)
    auto drain_dw = [&]() {
      if (gp >= 1) mbar_wait(&bar_f[(gp - 1) & 1u], (((gp - 1) >> 1)) & 1u);
      fence_after();
    };
    // this thread's 16 columns x 1 lane of accumulator l -> the CTA's fp32 partial sums
    auto flush_dw = [&](int l, int e) {
      const uint32_t tW = tmem + T_ACC + (uint32_t)(l - 1) * 64u + lane_base + j0;
      float w[16];
      tmem_ld16(tW, w);
      const float sc = exp2f((float)-e);
      float* dst = dwp + ((size_t)(l - 1) * HP + j0) * 128 + prow;
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(dst + i * 128, w[i] * sc);
      fence_before();
    };

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int set = (int)((tile - blockIdx.x) / gridDim.x) & 1;
      const int64_t gpt = tile * TP + prow;
      const bool valid = gpt < a.n;
      float xr[MAXIO];
#pragma unroll
      for (int cc = 0; cc < MAXIO; ++cc) xr[cc] = (valid && cc < d_in) ? __ldg(a.x + gpt * d_in + cc) : 0.0f;
      // adjoint of output stream s, column o, of this thread's point
      auto load_ub = [&](int s, int o) -> float {
        if (!(valid && o < d_out)) return 0.0f;
        if (s == 0) return a.gu ? __ldg(a.gu + gpt * d_out + o) : 0.0f;
        if (LAP && s == S - 1) return a.glap ? __ldg(a.glap + gpt * d_out + o) : 0.0f;
        return a.gdu ? __ldg(a.gdu + (gpt * d_out + o) * ND + (s - 1)) : 0.0f;
      };
      if (chunk == 0) {                        // tile maxima of |ubar_s|; output bias gradient
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float m = 0.0f;
#pragma unroll
          for (int o = 0; o < MAXIO; ++o) {
            const float v = load_ub(s, o);
            m = fmaxf(m, fabsf(v));
            if (s == 0) {
              float sum = v;
#pragma unroll
              for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
              if (lane == 0 && o < d_out) accs[ACC_BL + quad * MAXIO + o] += sum;
            }
          }
          m = warp_max_nonneg(m);
          if (lane == 0) atomicMax(mub + set * 4 + s, __float_as_uint(m));
        }
      }
      named_bar_sync(BAR_TILE, 32 * ACT_WARPS);
      if (warp == 0) {                         // clear the other set for the next tile (every warp has left the previous tile)
        if (lane < 4) mub[(set ^ 1) * 4 + lane] = 0u;
        if (lane < B_MAXL * 4) mzb[(set ^ 1) * B_MAXL * 4 + lane] = 0u;
      }
      const bool periodic = (tcount >= FLUSH_TILES);
      if (periodic) tcount = 0;
      ++tcount;

      const float* ck_tile = a.ckpt + (size_t)tile * ((size_t)(L - 1) * S * HP * TP) + (size_t)j0 * TP + prow;
      const float* st_tile = a.stats + (size_t)tile * (L * STATS);
      float ginv[S];                           // unscale factors of the adjoint blocks in TMEM
#pragma unroll
      for (int s = 0; s < S; ++s) ginv[s] = 0.0f;
      int step = 0;

      for (int l = L - 1; l >= 1; --l, ++step) {
        const bool top = (l == L - 1);
        const float* cku = ck_tile + (size_t)(l - 1) * S * HP * TP;                       // checkpoints of layer l
        const float* ckd = (l >= 2) ? ck_tile + (size_t)(l - 2) * S * HP * TP : nullptr;  // of layer l - 1 (layer 0: recomputed)
        const float* bias0 = small + lay.bias(0) + j0;
        // exact tile maxima from the forward sweep
        float Mz[S], Ma[S];
        Ma[0] = 1.0f; Mz[0] = 0.0f;
#pragma unroll
        for (int s = 1; s < S; ++s) {
          Mz[s] = __ldg(st_tile + l * STATS + 4 + (s - 1));
          Ma[s] = __ldg(st_tile + (l - 1) * STATS + s);
        }
        const float wnorm = top ? (float)d_out * small[lay.wlmax()] : small[lay.norm(l + 1) + 2];
        const float itau = small[lay.norm(l) + 0];
        float Ba[S];                           // bounds of |g_s| over the tile (filled as the passes wait for their blocks)
        float ginv_next[S];
        int emin_cur = 1 << 20;
        int my_epi = 0, my_emin = 1 << 20;     // this matrix's entries of epi / emin_prev (dynamic index -> select chain)
#pragma unroll
        for (int q = 0; q < B_MAXL; ++q) { if (q == l) { my_epi = epi[q]; my_emin = emin_prev[q]; } }
        bool my_live = (live >> l) & 1u;
        // held across the passes of this step (the adjoint blocks are overwritten pass by pass): the Laplacian block and
        // everything the value stream's Zbar needs from the other streams
        float gLh[CHN], rv[CHN];
#pragma unroll
        for (int i = 0; i < CHN; ++i) { gLh[i] = 0.0f; rv[i] = 0.0f; }

#pragma unroll
        for (int p = 0; p < S; ++p, ++gp) {
          const int s = S - 1 - p;
          const bool is_lap = LAP && s == S - 1, is_val = (s == 0);
          // ---- the adjoint block of this stream
          if (!top) {
            mbar_wait(&bar_d[s], (ph_d >> s) & 1u);
            ph_d ^= (1u << s);
            fence_after();
            Ba[s] = wnorm * __uint_as_float(mzb[(set * B_MAXL + (step - 1)) * 4 + s]);
          } else {
            Ba[s] = wnorm * __uint_as_float(mub[set * 4 + s]);
          }
          // ---- bound of |Zbar_s|
          float Bz;
          if (is_lap) {
            Bz = Ba[s];
          } else if (!is_val) {
            Bz = Ba[s];
            if (LAP) Bz = fmaf(2.0f * fabsf(lapw[s - 1]) * s2_max<ACT>() * Mz[s], Ba[S - 1], Bz);
          } else {
            Bz = Ba[0];
            float q = 0.0f;
#pragma unroll
            for (int k = 0; k < ND; ++k) {
              Bz = fmaf(s2_max<ACT>() * Mz[1 + k], Ba[1 + k], Bz);
              q = fmaf(fabsf(lapw[k]) * Mz[1 + k], Mz[1 + k], q);
            }
            if (LAP) Bz = fmaf(Ba[S - 1], fmaf(s3_max<ACT>(), q, s2_max<ACT>() * Mz[S - 1]), Bz);
          }
          const int ezb = scale_exp_for(Bz), eab = scale_exp_for(Ma[s]);     // biased exponents of the largest safe scales
          const int esum = (ezb - 127) + (eab - 127);
          emin_cur = esum < emin_cur ? esum : emin_cur;
          // ---- accumulator scale: sticky; flush when this tile does not fit, is far below it, or periodically
          {
            bool need = my_live && (esum < my_epi || esum > my_epi + 24);
            if (p == 0 && periodic && my_live) need = true;
            if (need) {
              drain_dw();
              flush_dw(l, my_epi);
              my_live = false;
            }
            if (!my_live) my_epi = (esum < my_emin ? esum : my_emin) - 4;
          }
          const int ea = my_epi - (ezb - 127);                                // real exponent of the A-image scale (<= eab - 127)
          const float zsc = exp_to_float(ezb), asc = exp_to_float(ea + 127);
          ginv_next[s] = inv_of_exp(ezb) * itau;
          // ---- staging region of this pass: the MMAs of pass gp - 2 have finished reading it
          const uint32_t r = gp & 1u;
          if (gp >= 2) mbar_wait(&bar_f[r], ((gp >> 1) - 1) & 1u);
          const uint32_t reg = stage0 + r * REGION + (uint32_t)prow * 128u;
          float ubs[MAXIO];                    // top step: adjoints of this stream's outputs
#pragma unroll
          for (int o = 0; o < MAXIO; ++o) ubs[o] = top ? load_ub(s, o) : 0.0f;
          float mz_local = 0.0f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int jj = 8 * h;              // neurons j0 + jj .. + 7
            // -- adjoint of a_{l,s}
            float g[8];
            if (!top) {
              uint32_t raw[8];
              tmem_ld8_nowait(tmem + lane_base + s * 64 + j0 + jj, raw);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = __uint_as_float(raw[i]) * ginv[s];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v = 0.0f;
#pragma unroll
                for (int o = 0; o < MAXIO; ++o)
                  if (o < d_out) v = fmaf(WL[o * HP + j0 + jj + i], ubs[o], v);
                g[i] = v;
              }
            }
            // -- Zbar_s of 8 neurons
            float zb[8];
            float aup[8];                      // top step only: a_{L-1,s} for dWL
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float keep = __ldcs(cku + (0 * HP + jj + i) * TP);
              float t, s1, s2;
              act_from_keep<ACT>(keep, t, s1, s2);
              if (is_lap) {
                float q = 0.0f;
#pragma unroll
                for (int k = 0; k < ND; ++k) { const float dz = __ldcs(cku + ((1 + k) * HP + jj + i) * TP); q = fmaf(lapw[k] * dz, dz, q); }
                const float lz = __ldcs(cku + ((S - 1) * HP + jj + i) * TP);
                const float s3 = (ACT == 0) ? -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f) : -s1;
                zb[i] = g[i] * s1;
                gLh[jj + i] = g[i];
                rv[jj + i] = g[i] * fmaf(s3, q, s2 * lz);
                if (top) aup[i] = fmaf(s2, q, s1 * lz);
              } else if (!is_val) {
                const float dz = __ldcs(cku + (s * HP + jj + i) * TP);
                float v = g[i] * s1;
                if (LAP) v = fmaf(2.0f * lapw[s - 1] * gLh[jj + i] * s2, dz, v);
                zb[i] = v;
                rv[jj + i] = fmaf(g[i] * s2, dz, rv[jj + i]);
                if (top) aup[i] = s1 * dz;
              } else {
                zb[i] = fmaf(g[i], s1, rv[jj + i]);
                if (top) aup[i] = t;
              }
              mz_local = fmaxf(mz_local, fabsf(zb[i]));
            }
            if (is_val) {                      // bias gradient of layer l
              const float rs = reduce8(zb, lane);
              if (lane < 8) acc[ACC_DB + l * HP + j0 + jj + lane] += rs;
            }
            if (top) {                         // output-layer weights: dWL[o][j] += sum_p ubar_s[p][o] a_{L-1,s}[p][j]
#pragma unroll
              for (int o = 0; o < MAXIO; ++o)
                if (o < d_out) {
                  float w8[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) w8[i] = ubs[o] * aup[i];
                  const float rs = reduce8(w8, lane);
                  if (lane < 8) acc[ACC_WL + o * HP + j0 + jj + lane] += rs;
                }
            }
            {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) split2(zb[2 * v] * zsc, zb[2 * v + 1] * zsc, hi[v], lo[v]);
              const uint32_t off = (uint32_t)(((2 * chunk + h) ^ (prow & 7)) << 4);
              st_shared_v4(reg + off, hi[0], hi[1], hi[2], hi[3]);
              st_shared_v4(reg + IMG + off, lo[0], lo[1], lo[2], lo[3]);
            }
            // -- a_{l-1,s} of 8 neurons
            float av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float keep;
              if (ckd) {
                keep = __ldcs(ckd + (0 * HP + jj + i) * TP);
              } else {
                float z0 = bias0[jj + i];
#pragma unroll
                for (int cc = 0; cc < MAXIO; ++cc)
                  if (cc < d_in) z0 = fmaf(W0[cc * HP + j0 + jj + i], xr[cc], z0);
                keep = (ACT == 0) ? tanh_fast(z0) : z0;
              }
              float t, s1, s2;
              act_from_keep<ACT>(keep, t, s1, s2);
              if (is_val) {
                av[i] = t;
              } else if (!is_lap) {
                const float dz = ckd ? __ldcs(ckd + (s * HP + jj + i) * TP) : W0[dcol[s - 1] * HP + j0 + jj + i];
                av[i] = s1 * dz;
              } else {
                float q = 0.0f;
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                  const float dz = ckd ? __ldcs(ckd + ((1 + k) * HP + jj + i) * TP) : W0[dcol[k] * HP + j0 + jj + i];
                  q = fmaf(lapw[k] * dz, dz, q);
                }
                const float lz = ckd ? __ldcs(ckd + ((S - 1) * HP + jj + i) * TP) : 0.0f;
                av[i] = fmaf(s2, q, s1 * lz);
              }
            }
            {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) split2(av[2 * v] * asc, av[2 * v + 1] * asc, hi[v], lo[v]);
              const uint32_t off = (uint32_t)(((2 * chunk + h) ^ (prow & 7)) << 4);
              st_shared_v4(reg + 2 * IMG + off, hi[0], hi[1], hi[2], hi[3]);
              st_shared_v4(reg + 3 * IMG + off, lo[0], lo[1], lo[2], lo[3]);
            }
          }
          {
            const float m = warp_max_nonneg(mz_local);
            if (lane == 0) atomicMax(mzb + (set * B_MAXL + step) * 4 + s, __float_as_uint(m));
          }
          if (warp == 0 && lane == 0) flags[r] = my_live ? 1u : 0u;
          my_live = true;
          fence_async_smem();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_z[r]);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) ginv[s] = ginv_next[s];
#pragma unroll
        for (int q = 0; q < B_MAXL; ++q) { if (q == l) { epi[q] = my_epi; emin_prev[q] = emin_cur; } }
        live |= (1u << l);
      }
      // ------------------------------------------------ first layer: z = W0 x + b0, d_k z = W0[:, dcol_k], L z = 0
      {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          mbar_wait(&bar_d[s], (ph_d >> s) & 1u);
          ph_d ^= (1u << s);
        }
        fence_after();
        const float* bias0 = small + lay.bias(0) + j0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int jj = 8 * h;
          uint32_t raw[S][8];
#pragma unroll
          for (int s = 0; s < S; ++s) tmem_ld8_nowait(tmem + lane_base + s * 64 + j0 + jj, raw[s]);
          tmem_wait_ld();
          float zv[8], zk[NDX][8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float z0 = bias0[jj + i];
#pragma unroll
            for (int cc = 0; cc < MAXIO; ++cc)
              if (cc < d_in) z0 = fmaf(W0[cc * HP + j0 + jj + i], xr[cc], z0);
            float t, s1, s2;
            if constexpr (ACT == 0) {
              t = tanh_fast(z0);
              s1 = fmaf(-t, t, 1.0f);
              s2 = -2.0f * t * s1;
            } else {
              sincosf(z0, &t, &s1);
              s2 = -t;
            }
            const float gl = LAP ? __uint_as_float(raw[S - 1][i]) * ginv[S - 1] : 0.0f;
            float v = __uint_as_float(raw[0][i]) * ginv[0] * s1, q = 0.0f;
#pragma unroll
            for (int k = 0; k < ND; ++k) {
              const float dz = W0[dcol[k] * HP + j0 + jj + i];
              const float gk = __uint_as_float(raw[1 + k][i]) * ginv[1 + k];
              v = fmaf(gk * s2, dz, v);
              float vk = gk * s1;
              if (LAP) {
                q = fmaf(lapw[k] * dz, dz, q);
                vk = fmaf(2.0f * lapw[k] * gl * s2, dz, vk);
              }
              zk[k][i] = vk;
            }
            if (LAP) {
              const float s3 = (ACT == 0) ? -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f) : -s1;
              v = fmaf(gl, s3 * q, v);
            }
            zv[i] = v;
          }
          {
            const float rs = reduce8(zv, lane);
            if (lane < 8) acc[ACC_DB + 0 * HP + j0 + jj + lane] += rs;
          }
          float rk[NDX];
#pragma unroll
          for (int k = 0; k < ND; ++k) rk[k] = reduce8(zk[k], lane);
#pragma unroll
          for (int cc = 0; cc < MAXIO; ++cc)
            if (cc < d_in) {
              float w8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w8[i] = zv[i] * xr[cc];
              float rs = reduce8(w8, lane);
#pragma unroll
              for (int k = 0; k < ND; ++k)
                if (dcol[k] == cc) rs += rk[k];
              if (lane < 8) acc[ACC_W0 + cc * HP + j0 + jj + lane] += rs;
            }
        }
        fence_before();                        // the TMEM reads of this tile precede the next tile's adjoint MMAs (ordered by bar_z)
      }
    }

    // ------------------------------------------------------------------ flush everything
    drain_dw();
    for (int l = 1; l < L; ++l)
      if ((live >> l) & 1u) {
        int e = 0;
#pragma unroll
        for (int q = 0; q < B_MAXL; ++q) { if (q == l) e = epi[q]; }
        flush_dw(l, e);
      }
    __threadfence();
    named_bar_sync(BAR_ALL, 32 * ACT_WARPS);
    const int HP32 = lay32.HP;
    // small gradients: the four quadrant copies of every entry
    for (int v = tid; v < ACC_Q; v += 32 * ACT_WARPS) {
      const float sum = accs[v] + accs[ACC_Q + v] + accs[2 * ACC_Q + v] + accs[3 * ACC_Q + v];
      const int j = v % HP;
      if (j >= HP32) continue;
      if (v < ACC_W0) {
        const int l = v / HP;
        if (l < L) atomicAdd(a.gacc + (l == 0 ? lay32.b0() : lay32.b(l)) + j, (double)sum);
      } else if (v < ACC_WL) {
        const int cc = (v - ACC_W0) / HP;
        if (cc < d_in) atomicAdd(a.gacc + lay32.w0() + (size_t)cc * HP32 + j, (double)sum);
      } else {
        const int o = (v - ACC_WL) / HP;
        if (o < d_out) atomicAdd(a.gacc + lay32.wl() + (size_t)o * HP32 + j, (double)sum);
      }
    }
    if (tid < MAXIO && tid < d_out)
      atomicAdd(a.gacc + lay32.bl() + tid, (double)accs[ACC_BL + tid] + (double)accs[ACC_BL + MAXIO + tid] +
                                               (double)accs[ACC_BL + 2 * MAXIO + tid] + (double)accs[ACC_BL + 3 * MAXIO + tid]);
    // hidden matrices: rows j (hi) and 64 + j (lo) of the partial sums both belong to dW[j][i]
    if (ntiles > blockIdx.x) {
      for (int l = 1; l < L; ++l)
        for (int e = tid; e < HP * HP; e += 32 * ACT_WARPS) {
          const int i = e / HP, j = e % HP;
          const float* src = dwp + ((size_t)(l - 1) * HP + i) * 128;
          const double v = (double)__ldcg(src + j) + (double)__ldcg(src + 64 + j);
          if (j < HP32 && i < HP32) atomicAdd(a.gacc + lay32.w(l) + (size_t)j * HP32 + i, v);
        }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == ACT_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace tcp

using namespace tcp;

bool tcp_bwd_supported(const tpx_fcn_desc* net, int nd, int lap) {
  if (!tcp_supported(net, nd, lap)) return false;
  if (net->n_hidden > B_MAXL) return false;
  const tcp::Layout lay{net->n_hidden, net->d_in, net->d_out};
  if ((size_t)bwd_smem(lay).total + 1024 > 227 * 1024) return false;
  return true;
}

static size_t tcp_dwpart_bytes() { return (size_t)MAX_BWD_CTAS * (B_MAXL - 1) * HP * 128 * sizeof(float); }
static size_t tcp_tile_bytes(int L, int S) { return (tcp_ckpt_floats_per_tile(L, S) + (size_t)L * STATS) * sizeof(float); }

// saved buffer: [checkpoints of all tiles][stats of all tiles][dW partial sums]
size_t tcp_saved_bytes(const tpx_fcn_desc* net, int nd, int lap, int64_t n) {
  const int64_t ntiles = (n + TP - 1) / TP;
  return (size_t)ntiles * tcp_tile_bytes(net->n_hidden, 1 + nd + lap) + tcp_dwpart_bytes();
}
int64_t tcp_tiles_in(const tpx_fcn_desc* net, int nd, int lap, size_t bytes) {
  if (bytes <= tcp_dwpart_bytes()) return 0;
  return (int64_t)((bytes - tcp_dwpart_bytes()) / tcp_tile_bytes(net->n_hidden, 1 + nd + lap));
}
void tcp_saved_split(const tpx_fcn_desc* net, int nd, int lap, int64_t ntiles, float* base, float** ckpt, float** stats, float** dwpart) {
  const int S = 1 + nd + lap;
  *ckpt = base;
  *stats = base + (size_t)ntiles * tcp_ckpt_floats_per_tile(net->n_hidden, S);
  *dwpart = *stats + (size_t)ntiles * net->n_hidden * STATS;
}

template <int ND, int LAP, int ACT>
static int launch_tcp_bwd_act(const BwdArgs& a, int sm_count, cudaStream_t stream) {
  const size_t smem = (size_t)bwd_smem(a.lay).total + 1024;
  if (smem > 227 * 1024) return fail(TPX_E_UNSUPPORTED, "tcp_bwd_kernel needs %zu bytes of shared memory", smem);
  const int64_t ntiles = (a.n + TP - 1) / TP;
  int64_t grid = sm_count < ntiles ? sm_count : ntiles;
  if (grid > MAX_BWD_CTAS) grid = MAX_BWD_CTAS;
  if (grid < 1) return 0;
  cudaError_t e = cudaFuncSetAttribute(tcp_bwd_kernel<ND, LAP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tcp_bwd_kernel: shared memory attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
  tcp_bwd_kernel<ND, LAP, ACT><<<(unsigned)grid, THREADS, smem, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TPX_E_CUDA, "tcp_bwd_kernel launch (%lld CTAs, %zu bytes smem): %s", (long long)grid, smem, cudaGetErrorString(e));
  return 0;
}
template <int ND, int LAP>
static int launch_tcp_bwd(const BwdArgs& a, int act, int sm_count, cudaStream_t stream) {
  return act == TPX_ACT_SIN ? launch_tcp_bwd_act<ND, LAP, 1>(a, sm_count, stream) : launch_tcp_bwd_act<ND, LAP, 0>(a, sm_count, stream);
}
static int launch_tcp_bwd_any(const BwdArgs& a, int act, int nd, int lap, int sm_count, cudaStream_t stream) {
  switch (nd * 2 + lap) {
    case 0: return launch_tcp_bwd<0, 0>(a, act, sm_count, stream);
    case 2: return launch_tcp_bwd<1, 0>(a, act, sm_count, stream);
    case 3: return launch_tcp_bwd<1, 1>(a, act, sm_count, stream);
    case 4: return launch_tcp_bwd<2, 0>(a, act, sm_count, stream);
    case 5: return launch_tcp_bwd<2, 1>(a, act, sm_count, stream);
    case 6: return launch_tcp_bwd<3, 0>(a, act, sm_count, stream);
    default: return TPX_E_UNSUPPORTED;
  }
}

int tcp_backward(const tpx_fcn_desc* net, const tpx::Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                 int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* buf, size_t buf_bytes,
                 int saved, int sm_count, cudaStream_t stream) {
  BwdArgs a;
  a.lay = tcp::Layout{net->n_hidden, net->d_in, net->d_out};
  a.lay32 = lay32;
  a.jet = jet;
  a.img = img; a.gacc = gacc;
  const int nd = jet.nd;
  if (saved) {
    const size_t need = tcp_saved_bytes(net, nd, lap, n);
    if (buf_bytes < need) return fail(TPX_E_WORKSPACE, "checkpoint buffer of %zu bytes is too small (%zu needed)", buf_bytes, need);
    float *ck, *st, *dw;
    tcp_saved_split(net, nd, lap, (n + TP - 1) / TP, buf, &ck, &st, &dw);
    a.ckpt = ck; a.stats = st; a.dwpart = dw;
    a.x = x; a.n = n; a.gu = gu; a.gdu = gdu; a.glap = glap;
    return launch_tcp_bwd_any(a, net->activation, nd, lap, sm_count, stream);
  }
  // no saved checkpoints: chunks of (forward kernel writing checkpoints only) + reverse kernel through the workspace
  int64_t ctiles = tcp_tiles_in(net, nd, lap, buf_bytes);
  if (ctiles < 1) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes holds no tile", buf_bytes);
  const int64_t cap = (int64_t)sm_count * 4;            // keep the chunk's checkpoints L2-resident (148 x 4 tiles x 384 KB = 227 MB is too much: 2 tiles)
  (void)cap;
  const int64_t cpts = ctiles * TP;
  const int d_in = net->d_in, d_out = net->d_out;
  float *ck, *st, *dw;
  tcp_saved_split(net, nd, lap, ctiles, buf, &ck, &st, &dw);
  a.ckpt = ck; a.stats = st; a.dwpart = dw;
  for (int64_t p0 = 0; p0 < n; p0 += cpts) {
    const int64_t m = (n - p0 < cpts) ? n - p0 : cpts;
    int rc = tcp_forward(net, jet, lap, img, x + p0 * d_in, m, nullptr, nullptr, nullptr, ck, st, sm_count, stream);
    if (rc) return fail(rc, "tcp_fwd_kernel (checkpoint pass of the reverse sweep)");
    a.x = x + p0 * d_in; a.n = m;
    a.gu = gu ? gu + p0 * d_out : nullptr;
    a.gdu = gdu ? gdu + p0 * d_out * nd : nullptr;
    a.glap = glap ? glap + p0 * d_out : nullptr;
    rc = launch_tcp_bwd_any(a, net->activation, nd, lap, sm_count, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace tpx
