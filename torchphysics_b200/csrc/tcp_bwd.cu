// Points-as-lanes tensor-core reverse kernel (hidden width <= 64, tanh / sin, <= 4 streams); see tcp_common.cuh.
//
// One persistent CTA per SM, 16 activation warps (warp = chunk * 4 + quadrant: TMEM lanes 32 q .. 32 q + 31 = points,
// neurons 16 c .. 16 c + 15) + 1 MMA warp, one 128-point tile in flight.  It sweeps the layers of the tile backwards,
// reading the checkpoints of tcp_fwd.cu ([layer][stream][64 neurons][128 points]: tanh z | z, d_k z, L z).
//
// Reverse step of hidden matrix l (L-1 >= l >= 1): two PASSES, A = streams S-1 .. 2 (Laplacian, d_2, d_3), B = streams 1, 0
// (d_1, value).  Per pass
//   activation warps   g_s = adjoint of a_{l,s} (TMEM, product of the previous step; top step: WL^T ubar, thread-local)
//                      Zbar_{l,s} = adjoint jets(g, checkpoint_l),   a_{l-1,s} = jets(checkpoint_{l-1})
//                      both scaled by powers of two, split hi / lo (fp16) and stored as 128-byte rows (row = point,
//                      64 neurons, 128-byte swizzle) into a 64 KB staging region per stream  [Z_hi | Z_lo | A_hi | A_lo]
//                      (a ring of three regions: the stores of a stream reuse the region of the stream three before it)
//   MMA warp           adjoint   D_s[p][i] = sum_j Zbar_s[p][j] W_l[j][i]      A = the Z image (K-major), B = the image of
//                                row-major W_l read MN-major -> TMEM block s of X, consumed by the next step
//                      gradient  dW_l[j][i] += sum_p Zbar_s[p][j] a_s[p][i]    K = 128 points: A = [Z_hi ; Z_lo] read MN-major
//                                (M = 128), B = A_lo, then A_hi (MN-major, N = 64) -> 64 TMEM columns per hidden matrix
// The same Z image serves both products, and W needs no transposed copy.  The adjoint blocks are overwritten pass by pass,
// so pass A leaves what pass B needs from it in registers (the Laplacian adjoint and the cross terms of the value stream).
// W_l (16 KB) is fetched with cp.async.bulk behind the previous step.
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
constexpr int REGION = 65536;             // bytes of one staging region (one stream)
constexpr int NREG = 3;
constexpr int IMG = 16384;                // bytes of one image (128 rows x 128 B)
constexpr int BAR_TILE = 5, BAR_ALL = 6;
constexpr int B_WARPS = 20;               // 16 activation warps, the MMA warp, 3 idle warps that complete its warpgroup
constexpr int B_THREADS = 32 * B_WARPS;   // (setmaxnreg moves registers between whole warpgroups)
constexpr int ACT_REGS = 104, MMA_REGS = 64;   // balanced: 4 x 128 x (104 - 96) registers taken = 128 x (96 - 64) released

struct BwdSmem {
  int stage, wimg, small, acc, mub, mzb, flags, total;   // byte offsets from the 1 KB aligned base
};
// shared accumulators (fp32 atomics): [db: B_MAXL x 64][dW0: MAXIO x 64][dWL: MAXIO x 64][dbL: MAXIO]
constexpr int ACC_DB = 0, ACC_W0 = B_MAXL * HP, ACC_WL = ACC_W0 + MAXIO * HP, ACC_BL = ACC_WL + MAXIO * HP, ACC_TOTAL = ACC_BL + MAXIO;
__host__ __device__ inline BwdSmem bwd_smem(const Layout& lay) {
  BwdSmem s;
  s.stage = 0;
  s.wimg = NREG * REGION;                 // ONE weight image (hi | lo): the matrix of the current step
  s.small = s.wimg + 16384;
  s.acc = s.small + (int)lay.small() * 4;
  s.mub = s.acc + ACC_TOTAL * 4;          // [2 sets][4] tile maxima of |ubar_s|
  s.mzb = s.mub + 2 * 4 * 4;              // [2 sets][B_MAXL steps][4] tile maxima of |Zbar_s|
  s.flags = s.mzb + 2 * B_MAXL * 4 * 4;   // [2 passes] accumulate flag of the pass's first gradient MMA
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
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// streams of pass P (0 = A, 1 = B), in processing order
template <int S> __host__ __device__ constexpr int pass_count(int P) { return P == 0 ? (S > 2 ? S - 2 : 0) : (S < 2 ? S : 2); }
template <int S> __host__ __device__ constexpr int pass_stream(int P, int u) { return P == 0 ? S - 1 - u : (S < 2 ? 0 : 1 - u); }

template <int ND, int LAP, int ACT>
__global__ void __launch_bounds__(B_THREADS, 1) tcp_bwd_kernel(BwdArgs a) {
  constexpr int S = 1 + ND + LAP;
  constexpr int NDX = ND > 0 ? ND : 1;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_z[2], bar_f[NREG], bar_d[4], bar_w;
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Layout lay = a.lay;
  const BwdSmem so = bwd_smem(lay);
  const int L = lay.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* small = reinterpret_cast<float*>(sm + so.small);
  float* accs = reinterpret_cast<float*>(sm + so.acc);
  uint32_t* mub = reinterpret_cast<uint32_t*>(sm + so.mub);
  uint32_t* mzb = reinterpret_cast<uint32_t*>(sm + so.mzb);
  uint32_t* flags = reinterpret_cast<uint32_t*>(sm + so.flags);

  for (int v = tid; v < (int)lay.small(); v += B_THREADS) small[v] = __ldg(a.img + v);
  for (int v = tid; v < ACC_TOTAL; v += B_THREADS) accs[v] = 0.0f;
  for (int v = tid; v < 2 * 4 + 2 * B_MAXL * 4 + 4; v += B_THREADS) mub[v] = 0u;     // mub, mzb, flags are contiguous
  float* dwp = a.dwpart + (size_t)blockIdx.x * ((B_MAXL - 1) * HP * 128);
  for (int v = tid; v < (L - 1) * HP * 128; v += B_THREADS) dwp[v] = 0.0f;
  if (warp == ACT_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int r = 0; r < 2; ++r) mbar_init(&bar_z[r], ACT_WARPS);
    for (int r = 0; r < NREG; ++r) mbar_init(&bar_f[r], 1);
    for (int s = 0; s < 4; ++s) mbar_init(&bar_d[s], 1);
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const int64_t ntiles = (a.n + TP - 1) / TP;
  const uint32_t stage0 = smem_u32(sm + so.stage);

  if (warp >= ACT_WARPS) {
    reg_dec<MMA_REGS>();
    if (warp == ACT_WARPS) {
      // =========================== weight producer + MMA issuer ===========================
      const uint32_t wimg = smem_u32(sm + so.wimg);
      constexpr uint32_t idesc_adj = idesc_f16(128, HP, 0, 1);
      constexpr uint32_t idesc_dw = idesc_f16(128, HP, 1, 1);
      const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const int64_t total_steps = my_tiles * (L - 1);
      uint32_t gs = 0, pp = 0, ph_w = 0, ph_last = 0;
      constexpr int last_stream = 0;            // the adjoint product issued last in a step
      if (total_steps > 0 && elect_one()) bulk_load(wimg, a.img + lay.mat(L - 1, 0), 16384u, &bar_w);
      __syncwarp();
      int64_t step_no = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tile + gridDim.x < ntiles && lane == 0) {     // the next tile's checkpoints on their way into L2
          const float* nxt = a.ckpt + (size_t)(tile + gridDim.x) * ((size_t)(L - 1) * S * HP * TP);
          for (int c = 0; c < (L - 1) * S; ++c) bulk_prefetch_l2(nxt + (size_t)c * HP * TP, HP * TP * 4);
        }
        for (int l = L - 1; l >= 1; --l, ++step_no) {
          const uint32_t wh = wimg, wl = wimg + 8192u;
          const uint32_t tacc = tmem + T_ACC + (uint32_t)(l - 1) * 64u;
          if (L > 2 || step_no == 0) {
            mbar_wait(&bar_w, ph_w);
            ph_w ^= 1;
          }
#pragma unroll
          for (int P = 0; P < 2; ++P) {
            const int np = pass_count<S>(P);
            if (np == 0) continue;
            mbar_wait(&bar_z[pp & 1u], (pp >> 1) & 1u);
            fence_after();
            const uint32_t acc_flag = flags[pp & 1u];
            if (elect_one()) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (u >= np) continue;
                const int s = pass_stream<S>(P, u);
                const uint32_t r = (gs + u) % NREG;
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
                  mma_f16_ss(tacc, za, desc_sw128(al + kp * 2048, 16384, 1024), idesc_dw, (kp > 0 || u > 0) ? 1u : acc_flag);
                  mma_f16_ss(tacc, za, desc_sw128(ah + kp * 2048, 16384, 1024), idesc_dw, 1);
                }
                mma_commit(&bar_f[r]);
              }
            }
            __syncwarp();
            gs += np;
            ++pp;
          }
          // the weight image is free once the step's last adjoint product is done: fetch the next step's behind the
          // activation warps' work (a single hidden matrix stays resident)
          if (L > 2) {
            mbar_wait(&bar_d[last_stream], ph_last);
            ph_last ^= 1;
            if (step_no + 1 < total_steps) {
              const int ln = (l > 1) ? l - 1 : L - 1;
              if (elect_one()) bulk_load(wimg, a.img + lay.mat(ln, 0), 16384u, &bar_w);
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    reg_inc<ACT_REGS>();
    // =========================== activation warps ===========================
    const tpx::Layout lay32 = a.lay32;
    const int d_in = lay.d_in, d_out = lay.d_out;
    const int quad = warp & 3, chunk = warp >> 2;
    const int j0 = chunk * CHN;
    const int prow = quad * 32 + lane;                       // row of this thread's point in the images
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float* W0 = small + lay.w0() + j0;
    const float* WL = small + lay.wl() + j0;
    const float* bias0 = small + lay.bias(0) + j0;
    float lapw[NDX];
    int dcol[NDX];
#pragma unroll
    for (int k = 0; k < NDX; ++k) {
      lapw[k] = (LAP && k < ND) ? a.jet.lap_w[k] : 0.0f;
      dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
    }
    uint32_t gs = 0, pp = 0;                  // stream-slot and pass counters (same sequence as the MMA warp's)
    uint32_t ph_d = 0;                        // bit s: parity of the next completion of bar_d[s] to wait for
    int tcount = 0;                           // tiles accumulated since the last periodic flush
    // per hidden matrix l: the accumulator holds products (bit l of live); real exponent of its scale; smallest exponent
    // sum a stream of the previous tile allowed (all threads keep identical copies: every input of these is CTA-uniform)
    uint32_t live = 0;
    int epi[B_MAXL], emin_prev[B_MAXL];
#pragma unroll
    for (int l = 0; l < B_MAXL; ++l) { epi[l] = 0; emin_prev[l] = 1 << 20; }

    // all gradient MMAs issued so far have completed (the commit of a stream covers every earlier MMA)
    auto drain_dw = [&]() {
      if (gs >= 1) mbar_wait(&bar_f[(gs - 1) % NREG], ((gs - 1) / NREG) & 1u);
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
          for (int o = 0; o < d_out; ++o) {
            const float v = load_ub(s, o);
            m = fmaxf(m, fabsf(v));
            if (s == 0) {
              float sum = v;
#pragma unroll
              for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
              if (lane == 0) atomicAdd(accs + ACC_BL + o, sum);
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
        // held from pass A to pass B (the adjoint blocks are overwritten pass by pass): the Laplacian block and the cross
        // terms of the value stream's Zbar
        float gLh[CHN], rv[CHN];
#pragma unroll
        for (int i = 0; i < CHN; ++i) { gLh[i] = 0.0f; rv[i] = 0.0f; }

#pragma unroll
        for (int P = 0; P < 2; ++P) {
          constexpr int NPmax = 2;
          const int np = pass_count<S>(P);
          if (np == 0) continue;
          float zsc[NPmax], asc[NPmax];
          int esum_min = 1 << 20;
          int ezb[NPmax];
          // ---- adjoint blocks, bounds and scales of the pass's streams
#pragma unroll
          for (int u = 0; u < NPmax; ++u) {
            if (u >= np) continue;
            const int s = pass_stream<S>(P, u);
            const bool is_lap = LAP && s == S - 1, is_val = (s == 0);
            if (!top) {
              mbar_wait(&bar_d[s], (ph_d >> s) & 1u);
              ph_d ^= (1u << s);
              Ba[s] = wnorm * __uint_as_float(mzb[(set * B_MAXL + (step - 1)) * 4 + s]);
            } else {
              Ba[s] = wnorm * __uint_as_float(mub[set * 4 + s]);
            }
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
            ezb[u] = scale_exp_for(Bz);
            const int esum = (ezb[u] - 127) + (scale_exp_for(Ma[s]) - 127);
            esum_min = esum < esum_min ? esum : esum_min;
            ginv_next[s] = inv_of_exp(ezb[u]) * itau;
          }
          if (!top) fence_after();
          emin_cur = esum_min < emin_cur ? esum_min : emin_cur;
          // ---- accumulator scale: sticky; flush when this tile does not fit, or periodically
          {
            bool need = my_live && esum_min < my_epi;
            if (P == (pass_count<S>(0) > 0 ? 0 : 1) && periodic && my_live) need = true;
            if (need) {
              drain_dw();
              flush_dw(l, my_epi);
              my_live = false;
            }
            if (!my_live) my_epi = (esum_min < my_emin ? esum_min : my_emin) - 4;
          }
#pragma unroll
          for (int u = 0; u < NPmax; ++u) {
            if (u >= np) continue;
            zsc[u] = exp_to_float(ezb[u]);
            asc[u] = exp_to_float(my_epi - (ezb[u] - 127) + 127);
          }
          // ---- staging regions of the pass's streams: the MMAs of the streams three slots earlier have finished reading
          uint32_t reg[NPmax];
#pragma unroll
          for (int u = 0; u < NPmax; ++u) {
            if (u >= np) continue;
            const uint32_t slot = gs + u, r = slot % NREG;
            if (slot >= NREG) mbar_wait(&bar_f[r], ((slot / NREG) - 1) & 1u);
            reg[u] = stage0 + r * REGION + (uint32_t)prow * 128u;
          }
          float ubs[NPmax][MAXIO];             // top step: adjoints of the pass's output streams
          if (top) {
#pragma unroll
            for (int u = 0; u < NPmax; ++u)
#pragma unroll
              for (int o = 0; o < MAXIO; ++o) ubs[u][o] = (u < np) ? load_ub(pass_stream<S>(P, u < np ? u : 0), o) : 0.0f;
          }
          float mz_local[NPmax] = {0.0f, 0.0f};

#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int jj = 8 * h;              // neurons j0 + jj .. + 7
            const uint32_t off = (uint32_t)(((2 * chunk + h) ^ (prow & 7)) << 4);
            // -- adjoints of a_{l,s} for the pass's streams
            float g[NPmax][8];
            if (!top) {
              uint32_t raw[NPmax][8];
#pragma unroll
              for (int u = 0; u < NPmax; ++u)
                if (u < np) tmem_ld8_nowait(tmem + lane_base + pass_stream<S>(P, u) * 64 + j0 + jj, raw[u]);
              tmem_wait_ld();
#pragma unroll
              for (int u = 0; u < NPmax; ++u)
                if (u < np) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) g[u][i] = __uint_as_float(raw[u][i]) * ginv[pass_stream<S>(P, u)];
                }
            } else {
#pragma unroll
              for (int u = 0; u < NPmax; ++u)
                if (u < np) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float v = 0.0f;
                    for (int o = 0; o < d_out; ++o) v = fmaf(WL[o * HP + jj + i], ubs[u][o], v);
                    g[u][i] = v;
                  }
                }
            }
            // -- Zbar of 8 neurons for the pass's streams (aup: a_{L-1,s} of the top step, for dWL)
            float zb[NPmax][8], aup[NPmax][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float keep = __ldcs(cku + (0 * HP + jj + i) * TP);
              float t, s1, s2;
              act_from_keep<ACT>(keep, t, s1, s2);
              if (P == 0) {
                // streams S-1 .. 2: [Laplacian,] d_k for k >= 1
                float dz[NDX];
#pragma unroll
                for (int k = 0; k < ND; ++k) dz[k] = (LAP || k >= 1) ? __ldcs(cku + ((1 + k) * HP + jj + i) * TP) : 0.0f;
                float r = 0.0f, gl = 0.0f;
                if (LAP) {
                  float q = 0.0f;
#pragma unroll
                  for (int k = 0; k < ND; ++k) q = fmaf(lapw[k] * dz[k], dz[k], q);
                  const float lz = __ldcs(cku + ((S - 1) * HP + jj + i) * TP);
                  const float s3 = (ACT == 0) ? -2.0f * s1 * fmaf(-3.0f * t, t, 1.0f) : -s1;
                  gl = g[0][i];
                  zb[0][i] = gl * s1;
                  gLh[jj + i] = gl;
                  r = gl * fmaf(s3, q, s2 * lz);
                  if (top) aup[0][i] = fmaf(s2, q, s1 * lz);
                }
#pragma unroll
                for (int u = LAP; u < NPmax; ++u)
                  if (u < np) {
                    const int k = pass_stream<S>(0, u) - 1;      // direction of this stream (>= 1)
                    const int kc = (k >= 0 && k < ND) ? k : 0;
                    float v = g[u][i] * s1;
                    if (LAP) v = fmaf(2.0f * lapw[kc] * gl * s2, dz[kc], v);
                    zb[u][i] = v;
                    r = fmaf(g[u][i] * s2, dz[kc], r);
                    if (top) aup[u][i] = s1 * dz[kc];
                  }
                rv[jj + i] = r;
              } else {
                // streams 1 (direction 0), 0 (value)
                float r = rv[jj + i];
                if (ND >= 1) {
                  const float dz = __ldcs(cku + (1 * HP + jj + i) * TP);
                  float v = g[0][i] * s1;
                  if (LAP) v = fmaf(2.0f * lapw[0] * gLh[jj + i] * s2, dz, v);
                  zb[0][i] = v;
                  r = fmaf(g[0][i] * s2, dz, r);
                  if (top) aup[0][i] = s1 * dz;
                }
                constexpr int uv = (S < 2) ? 0 : 1;
                zb[uv][i] = fmaf(g[uv][i], s1, r);
                if (top) aup[uv][i] = t;
              }
            }
#pragma unroll
            for (int u = 0; u < NPmax; ++u)
              if (u < np) {
#pragma unroll
                for (int i = 0; i < 8; ++i) mz_local[u] = fmaxf(mz_local[u], fabsf(zb[u][i]));
              }
            if (P == 1) {                      // bias gradient of layer l (value stream)
              constexpr int uv = (S < 2) ? 0 : 1;
              const float rs = reduce8(zb[uv], lane);
              if (lane < 8) atomicAdd(accs + ACC_DB + l * HP + j0 + jj + lane, rs);
            }
            if (top) {                         // output-layer weights: dWL[o][j] += sum_p ubar_s[p][o] a_{L-1,s}[p][j]
              for (int o = 0; o < d_out; ++o) {
                float w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float v = 0.0f;
#pragma unroll
                  for (int u = 0; u < NPmax; ++u)
                    if (u < np) v = fmaf(ubs[u][o], aup[u][i], v);
                  w8[i] = v;
                }
                const float rs = reduce8(w8, lane);
                if (lane < 8) atomicAdd(accs + ACC_WL + o * HP + j0 + jj + lane, rs);
              }
            }
#pragma unroll
            for (int u = 0; u < NPmax; ++u)
              if (u < np) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) split2(zb[u][2 * v] * zsc[u], zb[u][2 * v + 1] * zsc[u], hi[v], lo[v]);
                st_shared_v4(reg[u] + off, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(reg[u] + IMG + off, lo[0], lo[1], lo[2], lo[3]);
              }
            // -- a_{l-1,s} of 8 neurons for the pass's streams
            float av[NPmax][8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float keep, dzl[NDX], lzl = 0.0f;
              if (ckd) {
                keep = __ldcs(ckd + (0 * HP + jj + i) * TP);
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                  const bool used = (P == 0) ? (LAP || k >= 1) : (k == 0);
                  dzl[k] = used ? __ldcs(ckd + ((1 + k) * HP + jj + i) * TP) : 0.0f;
                }
                if (LAP && P == 0) lzl = __ldcs(ckd + ((S - 1) * HP + jj + i) * TP);
              } else {
                float z0 = bias0[jj + i];
                for (int cc = 0; cc < d_in; ++cc) z0 = fmaf(W0[cc * HP + jj + i], xr[cc], z0);
                keep = (ACT == 0) ? tanh_fast(z0) : z0;
#pragma unroll
                for (int k = 0; k < ND; ++k) dzl[k] = W0[dcol[k] * HP + jj + i];
              }
              float t, s1, s2;
              act_from_keep<ACT>(keep, t, s1, s2);
              if (P == 0) {
                if (LAP) {
                  float q = 0.0f;
#pragma unroll
                  for (int k = 0; k < ND; ++k) q = fmaf(lapw[k] * dzl[k], dzl[k], q);
                  av[0][i] = fmaf(s2, q, s1 * lzl);
                }
#pragma unroll
                for (int u = LAP; u < NPmax; ++u)
                  if (u < np) {
                    const int k = pass_stream<S>(0, u) - 1;
                    const int kc = (k >= 0 && k < ND) ? k : 0;
                    av[u][i] = s1 * dzl[kc];
                  }
              } else {
                if (ND >= 1) av[0][i] = s1 * dzl[0];
                constexpr int uv = (S < 2) ? 0 : 1;
                av[uv][i] = t;
              }
            }
#pragma unroll
            for (int u = 0; u < NPmax; ++u)
              if (u < np) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) split2(av[u][2 * v] * asc[u], av[u][2 * v + 1] * asc[u], hi[v], lo[v]);
                st_shared_v4(reg[u] + 2 * IMG + off, hi[0], hi[1], hi[2], hi[3]);
                st_shared_v4(reg[u] + 3 * IMG + off, lo[0], lo[1], lo[2], lo[3]);
              }
          }
#pragma unroll
          for (int u = 0; u < NPmax; ++u)
            if (u < np) {
              const float m = warp_max_nonneg(mz_local[u]);
              if (lane == 0) atomicMax(mzb + (set * B_MAXL + step) * 4 + pass_stream<S>(P, u), __float_as_uint(m));
            }
          if (warp == 0 && lane == 0) flags[pp & 1u] = my_live ? 1u : 0u;
          my_live = true;
          fence_async_smem();
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_z[pp & 1u]);
          gs += np;
          ++pp;
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
            for (int cc = 0; cc < d_in; ++cc) z0 = fmaf(W0[cc * HP + jj + i], xr[cc], z0);
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
              const float dz = W0[dcol[k] * HP + jj + i];
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
            if (lane < 8) atomicAdd(accs + ACC_DB + 0 * HP + j0 + jj + lane, rs);
          }
          float rk[NDX];
#pragma unroll
          for (int k = 0; k < ND; ++k) rk[k] = reduce8(zk[k], lane);
          for (int cc = 0; cc < d_in; ++cc) {
            float w8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = zv[i] * xr[cc];
            float rs = reduce8(w8, lane);
#pragma unroll
            for (int k = 0; k < ND; ++k)
              if (dcol[k] == cc) rs += rk[k];
            if (lane < 8) atomicAdd(accs + ACC_W0 + cc * HP + j0 + jj + lane, rs);
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
    for (int v = tid; v < ACC_BL; v += 32 * ACT_WARPS) {
      const float sum = accs[v];
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
    if (tid < d_out) atomicAdd(a.gacc + lay32.bl() + tid, (double)accs[ACC_BL + tid]);
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
  tcp_bwd_kernel<ND, LAP, ACT><<<(unsigned)grid, B_THREADS, smem, stream>>>(a);
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
  const int64_t ctiles = tcp_tiles_in(net, nd, lap, buf_bytes);
  if (ctiles < 1) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes holds no tile", buf_bytes);
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
