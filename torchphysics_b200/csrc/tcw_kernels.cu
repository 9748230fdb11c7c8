// Tensor-core (tcgen05 / TMEM) jet kernels for hidden widths 65..256 ("wide" path), tanh or sin, <= 5 streams.
// (Written for one 128-neuron block per layer; widths 129..256 run the same kernels on 128 x 128 blocks, see Launch.)
//
// Formulation: NEURONS ARE THE UMMA ROWS.  For a hidden matrix W (128 x 128, zero padded) and a tile of
// P = 16 collocation points with S streams (value, d_k, weighted Laplacian), N = S * P columns:
//     D[j][n] = sum_i W[j][i] B[n][i]          M = 128 (output neuron j), N = 16 S, K = 128 (input neuron i)
//   * A = W lives in TMEM for the whole launch (hi | lo tf32 parts, 256 columns), so it costs no shared-memory
//     bandwidth; B = the activation streams of the tile in shared memory (K-major, 128-byte swizzle, hi | lo);
//     three tf32 products  W_lo B_hi + W_hi B_lo + W_hi B_hi  (small ones first: the accumulator rounds toward zero) give
//     fp32-grade accuracy (tc_common.cuh).
//   * TMEM lane = neuron: the thread that reads row j of D owns ALL streams of ALL points of the tile for its
//     neuron, so tanh, the jet formulas and their adjoints are thread-local -- no stream exchange through shared
//     memory (the width-64 kernels, whose UMMA rows are (stream, point), need one per layer).
//   * activations in HBM are [tile][n = s * 16 + p][128 neurons]: a warp of 32 consecutive neurons reads / writes
//     128-byte rows, and the same rows are 128-byte rows of the K-major B operand (conflict-free scalar stores).
//
// The sweep is LAYER-MAJOR: one launch per hidden matrix and direction (a 128 x 128 matrix in split tf32 is
// 128 KB: one matrix fills half of TMEM, and the reverse sweep's dW accumulators of all layers do not fit
// on chip).  Per launch every CTA (persistent, one per SM) keeps the matrix in TMEM and streams tiles through a
// two-stage pipeline: producer warps (global -> jets -> split -> shared), one MMA-issuer warp, four epilogue
// warps (TMEM -> jets / adjoint jets -> global).  Kernels:
//   w_layer_kernel<FWD>  ckpt_{l-1} (or x for l = 1) -> a_{l-1} -> z_l = W_l a_{l-1} + b_l -> ckpt_l = (tanh z, d_k z, L z)
//   w_out_fwd_kernel     ckpt_{L-1} -> u, du, lap      (linear output layer, CUDA cores: d_out x 128 per stream)
//   w_out_bwd_kernel     output adjoints -> dWL, dbL, Zbar_{L-1} (adjoint of the pre-activation streams), db_{L-1}
//   w_rev_kernel         one read of Zbar_l and ckpt_{l-1} for BOTH reverse products of matrix l:
//                        dW_l[j][i] += sum_n Zbar_l[n][j] a_{l-1}[n][i]   (A = Zbar^T in TMEM, written by its lane-bound
//                        loader warps; B = a^T in shared memory, 16-byte stores of a thread's own row; K = n) and
//                        Abar_{l-1} = W_l^T Zbar_l -> adjoint jets with ckpt_{l-1} -> Zbar_{l-1}, db_{l-1} (l = 1: dW0, db0)
//   w_dw_reduce_kernel   float64 sum over CTAs of w_rev_kernel's per-CTA fp32 partial sums of dW -> gradient image
//   w_out_gather[_t]_kernel / w_out_seed[_t]_kernel   9 or more outputs (hidden width <= 128; DeepONet trunks): the output layer runs as
//                        "matrix L" of w_layer_kernel<FWD> (raw-product exit) / w_rev_kernel, 128 outputs per launch; these move the
//                        products / output adjoints between the [tile][n][128] arrays and the caller's (optionally output-major) arrays
//   w_dw_kernel, w_layer_kernel<ADJ>   the same two products as separate launches (TPX_TCW_SPLIT=1, development)
// HBM traffic per point and hidden matrix: forward 2, reverse 3 activation rows of S * 512 bytes.  The forward launches
// are HBM-bound (~75 % of the copy peak), the fused reverse launch is bound by the tensor pipe and load latency
// (DESIGN.md section 3.2b).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fcn_common.cuh"
#include "tc_common.cuh"
#include "tc_kernels.h"
#include "tpx_internal.h"

// This file is compiled twice: as it is for tanh (namespace tcw, entry points tcw_*) and from tcw_sin.cu with
// TCW_ACT_SIN defined for tp.models.Sinus (namespace tcws, entry points tcws_*).
#ifdef TCW_ACT_SIN
#define TCW_NS tcws
#else
#define TCW_NS tcw
#endif

namespace tpx {
namespace TCW_NS {

using namespace tc;

constexpr int HW = 128;            // padded hidden width = UMMA M
constexpr int P = 16;              // points per tile
constexpr int EPI_WARPS = 4;       // one per TMEM quadrant (lane = neuron)
constexpr int PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS + PROD_WARPS;
constexpr int THREADS = 32 * (MMA_WARP + 1);
constexpr int MAXD = 4;            // d_in supported by the layer-0 code
constexpr int MAXO = 8;            // d_out supported by the CUDA-core output-layer kernels
constexpr int MAXO_WIDE = TPX_MAX_DOUT;   // ... 9 or more outputs (hidden width <= 128): the output layer is one more UMMA product per 128 outputs
constexpr int FWD = 0, ADJ = 1;
// The TMEM accumulator of dW rounds toward zero on every add (one ulp of the running sum): after thousands of steps of a
// CTA the sum would be ~1e-4 short.  Every FLUSH steps the loader warps move it into a per-CTA fp32 array in global
// memory (round-to-nearest fp32 reductions at L2, fire and forget; every element belongs to one thread) and the next
// product starts from zero.
constexpr int FLUSH = 64;
constexpr int MAX_REV_CTAS = 160;   // per-CTA partial arrays are sized for this many CTAs (B200: 148 SMs)

// tc::tanh_fast (ex2 + rcp + one Newton step, absolute error <= 3e-7).  The library tanhf gives the same parity
// margins (measured: the error of these kernels comes from the accumulator's round-toward-zero adds, not from tanh)
// and makes the l = 1 launches, which recompute layer 0 from x, compute-bound; -DTPX_TCW_LIBM_TANH selects it.
#ifdef TPX_TCW_LIBM_TANH
__device__ __forceinline__ float tanh_w(float x) { return tanhf(x); }
#else
__device__ __forceinline__ float tanh_w(float x) { return tanh_fast(x); }
#endif

// What the value stream of a checkpoint keeps, and the activation with its first two (three) derivatives from it: tanh keeps
// t = tanh z (every derivative is a polynomial in t), sin keeps z (reference models/activation_fn.py:76-85).
#ifdef TCW_ACT_SIN
__device__ __forceinline__ float keep_of(float z) { return z; }
__device__ __forceinline__ void act_derivs(float keep, float& s0, float& s1, float& s2) {
  sincosf(keep, &s0, &s1);
  s2 = -s0;
}
__device__ __forceinline__ float act_s3(float, float s1) { return -s1; }
#else
__device__ __forceinline__ float keep_of(float z) { return tanh_w(z); }
__device__ __forceinline__ void act_derivs(float keep, float& s0, float& s1, float& s2) {
  s0 = keep;
  s1 = fmaf(-keep, keep, 1.0f);
  s2 = -2.0f * keep * s1;
}
__device__ __forceinline__ float act_s3(float keep, float s1) { return -2.0f * s1 * fmaf(-3.0f * keep, keep, 1.0f); }
#endif

// mbarrier wait with a watchdog: a protocol error (or a fault in another role) traps after ~10 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  if (done) return;
  const long long t0 = clock64();
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 20000000000ll) __trap();
  } while (!done);
}

static int env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

// low part of the tf32 split of an activation: x - hi is exact; the MMA truncates it to tf32 (2^-23 |x|, toward zero).
// Rounding it here instead costs two more integer instructions per element in kernels that are bound by instruction
// issue (13 warps per SM) and changed the measured parity margins by < 10 % (-DTPX_TCW_ROUND_LO selects it).
#ifdef TPX_TCW_ROUND_LO
__device__ __forceinline__ float lo_part(float x, float hi) { return tf32_hi(x - hi); }
#else
__device__ __forceinline__ float lo_part(float x, float hi) { return x - hi; }
#endif

struct WArgs {
  Layout lay;            // FP32 parameter image (HP = 128)
  JetArgs jet;
  const float* packed;
  const float* x;
  int64_t n;
  int l;                 // hidden matrix 1..L-1
  const float* in;       // FWD: ckpt_{l-1} (unused for l = 1); ADJ / DW: Zbar_l
  const float* ck;       // ADJ / DW: ckpt_{l-1} (unused for l = 1)
  float* out;            // FWD: ckpt_l; ADJ: Zbar_{l-1} (unused for l = 1)
  double* gacc;
  // hidden width 129..256: the matrices are processed as 128 x 128 blocks, activations are stored as two half-arrays
  int in0;               // d_in > 4: layer 0 has its own kernels; matrix 1 reads its checkpoints from `in` / `ck` like any other
  int jb, kb;            // FWD: output-neuron block jb, input-neuron block kb.  REV: Zbar (output-neuron) block jb, input-neuron block kb
  int part_out, part_in; // the product is one K block of a sum: write it raw to `part` / add `part` to it before the epilogue
  float* part;           // half-array of partial products
  int64_t half_stride;   // output kernels: floats between the half-arrays of one layer
  float* dwpart;         // reverse kernel: per-CTA fp32 partial sums of dW, [CTA][input neuron i][output neuron j]
  long long* dbg;        // development (-DTPX_TIMELINE): clock64 stamps of steps 40..43 of CTA 0 of the reverse kernel
};

// layer-0 pre-activation of neuron i at one point: z = b0 + W0 x
struct Layer0 {
  float w[MAXD], b;
  int d_in;
  __device__ __forceinline__ void load(const Layout& lay, const float* __restrict__ packed, int i) {
    d_in = lay.d_in;
#pragma unroll
    for (int c = 0; c < MAXD; ++c) w[c] = (c < d_in) ? __ldg(packed + lay.w0() + (size_t)c * lay.HP + i) : 0.0f;
    b = __ldg(packed + lay.b0() + i);
  }
  __device__ __forceinline__ float col(int c) const {
    float r = w[0];
#pragma unroll
    for (int k = 1; k < MAXD; ++k) r = (c == k) ? w[k] : r;
    return r;
  }
};

// checkpoint streams (t = tanh z, d_k z, L z) -> activation streams, in place
template <int ND, int LAP>
__device__ __forceinline__ void jets_from_ckpt(float (&v)[1 + ND + LAP], const float (&lapw)[ND > 0 ? ND : 1]) {
  float s0, s1, s2;
  act_derivs(v[0], s0, s1, s2);
  v[0] = s0;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    const float dz = v[1 + k];
    if (LAP) q = fmaf(lapw[k] * dz, dz, q);
    v[1 + k] = s1 * dz;
  }
  if (LAP) v[ND + 1] = fmaf(s2, q, s1 * v[ND + 1]);
}

// adjoint of the activation jets: hb = adjoints of the activation streams (in), ck = checkpoint streams;
// on return hb holds the adjoints of the pre-activation streams (oracle/jet_numpy.py jet_backward)
template <int ND, int LAP>
__device__ __forceinline__ void adjoint_jets(float (&hb)[1 + ND + LAP], const float (&ck)[1 + ND + LAP],
                                             const float (&lapw)[ND > 0 ? ND : 1]) {
  float s0, s1, s2;
  act_derivs(ck[0], s0, s1, s2);
  float dot = 0.0f, q = 0.0f;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    dot = fmaf(hb[1 + k], ck[1 + k], dot);
    if (LAP) q = fmaf(lapw[k] * ck[1 + k], ck[1 + k], q);
  }
  float zb = fmaf(hb[0], s1, dot * s2);
  if (LAP) {
    const float lhb = hb[ND + 1];
    const float s3 = act_s3(ck[0], s1);
    zb = fmaf(lhb, fmaf(s3, q, s2 * ck[ND + 1]), zb);
#pragma unroll
    for (int k = 0; k < ND; ++k) hb[1 + k] = fmaf(hb[1 + k], s1, 2.0f * lhb * s2 * lapw[k] * ck[1 + k]);
    hb[ND + 1] = lhb * s1;
  } else {
#pragma unroll
    for (int k = 0; k < ND; ++k) hb[1 + k] *= s1;
  }
  hb[0] = zb;
}

// =============================================================================================================
// forward / adjoint layer kernel
// =============================================================================================================
template <int ND, int LAP, int MODE>
__global__ void __launch_bounds__(THREADS, 1) w_layer_kernel(WArgs a) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  constexpr int NDX = ND > 0 ? ND : 1;
  constexpr int STAGE = 2 * N * HW;          // floats: hi | lo
  // the heavy role gets 8 warps: forward = producers (jets + split), adjoint = epilogue (checkpoint reads, adjoint
  // jets, Zbar stores; two warps per TMEM quadrant, 8 points each, so that one batch of checkpoint loads per tile
  // is in flight while the warp waits for the tile's product)
  constexpr int NEPI = (MODE == FWD) ? 4 : 8;
  constexpr int NPROD = MMA_WARP - NEPI;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_dfull[2], bar_dempty[2];
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const Layout lay = a.lay;
  const int l = a.l;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.n + P - 1) / P;

  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], NPROD);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_dfull[s], 1);
      mbar_init(&bar_dempty[s], NEPI);
    }
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  // ---- the matrix (A operand) into TMEM: lane = row, columns [0,128) hi, [128,256) lo
  if (warp < EPI_WARPS) {
    const float* Wrow = a.packed + (MODE == FWD ? lay.w(l) : lay.wt(l)) + (size_t)(a.jb * HW + warp * 32 + lane) * lay.HP + a.kb * HW;
    const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < HW / 16; ++c) {
      float hi[16], lo[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wrow + c * 16) + v);
        hi[4 * v] = w4.x; hi[4 * v + 1] = w4.y; hi[4 * v + 2] = w4.z; hi[4 * v + 3] = w4.w;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float h = tf32_hi(hi[e]);
        lo[e] = tf32_hi(hi[e] - h);
        hi[e] = h;
      }
      tmem_st16(tw + c * 16, hi);
      tmem_st16(tw + HW + c * 16, lo);
    }
    tmem_wait_st();
  }
  fence_before();
  __syncthreads();
  fence_after();

  float lapw[NDX];
  int dcol[NDX];
#pragma unroll
  for (int k = 0; k < NDX; ++k) {
    lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
    dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
  }

  if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = idesc_tf32(128, N);
    const uint32_t sbase = smem_u32(sm);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait_w(&bar_full[st], ph);
      if (it >= 2) mbar_wait_w(&bar_dempty[st], ph ^ 1);
      fence_after();
      if (elect_one()) {
        const uint32_t tD = tmem + 256 + st * 128;
        const uint32_t bh = sbase + (uint32_t)st * (STAGE * 4), bl = bh + N * HW * 4;
        // the small correction products first: the accumulator rounds toward zero on every add, and an add
        // costs up to one ulp of the running sum (measured: interleaving the products triples the bias)
#pragma unroll
        for (int ks = 0; ks < HW / 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (N * 128) + (ks & 3) * 32;
          mma_tf32_ts(tD, tmem + HW + ks * 8, desc_k_sw128(bh + off), idesc, ks > 0);
          mma_tf32_ts(tD, tmem + ks * 8, desc_k_sw128(bl + off), idesc, 1);
        }
#pragma unroll
        for (int ks = 0; ks < HW / 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (N * 128) + (ks & 3) * 32;
          mma_tf32_ts(tD, tmem + ks * 8, desc_k_sw128(bh + off), idesc, 1);
        }
        mma_commit(&bar_empty[st]);
        mma_commit(&bar_dfull[st]);
      }
      __syncwarp();
    }
  } else if (warp >= NEPI) {
    // =========================== producers: B operand of the tile ===========================
    const int w = warp - NEPI, kb = w & 3, half = w >> 2;
    const int i = kb * 32 + lane;
    Layer0 l0;
    float dz0[NDX];
    const bool first = (l == 1) && !a.in0;     // layer 0 recomputed from x on the fly (d_in <= 4)
    if (MODE == FWD && first) {
      l0.load(lay, a.packed, a.kb * HW + i);
#pragma unroll
      for (int k = 0; k < NDX; ++k) dz0[k] = l0.col(dcol[k]);
    }
    int it = 0;
    constexpr bool PREFETCH = S <= 4;          // S = 5: no registers for it (128 per thread)
    float vn[8][S];                            // forward, checkpoints from memory: the next tile's rows, in flight during this tile
    if (PREFETCH && MODE == FWD && !first) {
      const float* src = a.in + ((size_t)blockIdx.x * N + half * 8) * HW + i;
#pragma unroll
      for (int pp = 0; pp < 8; ++pp)
#pragma unroll
        for (int s = 0; s < S; ++s) vn[pp][s] = __ldcs(src + (size_t)(s * P + pp) * HW);
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      float* bh = sm + st * STAGE;
      float* bl = bh + N * HW;
      if (MODE == FWD) {
        // rows n = s * P + p for this warp's 8 points
        float v[8][S];
        if (first) {
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const int64_t gp = tile * P + half * 8 + pp;
            float z = l0.b;
            if (gp < a.n) {
#pragma unroll
              for (int c = 0; c < MAXD; ++c)
                if (c < l0.d_in) z = fmaf(l0.w[c], __ldg(a.x + gp * l0.d_in + c), z);
            }
            v[pp][0] = keep_of(z);
#pragma unroll
            for (int k = 0; k < ND; ++k) v[pp][1 + k] = dz0[k];
            if (LAP) v[pp][S - 1] = 0.0f;
          }
        } else {
          if (PREFETCH) {
#pragma unroll
            for (int pp = 0; pp < 8; ++pp)
#pragma unroll
              for (int s = 0; s < S; ++s) v[pp][s] = vn[pp][s];
          } else {
            const float* src = a.in + ((size_t)tile * N + half * 8) * HW + i;
#pragma unroll
            for (int pp = 0; pp < 8; ++pp)
#pragma unroll
              for (int s = 0; s < S; ++s) v[pp][s] = __ldcs(src + (size_t)(s * P + pp) * HW);
          }
          if (PREFETCH && tile + gridDim.x < ntiles) {
            const float* src = a.in + ((size_t)(tile + gridDim.x) * N + half * 8) * HW + i;
#pragma unroll
            for (int pp = 0; pp < 8; ++pp)
#pragma unroll
              for (int s = 0; s < S; ++s) vn[pp][s] = __ldcs(src + (size_t)(s * P + pp) * HW);
          }
        }
        if (it >= 2) mbar_wait_w(&bar_empty[st], ph ^ 1);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int p = half * 8 + pp;
          jets_from_ckpt<ND, LAP>(v[pp], lapw);
          const bool valid = tile * P + p < a.n;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float val = valid ? v[pp][s] : 0.0f;
            const float h = tf32_hi(val);
            const int o = swz128(s * P + p, i, N);
            bh[o] = h;
            bl[o] = lo_part(val, h);
          }
        }
      } else {
        // all N rows of this warp's 32 neurons: plain copy + split of Zbar_l (half == 0: 4 producer warps)
        float v[N];
        const float* src = a.in + (size_t)tile * N * HW + i;
#pragma unroll
        for (int r = 0; r < N; ++r) v[r] = __ldcs(src + (size_t)r * HW);
        if (it >= 2) mbar_wait_w(&bar_empty[st], ph ^ 1);
#pragma unroll
        for (int r = 0; r < N; ++r) {
          const float h = tf32_hi(v[r]);
          const int o = swz128(r, i, N);
          bh[o] = h;
          bl[o] = lo_part(v[r], h);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[st]);
    }
  } else {
    // =========================== epilogue: row j of D = all streams / points of neuron j ===========================
    const int q = warp & 3, h = warp >> 2;     // TMEM quadrant; (adjoint) which 8 points of the tile
    const int j = q * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + 256;
    if (MODE == FWD) {
      const float bias = __ldg(a.packed + lay.b(l) + a.jb * HW + j);
      const bool part_in = a.part_in != 0, part_out = a.part_out != 0;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        float* dst = (part_out ? a.part : a.out) + (size_t)tile * N * HW + j;
        const float* pin = a.part + (size_t)tile * N * HW + j;
        // the other K block of a 256-wide layer: its partial products are in flight while the warp waits for this tile's
        // product, stream s + 1 while stream s is finished (loaded after the TMEM read they were an exposed HBM latency per stream)
        // S <= 3: all S x 16 values of the thread are requested before the wait (only four epilogue warps read this array:
        // with one stream in flight per thread, Little's law holds the whole launch to ~0.7 TB/s); S > 3: one stream ahead
        constexpr bool PART_ALL = S <= 3;
        float pv[PART_ALL ? S : 1][16];
        if (part_in) {
#pragma unroll
          for (int s = 0; s < (PART_ALL ? S : 1); ++s)
#pragma unroll
            for (int p = 0; p < P; ++p) pv[s][p] = __ldcs(pin + (size_t)(s * P + p) * HW);
        }
        mbar_wait_w(&bar_dfull[st], (it >> 1) & 1);
        fence_after();
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float d[16];
          tmem_ld16(tq + st * 128 + s * P, d);
          if (s == S - 1) {                    // D of this stage is in registers: hand it back to the MMA warp
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_dempty[st]);
          }
          if (part_in) {
#pragma unroll
            for (int p = 0; p < P; ++p) d[p] += pv[PART_ALL ? s : 0][p];
            if (!PART_ALL && s + 1 < S) {
#pragma unroll
              for (int p = 0; p < P; ++p) pv[0][p] = __ldcs(pin + (size_t)((s + 1) * P + p) * HW);
            }
          }
          if (s == 0 && !part_out) {
#pragma unroll
            for (int p = 0; p < P; ++p) d[p] = keep_of(d[p] + bias);
          }
#pragma unroll
          for (int p = 0; p < P; ++p) __stcs(dst + (size_t)(s * P + p) * HW, d[p]);
        }
      }
    } else {
      // adjoint jets with the checkpoints of layer l-1; for l = 1 the gradient of the first layer
      Layer0 l0;
      float dz0[NDX];
      if (l == 1) {
        l0.load(lay, a.packed, j);
#pragma unroll
        for (int k = 0; k < NDX; ++k) dz0[k] = l0.col(dcol[k]);
      }
      double acc_db = 0.0, acc_w0[MAXD];
#pragma unroll
      for (int c = 0; c < MAXD; ++c) acc_w0[c] = 0.0;
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        float part_db = 0.0f, part_w0[MAXD];
#pragma unroll
        for (int c = 0; c < MAXD; ++c) part_w0[c] = 0.0f;
        const float* cks = a.ck + ((size_t)tile * N + h * 8) * HW + j;
        float* dst = a.out + ((size_t)tile * N + h * 8) * HW + j;
        // this tile's checkpoints: in flight while the warp waits for the product
        float ck[8][S], xr[8][MAXD];
        if (l == 1) {
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const int64_t gp = tile * P + h * 8 + pp;
            float z = l0.b;
#pragma unroll
            for (int c = 0; c < MAXD; ++c) {
              xr[pp][c] = (gp < a.n && c < l0.d_in) ? __ldg(a.x + gp * l0.d_in + c) : 0.0f;
              z = fmaf(l0.w[c], xr[pp][c], z);
            }
            ck[pp][0] = keep_of(z);
#pragma unroll
            for (int k = 0; k < ND; ++k) ck[pp][1 + k] = dz0[k];
            if (LAP) ck[pp][S - 1] = 0.0f;
          }
        } else {
#pragma unroll
          for (int pp = 0; pp < 8; ++pp)
#pragma unroll
            for (int s = 0; s < S; ++s) ck[pp][s] = __ldcs(cks + (size_t)(s * P + pp) * HW);
        }
        mbar_wait_w(&bar_dfull[st], (it >> 1) & 1);
        fence_after();
        float ab[S][8];
#pragma unroll
        for (int s = 0; s < S; ++s) tmem_ld8(tq + st * 128 + s * P + h * 8, ab[s]);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_dempty[st]);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          float hb[S];
#pragma unroll
          for (int s = 0; s < S; ++s) hb[s] = ab[s][pp];
          adjoint_jets<ND, LAP>(hb, ck[pp], lapw);
          part_db += hb[0];
          if (l == 1) {
#pragma unroll
            for (int c = 0; c < MAXD; ++c) {
              float g = hb[0] * xr[pp][c];
#pragma unroll
              for (int k = 0; k < ND; ++k) g += (dcol[k] == c) ? hb[1 + k] : 0.0f;
              part_w0[c] += g;
            }
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) __stcs(dst + (size_t)(s * P + pp) * HW, hb[s]);
          }
        }
        acc_db += (double)part_db;
#pragma unroll
        for (int c = 0; c < MAXD; ++c) acc_w0[c] += (double)part_w0[c];
      }
      atomicAdd(a.gacc + (l == 1 ? lay.b0() : lay.b(l - 1)) + j, acc_db);
      if (l == 1) {
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < lay.d_in) atomicAdd(a.gacc + lay.w0() + (size_t)c * HW + j, acc_w0[c]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 512);
}

// =============================================================================================================
// dW kernel: dW_l[j][i] += sum_n Zbar_l[n][j] a_{l-1}[n][i]
// =============================================================================================================
template <int ND, int LAP>
__global__ void __launch_bounds__(THREADS, 1) w_dw_kernel(WArgs a) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  constexpr int NDX = ND > 0 ? ND : 1;
  constexpr int KA = (N + 31) / 32;          // 32-float K atoms of the a^T operand
  constexpr int HALF = KA * HW * 32;         // floats of its hi (or lo) part
  constexpr int STAGE = 2 * HALF;
  constexpr uint32_t ZCOLS = 160;            // TMEM columns of one Zbar^T stage (hi N | lo N), N <= 80
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_fa[2], bar_fz[2], bar_empty[2], bar_done;
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const Layout lay = a.lay;
  const int l = a.l;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.n + P - 1) / P;

  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_fa[s], PROD_WARPS);
      mbar_init(&bar_fz[s], EPI_WARPS);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    mbar_init_fence();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;

  float lapw[NDX];
  int dcol[NDX];
#pragma unroll
  for (int k = 0; k < NDX; ++k) {
    lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
    dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
  }

  if (warp == MMA_WARP) {
    constexpr uint32_t idesc = idesc_tf32(128, HW);
    const uint32_t sbase = smem_u32(sm);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait_w(&bar_fa[st], ph);
      mbar_wait_w(&bar_fz[st], ph);
      fence_after();
      if (elect_one()) {
        const uint32_t zh = tmem + HW + st * ZCOLS, zl = zh + N;
        const uint32_t ah = sbase + (uint32_t)st * (STAGE * 4), al = ah + HALF * 4;
#pragma unroll
        for (int ks = 0; ks < N / 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (HW * 128) + (ks & 3) * 32;
          mma_tf32_ts(tmem, zh + ks * 8, desc_k_sw128(ah + off), idesc, (it > 0 || ks > 0) ? 1u : 0u);
          mma_tf32_ts(tmem, zl + ks * 8, desc_k_sw128(ah + off), idesc, 1);
          mma_tf32_ts(tmem, zh + ks * 8, desc_k_sw128(al + off), idesc, 1);
        }
        mma_commit(&bar_empty[st]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar_done);
    __syncwarp();
  } else if (warp >= EPI_WARPS) {
    // =========================== producers: a_{l-1}^T, row = neuron i, K = n ===========================
    const int w = warp - EPI_WARPS, kb = w & 3, half = w >> 2;
    const int i = kb * 32 + lane;
    Layer0 l0;
    float dz0[NDX];
    if (l == 1) {
      l0.load(lay, a.packed, i);
#pragma unroll
      for (int k = 0; k < NDX; ++k) dz0[k] = l0.col(dcol[k]);
    }
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      float* ah = sm + st * STAGE;
      float* al = ah + HALF;
      float v[8][S];
      if (l == 1) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int64_t gp = tile * P + half * 8 + pp;
          float z = l0.b;
          if (gp < a.n) {
#pragma unroll
            for (int c = 0; c < MAXD; ++c)
              if (c < l0.d_in) z = fmaf(l0.w[c], __ldg(a.x + gp * l0.d_in + c), z);
          }
          v[pp][0] = keep_of(z);
#pragma unroll
          for (int k = 0; k < ND; ++k) v[pp][1 + k] = dz0[k];
          if (LAP) v[pp][S - 1] = 0.0f;
        }
      } else {
        const float* src = a.ck + ((size_t)tile * N + half * 8) * HW + i;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
#pragma unroll
          for (int s = 0; s < S; ++s) v[pp][s] = __ldcs(src + (size_t)(s * P + pp) * HW);
      }
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) jets_from_ckpt<ND, LAP>(v[pp], lapw);
      if (it >= 2) mbar_wait_w(&bar_empty[st], ph ^ 1);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int n0 = s * P + half * 8;                 // 8 consecutive K positions = two 16-byte chunks
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float4 h4, l4;
          float vv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) vv[e] = v[c * 4 + e][s];
          h4.x = tf32_hi(vv[0]); h4.y = tf32_hi(vv[1]); h4.z = tf32_hi(vv[2]); h4.w = tf32_hi(vv[3]);
          l4.x = lo_part(vv[0], h4.x); l4.y = lo_part(vv[1], h4.y); l4.z = lo_part(vv[2], h4.z); l4.w = lo_part(vv[3], h4.w);
          const int o = swz128(i, n0 + c * 4, HW);
          *reinterpret_cast<float4*>(ah + o) = h4;
          *reinterpret_cast<float4*>(al + o) = l4;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_fa[st]);
    }
  } else {
    // =========================== Zbar^T loaders (lane = neuron j), then the dW read-out ===========================
    const int j = warp * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const float* src = a.in + (size_t)tile * N * HW + j;
      const uint32_t zh = tq + HW + st * ZCOLS, zl = zh + N;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float hi[16], lo[16];
#pragma unroll
        for (int p = 0; p < P; ++p) hi[p] = __ldcs(src + (size_t)(s * P + p) * HW);
        if (s == 0 && it >= 2) mbar_wait_w(&bar_empty[st], ph ^ 1);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const float h = tf32_hi(hi[p]);
          lo[p] = lo_part(hi[p], h);
          hi[p] = h;
        }
        tmem_st16(zh + s * P, hi);
        tmem_st16(zl + s * P, lo);
      }
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_fz[st]);
    }
    mbar_wait_w(&bar_done, 0);
    fence_after();
    double* g = a.gacc + lay.w(l) + (size_t)j * HW;
#pragma unroll 1
    for (int c = 0; c < HW / 16; ++c) {
      float d[16];
      tmem_ld16(tq + c * 16, d);
#pragma unroll
      for (int e = 0; e < 16; ++e) atomicAdd(g + c * 16 + e, (double)d[e]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 512);
}

// =============================================================================================================
// fused reverse kernel of one hidden matrix: dW_l += Zbar_l^T a_{l-1}  AND  Zbar_{l-1} = adjoint jets(W_l^T Zbar_l)
// =============================================================================================================
// One read of Zbar_l and of ckpt_{l-1} serves both products (3 activation rows of HBM traffic per point instead of
// 5).  TMEM: W^T hi | lo (256 columns) + dW accumulator (128) leave 128 columns, so a step is a SUB-tile of 8
// points (rows r = 8 s + pp of the 16-point tile, half h):  D (NP = 8 S rounded up to 16 columns) and Zbar^T hi | lo
// (2 x 8 S columns), both single-buffered; the shared-memory operands are double-buffered:
//   Zn  [NP rows][128]   Zbar_l rows, K = j     B operand of the adjoint product (rows >= 8 S stay zero)
//   At  [128 rows][8 S]  a_{l-1}^T,   K = r     B operand of the gradient product
// Warps 0-3: Zbar loaders (lane = neuron j: Zn rows to shared memory, Zbar^T to TMEM).  Warps 4-11: two sets of four
// (lane = neuron i), set = stage = parity of the step: checkpoint loads -> a_{l-1} -> At, then the adjoint jets of the
// SAME checkpoints once the adjoint product has landed -> Zbar_{l-1} (l = 1: dW0, db0).  Warp 12 issues the MMAs:
// adjoint(it), gradient(it), adjoint(it+1), ...; Zbar^T of step it+1 is written while adjoint(it+1) runs.
// Every mbarrier is waited on by roles that can be at most one phase behind (per-stage barriers for the two sets).
template <int ND, int LAP, bool FIRST>      // FIRST: l = 1 (layer 0 is recomputed from x; dW0, db0 instead of Zbar_0)
__global__ void __launch_bounds__(THREADS, 1) w_rev_kernel(WArgs a) {
  constexpr int S = 1 + ND + LAP, N16 = S * P;
  constexpr int NS = 8 * S;                    // rows of a sub-tile
  constexpr int NP = (NS + 15) / 16 * 16;      // ... padded to a legal UMMA N
  constexpr int NDX = ND > 0 ? ND : 1;
  constexpr int KA = (NS + 31) / 32;
  constexpr int ZN = NP * HW;                  // floats of Zn hi (or lo)
  constexpr int AT = KA * HW * 32;             // floats of At hi (or lo)
  // Zn is double-buffered where shared memory allows (S <= 4); with one buffer (S = 5) it is rewritten while the
  // gradient product of the step runs, which leaves its ~1000-cycle write on the critical path
  constexpr int ZSTAGES = (S <= 4) ? 2 : 1;
  constexpr int AT_OFF = ZSTAGES * 2 * ZN, AT_STAGE = 2 * AT, RAW_OFF = AT_OFF + 2 * AT_STAGE, RAW_SLOT = NS * HW;
  constexpr uint32_t T_DW = 256, T_D = 384, T_ZH = 384 + NP, T_ZL = T_ZH + NS;
  static_assert(T_ZL + NS <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_fzn[2], bar_fzt, bar_fa[2], bar_dfull[2], bar_dempty, bar_zfree, bar_afree[2], bar_done;
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const Layout lay = a.lay;
  const int l = a.l;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nsub = (a.n + P - 1) / P * 2;
#ifdef TPX_TIMELINE
  long long* dbg = (a.dbg && !FIRST && blockIdx.x == 0 && lane == 0 && (warp == MMA_WARP || warp == 0 || warp == 4 || warp == 8)) ? a.dbg : nullptr;
#define TPX_WSTAMP(role, e) do { if (dbg && it >= 40 && it < 44) dbg[(role) * 64 + (it - 40) * 8 + (e)] = clock64(); } while (0)
#else
#define TPX_WSTAMP(role, e) do { } while (0)
#endif

  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_fa[s], 4);
      mbar_init(&bar_dfull[s], 1);
      mbar_init(&bar_afree[s], 1);
    }
    mbar_init(&bar_fzn[0], 4);
    mbar_init(&bar_fzn[1], 4);
    mbar_init(&bar_fzt, 4);
    mbar_init(&bar_dempty, 4);
    mbar_init(&bar_zfree, 1);
    mbar_init(&bar_done, 1);
    mbar_init_fence();
  }
  // padding rows of Zn (hi and lo) are zero for the whole launch
  if constexpr (NP > NS) {
    constexpr int PADF = (NP - NS) * HW;       // floats of the padding rows of one hi (or lo) part
    for (int v = tid; v < ZSTAGES * 2 * PADF; v += THREADS) {
      const int part = v / PADF, e = v % PADF;
      sm[part * ZN + swz128(NS + e / HW, e % HW, NP)] = 0.0f;
    }
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 4) {                              // W^T into TMEM: lane = input neuron i, K = output neuron j
    const float* Wrow = a.packed + lay.wt(l) + (size_t)(a.kb * HW + warp * 32 + lane) * lay.wts(l) + a.jb * HW;
    const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < HW / 16; ++c) {
      float hi[16], lo[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wrow + c * 16) + v);
        hi[4 * v] = w4.x; hi[4 * v + 1] = w4.y; hi[4 * v + 2] = w4.z; hi[4 * v + 3] = w4.w;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float h = tf32_hi(hi[e]);
        lo[e] = tf32_hi(hi[e] - h);
        hi[e] = h;
      }
      tmem_st16(tw + c * 16, hi);
      tmem_st16(tw + HW + c * 16, lo);
    }
    tmem_wait_st();
  }
  fence_before();
  __syncthreads();
  fence_after();

  float lapw[NDX];
  int dcol[NDX];
#pragma unroll
  for (int k = 0; k < NDX; ++k) {
    lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
    dcol[k] = (k < ND) ? a.jet.dcol[k] : 0;
  }

  if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_adj = idesc_tf32(128, NP), idesc_dw = idesc_tf32(128, HW);
    const uint32_t sbase = smem_u32(sm);
    int it = 0;
    for (int64_t sub = blockIdx.x; sub < nsub; sub += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph2 = (it >> 1) & 1;
      const uint32_t zn_h = sbase + (uint32_t)(it & (ZSTAGES - 1)) * (2 * ZN * 4), zn_l = zn_h + ZN * 4;
      const uint32_t at_h = sbase + (uint32_t)(AT_OFF + st * AT_STAGE) * 4, at_l = at_h + AT * 4;
      TPX_WSTAMP(0, 0);
      // one barrier per Zn buffer: with two buffers the loader may be a whole step ahead, and a single barrier two
      // phases ahead of its waiter passes the parity test the wrong way round (dead-lock)
      mbar_wait_w(&bar_fzn[it & (ZSTAGES - 1)], (ZSTAGES == 2 ? (it >> 1) : it) & 1);
      TPX_WSTAMP(0, 1);
      if (it >= 1) mbar_wait_w(&bar_dempty, (it - 1) & 1);
      fence_after();
      TPX_WSTAMP(0, 2);
      if (elect_one()) {
        // correction products first (the accumulator rounds toward zero on every add).  N = 48 costs as much as
        // N = 64 (measured); issuing it as N = 32 + N = 16 is slower still (small-N MMAs have a minimum cost)
#pragma unroll
        for (int ks = 0; ks < HW / 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (NP * 128) + (ks & 3) * 32;
          mma_tf32_ts(tmem + T_D, tmem + HW + ks * 8, desc_k_sw128(zn_h + off), idesc_adj, ks > 0);
          mma_tf32_ts(tmem + T_D, tmem + ks * 8, desc_k_sw128(zn_l + off), idesc_adj, 1);
        }
#pragma unroll
        for (int ks = 0; ks < HW / 8; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (NP * 128) + (ks & 3) * 32;
          mma_tf32_ts(tmem + T_D, tmem + ks * 8, desc_k_sw128(zn_h + off), idesc_adj, 1);
        }
        mma_commit(&bar_dfull[st]);
      }
      __syncwarp();
      TPX_WSTAMP(0, 3);
      mbar_wait_w(&bar_fzt, it & 1);
      TPX_WSTAMP(0, 4);
      mbar_wait_w(&bar_fa[st], ph2);
      fence_after();
      TPX_WSTAMP(0, 5);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < S; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (HW * 128) + (ks & 3) * 32;
          mma_tf32_ts(tmem + T_DW, tmem + T_ZL + ks * 8, desc_k_sw128(at_h + off), idesc_dw, (it % FLUSH != 0 || ks > 0) ? 1u : 0u);
          mma_tf32_ts(tmem + T_DW, tmem + T_ZH + ks * 8, desc_k_sw128(at_l + off), idesc_dw, 1);
        }
#pragma unroll
        for (int ks = 0; ks < S; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * (HW * 128) + (ks & 3) * 32;
          mma_tf32_ts(tmem + T_DW, tmem + T_ZH + ks * 8, desc_k_sw128(at_h + off), idesc_dw, 1);
        }
        mma_commit(&bar_zfree);
        mma_commit(&bar_afree[st]);
      }
      __syncwarp();
      TPX_WSTAMP(0, 6);
    }
    if (elect_one()) mma_commit(&bar_done);
    __syncwarp();
  } else if (warp < 4) {
    // =========================== Zbar_l loaders: lane = output neuron j ===========================
    const int j = warp * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
    int it = 0;
    float* dwp = a.dwpart + (size_t)blockIdx.x * (HW * HW) + j;     // element (i, j) at dwp[i * HW]: owned by this thread
#pragma unroll 8
    for (int c = 0; c < HW; ++c) dwp[c * HW] = 0.0f;
    // The rows of the next two steps are in flight (cp.async of a thread's OWN 8 S elements into a two-slot ring of raw
    // rows: nothing but the issuing thread reads them, so no barrier is involved): measured HBM latency under load is
    // ~3600 cycles, longer than a step, and a register prefetch of one step left the loader as the critical chain.
    auto issue = [&](int64_t sub, int slot) {
      if (sub < nsub) {
        const float* src = a.in + ((size_t)(sub >> 1) * N16 + (sub & 1) * 8) * HW + j;
        const uint32_t dst = smem_u32(sm + RAW_OFF + slot * RAW_SLOT + j);
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int pp = 0; pp < 8; ++pp)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)((s * 8 + pp) * HW * 4)),
                         "l"(src + (size_t)(s * P + pp) * HW) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(blockIdx.x, 0);
    issue((int64_t)blockIdx.x + gridDim.x, 1);
    for (int64_t sub = blockIdx.x; sub < nsub; sub += gridDim.x, ++it) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      float v[S][8];
      {
        const float* raw = sm + RAW_OFF + (it & 1) * RAW_SLOT + j;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) v[s][pp] = raw[(s * 8 + pp) * HW];
      }
      issue(sub + 2 * (int64_t)gridDim.x, it & 1);
      TPX_WSTAMP(1, 0);
      // this step's Zn buffer is free once the adjoint product of step it - ZSTAGES has completed
      float* zn_h = sm + (it & (ZSTAGES - 1)) * (2 * ZN);
      float* zn_l = zn_h + ZN;
      if (it >= ZSTAGES) mbar_wait_w(&bar_dfull[(it - ZSTAGES) & 1], ((it - ZSTAGES) >> 1) & 1);
      TPX_WSTAMP(1, 1);
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const float h = tf32_hi(v[s][pp]);
          const int o = swz128(s * 8 + pp, j, NP);
          zn_h[o] = h;
          zn_l[o] = lo_part(v[s][pp], h);
        }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_fzn[it & (ZSTAGES - 1)]);
      TPX_WSTAMP(1, 2);
      // Zbar^T in TMEM is free once the gradient product of step it-1 has completed
      if (it >= 1) mbar_wait_w(&bar_zfree, (it - 1) & 1);
      fence_after();
      TPX_WSTAMP(1, 3);
      if (it > 0 && it % FLUSH == 0) {           // all gradient products so far have completed; the next one starts from zero
#pragma unroll 1
        for (int c = 0; c < HW / 16; ++c) {
          float d[16];
          tmem_ld16(tq + T_DW + c * 16, d);
#pragma unroll
          for (int e = 0; e < 16; ++e) atomicAdd(dwp + (c * 16 + e) * HW, d[e]);   // RED.ADD.F32: no round trip
        }
        fence_before();
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float hi[8], lo[8];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          hi[pp] = tf32_hi(v[s][pp]);
          lo[pp] = lo_part(v[s][pp], hi[pp]);
        }
        tmem_st8(tq + T_ZH + s * 8, hi);
        tmem_st8(tq + T_ZL + s * 8, lo);
      }
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_fzt);
      TPX_WSTAMP(1, 4);
    }
    mbar_wait_w(&bar_done, 0);
    fence_after();
    // the rest of the TMEM accumulator joins the CTA's fp32 partial sums (a warp adds 128 consecutive bytes per instruction);
    // w_dw_reduce_kernel then sums the CTAs in float64.  (One float64 atomic per element and CTA straight into the gradient
    // image -- 32 different 1 KB-strided rows per warp instruction -- cost ~60 us per launch, more than the whole sweep of a
    // 16,384-point batch.)
#pragma unroll 1
    for (int c = 0; c < HW / 16; ++c) {
      float d[16];
      tmem_ld16(tq + T_DW + c * 16, d);
#pragma unroll
      for (int e = 0; e < 16; ++e) atomicAdd(dwp + (c * 16 + e) * HW, d[e]);
    }
  } else {
    // =========================== a_{l-1}^T producers + adjoint-jet epilogue: lane = input neuron i ===========================
    const int q = warp & 3, set = (warp - 4) >> 2;
    const int i = q * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    Layer0 l0;
    float dz0[NDX];
    if (FIRST) {
      l0.load(lay, a.packed, a.kb * HW + i);
#pragma unroll
      for (int k = 0; k < NDX; ++k) dz0[k] = l0.col(dcol[k]);
    }
    double acc_db = 0.0, acc_w0[MAXD];
#pragma unroll
    for (int c = 0; c < MAXD; ++c) acc_w0[c] = 0.0;
    float* at_h = sm + AT_OFF + set * AT_STAGE;
    float* at_l = at_h + AT;
    int it = set;
    for (int64_t sub = blockIdx.x + (int64_t)set * gridDim.x; sub < nsub; sub += 2 * (int64_t)gridDim.x, it += 2) {
      const uint32_t ph2 = (it >> 1) & 1;
      const size_t row0 = ((size_t)(sub >> 1) * N16 + (sub & 1) * 8) * HW + i;
      float ck[8][S], xr[8][FIRST ? MAXD : 1];
      if (FIRST) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int64_t gp = (sub >> 1) * P + (sub & 1) * 8 + pp;
          float z = l0.b;
#pragma unroll
          for (int c = 0; c < MAXD; ++c) {
            xr[pp][c] = (gp < a.n && c < l0.d_in) ? __ldg(a.x + gp * l0.d_in + c) : 0.0f;
            z = fmaf(l0.w[c], xr[pp][c], z);
          }
          ck[pp][0] = keep_of(z);
#pragma unroll
          for (int k = 0; k < ND; ++k) ck[pp][1 + k] = dz0[k];
          if (LAP) ck[pp][S - 1] = 0.0f;
        }
      } else {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
#pragma unroll
          for (int s = 0; s < S; ++s) ck[pp][s] = __ldcs(a.ck + row0 + (size_t)(s * P + pp) * HW);
      }
      TPX_WSTAMP(2 + set, 0);
      // At[set] is free once the gradient product of step it-2 has completed
      if (it >= 2) mbar_wait_w(&bar_afree[set], ph2 ^ 1);
      TPX_WSTAMP(2 + set, 1);
      {
        float av[8][S];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
#pragma unroll
          for (int s = 0; s < S; ++s) av[pp][s] = ck[pp][s];
          jets_from_ckpt<ND, LAP>(av[pp], lapw);
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float4 h4, l4;
            h4.x = tf32_hi(av[c * 4 + 0][s]); h4.y = tf32_hi(av[c * 4 + 1][s]); h4.z = tf32_hi(av[c * 4 + 2][s]); h4.w = tf32_hi(av[c * 4 + 3][s]);
            l4.x = lo_part(av[c * 4 + 0][s], h4.x); l4.y = lo_part(av[c * 4 + 1][s], h4.y);
            l4.z = lo_part(av[c * 4 + 2][s], h4.z); l4.w = lo_part(av[c * 4 + 3][s], h4.w);
            const int o = swz128(i, s * 8 + c * 4, HW);
            *reinterpret_cast<float4*>(at_h + o) = h4;
            *reinterpret_cast<float4*>(at_l + o) = l4;
          }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_fa[set]);
      TPX_WSTAMP(2 + set, 2);
      // 256-wide layers / several blocks of outputs: the partial sum of the other blocks, in flight during the wait (S <= 3: the
      // registers are there; loaded after the TMEM read it was an exposed HBM latency per step)
      constexpr bool PART_PREFETCH = S <= 3;
      float pv[PART_PREFETCH ? S : 1][8];
      if (PART_PREFETCH && a.part_in) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
#pragma unroll
          for (int s = 0; s < (PART_PREFETCH ? S : 1); ++s) pv[s][pp] = __ldcs(a.part + row0 + (size_t)(s * P + pp) * HW);
      }
      // the adjoint product of this step
      mbar_wait_w(&bar_dfull[set], ph2);
      fence_after();
      TPX_WSTAMP(2 + set, 3);
      float ab[S][8];
#pragma unroll
      for (int s = 0; s < S; ++s) tmem_ld8(tq + T_D + s * 8, ab[s]);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_dempty);
      TPX_WSTAMP(2 + set, 4);
      if (a.part_in) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (PART_PREFETCH) ab[s][pp] += pv[s][pp];
            else ab[s][pp] += __ldcs(a.part + row0 + (size_t)(s * P + pp) * HW);
          }
      }
      if (a.part_out) {                        // ... not complete yet: park the raw product, no adjoint jets
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
#pragma unroll
          for (int s = 0; s < S; ++s) __stcs(a.part + row0 + (size_t)(s * P + pp) * HW, ab[s][pp]);
        continue;
      }
      float part_db = 0.0f, part_w0[MAXD];
#pragma unroll
      for (int c = 0; c < MAXD; ++c) part_w0[c] = 0.0f;
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        float hb[S];
#pragma unroll
        for (int s = 0; s < S; ++s) hb[s] = ab[s][pp];
        adjoint_jets<ND, LAP>(hb, ck[pp], lapw);
        part_db += hb[0];
        if (FIRST) {
#pragma unroll
          for (int c = 0; c < MAXD; ++c) {
            float g = hb[0] * xr[pp][c];
#pragma unroll
            for (int k = 0; k < ND; ++k) g += (dcol[k] == c) ? hb[1 + k] : 0.0f;
            part_w0[c] += g;
          }
        } else {
#pragma unroll
          for (int s = 0; s < S; ++s) __stcs(a.out + row0 + (size_t)(s * P + pp) * HW, hb[s]);
        }
      }
      acc_db += (double)part_db;
#pragma unroll
      for (int c = 0; c < MAXD; ++c) acc_w0[c] += (double)part_w0[c];
      TPX_WSTAMP(2 + set, 5);
    }
    if (!a.part_out) {
      atomicAdd(a.gacc + ((FIRST || l == 1) ? lay.b0() : lay.b(l - 1)) + a.kb * HW + i, acc_db);
      if (FIRST) {
#pragma unroll
        for (int c = 0; c < MAXD; ++c)
          if (c < lay.d_in) atomicAdd(a.gacc + lay.w0() + (size_t)c * lay.HP + a.kb * HW + i, acc_w0[c]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 512);
}

// dW of one 128 x 128 block: float64 sum of the per-CTA fp32 partial sums [CTA][input neuron i][output neuron j] of w_rev_kernel,
// added to the gradient image (one writer per element: launches of one sweep are ordered on the stream)
__global__ void __launch_bounds__(256) w_dw_reduce_kernel(const float* __restrict__ dwpart, int nctas, double* __restrict__ g, int hp) {
  const int e = blockIdx.x * 256 + threadIdx.x;          // = i * HW + j
  if (e >= HW * HW) return;
  double acc = 0.0;
  int c = 0;
  for (; c + 4 <= nctas; c += 4) {
    const float v0 = __ldcg(dwpart + (size_t)c * (HW * HW) + e), v1 = __ldcg(dwpart + (size_t)(c + 1) * (HW * HW) + e);
    const float v2 = __ldcg(dwpart + (size_t)(c + 2) * (HW * HW) + e), v3 = __ldcg(dwpart + (size_t)(c + 3) * (HW * HW) + e);
    acc += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
  }
  for (; c < nctas; ++c) acc += (double)__ldcg(dwpart + (size_t)c * (HW * HW) + e);
  g[(size_t)(e % HW) * hp + e / HW] += acc;
}

// =============================================================================================================
// output layer (CUDA cores): one warp per point, lane owns neurons lane + 32 c
// =============================================================================================================
constexpr int OUT_THREADS = 256;

template <int ND, int LAP>
__global__ void __launch_bounds__(OUT_THREADS) w_out_fwd_kernel(WArgs a, float* __restrict__ u, float* __restrict__ du,
                                                                float* __restrict__ lapo) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  constexpr int NDX = ND > 0 ? ND : 1;
  __shared__ float WL[MAXO * 2 * HW];
  const Layout lay = a.lay;
  const int d_out = lay.d_out, hp = lay.HP, nblk = hp / HW;
  for (int v = threadIdx.x; v < d_out * hp; v += OUT_THREADS) WL[v] = __ldg(a.packed + lay.wl() + v);
  __syncthreads();
  float lapw[NDX];
#pragma unroll
  for (int k = 0; k < NDX; ++k) lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (OUT_THREADS / 32);
  for (int64_t pt = (int64_t)blockIdx.x * (OUT_THREADS / 32) + (threadIdx.x >> 5); pt < a.n; pt += nwarps) {
    const int64_t tile = pt / P;
    const int p = (int)(pt % P);
    float acc[MAXO][S];
#pragma unroll
    for (int o = 0; o < MAXO; ++o)
#pragma unroll
      for (int s = 0; s < S; ++s) acc[o][s] = 0.0f;
    for (int hb = 0; hb < nblk; ++hb) {        // half-arrays of a 256-wide layer
      const float* src = a.in + (size_t)hb * a.half_stride + ((size_t)tile * N + p) * HW + lane;
      float v[4][S];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int s = 0; s < S; ++s) v[c][s] = __ldcs(src + (size_t)s * P * HW + c * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) jets_from_ckpt<ND, LAP>(v[c], lapw);
#pragma unroll
      for (int o = 0; o < MAXO; ++o)
        if (o < d_out) {
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[o][s] = fmaf(WL[o * hp + hb * HW + c * 32 + lane], v[c][s], acc[o][s]);
        }
    }
    for (int o = 0; o < d_out; ++o) {
      float r[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float t = acc[0][s];
#pragma unroll
        for (int oo = 1; oo < MAXO; ++oo) t = (o == oo) ? acc[oo][s] : t;
        r[s] = t;
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1)
#pragma unroll
        for (int s = 0; s < S; ++s) r[s] += __shfl_xor_sync(0xFFFFFFFFu, r[s], m);
      if (lane == 0) {
        u[pt * d_out + o] = r[0] + __ldg(a.packed + lay.bl() + o);
#pragma unroll
        for (int k = 0; k < ND; ++k) du[(pt * d_out + o) * ND + k] = r[1 + k];
        if (LAP) lapo[pt * d_out + o] = r[S - 1];
      }
    }
  }
}

// output adjoints -> Zbar_{L-1} (written for every row of every tile, zeros for the padding points), dWL, dbL, db_{L-1}
template <int ND, int LAP, int DO>
__global__ void __launch_bounds__(OUT_THREADS, (DO <= 2) ? 3 : 2) w_out_bwd_kernel(WArgs a, const float* __restrict__ gu, const float* __restrict__ gdu,
                                                                const float* __restrict__ glap) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  constexpr int NDX = ND > 0 ? ND : 1;
  __shared__ float WL[DO * HW];
  __shared__ float red[(DO + 1) * HW + DO];
  const Layout lay = a.lay;
  const int d_out = lay.d_out;
  const int hp = lay.HP, hb = a.kb;              // this launch: neurons [128 hb, 128 hb + 128) of the last hidden layer
  for (int v = threadIdx.x; v < d_out * HW; v += OUT_THREADS) WL[v] = __ldg(a.packed + lay.wl() + (size_t)(v / HW) * hp + hb * HW + v % HW);
  for (int v = threadIdx.x; v < (DO + 1) * HW + DO; v += OUT_THREADS) red[v] = 0.0f;
  __syncthreads();
  float lapw[NDX];
#pragma unroll
  for (int k = 0; k < NDX; ++k) lapw[k] = (k < ND) ? a.jet.lap_w[k] : 0.0f;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (OUT_THREADS / 32);
  const int64_t npad = (a.n + P - 1) / P * P;
  float acc_wl[DO][4], acc_db[4], acc_bl = 0.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    acc_db[c] = 0.0f;
#pragma unroll
    for (int o = 0; o < DO; ++o) acc_wl[o][c] = 0.0f;
  }
  for (int64_t pt = (int64_t)blockIdx.x * (OUT_THREADS / 32) + (threadIdx.x >> 5); pt < npad; pt += nwarps) {
    const int64_t tile = pt / P;
    const int p = (int)(pt % P);
    const bool valid = pt < a.n;
    const float* src = a.in + ((size_t)tile * N + p) * HW + lane;
    float* dst = a.out + ((size_t)tile * N + p) * HW + lane;
    float ck[4][S];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int s = 0; s < S; ++s) ck[c][s] = __ldcs(src + (size_t)s * P * HW + c * 32);
    float g[DO][S];
#pragma unroll
    for (int o = 0; o < DO; ++o) {
      const bool on = valid && o < d_out;
      g[o][0] = (on && gu) ? __ldg(gu + pt * d_out + o) : 0.0f;
#pragma unroll
      for (int k = 0; k < ND; ++k) g[o][1 + k] = (on && gdu) ? __ldg(gdu + (pt * d_out + o) * ND + k) : 0.0f;
      if (LAP) g[o][S - 1] = (on && glap) ? __ldg(glap + pt * d_out + o) : 0.0f;
    }
    if (lane < DO) {
      float t = 0.0f;
#pragma unroll
      for (int o = 0; o < DO; ++o) t = (lane == o) ? g[o][0] : t;
      acc_bl += t;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float act[S], hb[S];
#pragma unroll
      for (int s = 0; s < S; ++s) { act[s] = ck[c][s]; hb[s] = 0.0f; }
      jets_from_ckpt<ND, LAP>(act, lapw);
#pragma unroll
      for (int o = 0; o < DO; ++o) {
        if (o < d_out) {
          const float w = WL[o * HW + c * 32 + lane];
          float dw = 0.0f;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            hb[s] = fmaf(w, g[o][s], hb[s]);
            dw = fmaf(g[o][s], act[s], dw);
          }
          acc_wl[o][c] += dw;
        }
      }
      adjoint_jets<ND, LAP>(hb, ck[c], lapw);
      acc_db[c] += hb[0];
#pragma unroll
      for (int s = 0; s < S; ++s) __stcs(dst + (size_t)s * P * HW + c * 32, hb[s]);
    }
  }
  // CTA-level reduction in shared memory, then one float64 atomic per element and CTA
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    atomicAdd(&red[DO * HW + c * 32 + lane], acc_db[c]);
#pragma unroll
    for (int o = 0; o < DO; ++o)
      if (o < d_out) atomicAdd(&red[o * HW + c * 32 + lane], acc_wl[o][c]);
  }
  if (lane < DO) atomicAdd(&red[(DO + 1) * HW + lane], acc_bl);
  __syncthreads();
  for (int v = threadIdx.x; v < d_out * HW; v += OUT_THREADS)
    atomicAdd(a.gacc + lay.wl() + (size_t)(v / HW) * hp + hb * HW + v % HW, (double)red[v]);
  for (int v = threadIdx.x; v < HW; v += OUT_THREADS) atomicAdd(a.gacc + lay.b(lay.L - 1) + hb * HW + v, (double)red[DO * HW + v]);
  if (hb == 0 && threadIdx.x < d_out) atomicAdd(a.gacc + lay.bl() + threadIdx.x, (double)red[(DO + 1) * HW + threadIdx.x]);
}

// ---- 9 or more outputs (Layout::wide_out: DeepONet trunks, the DeepONet inner product folded into the output layer): the output
// layer runs as matrix L of the layer kernels, one 128-output block (a.jb) at a time ------------------------------------------
// forward: w_layer_kernel<FWD> with part_out leaves the raw products W_L a_{L-1} as [tile][n = s * 16 + p][128]; this kernel adds the
// bias to the value stream and writes the caller's u / du / lap arrays (a warp reads 128-byte rows, consecutive outputs)
template <int ND, int LAP>
__global__ void __launch_bounds__(OUT_THREADS) w_out_gather_kernel(WArgs a, const float* __restrict__ raw, float* __restrict__ u,
                                                                   float* __restrict__ du, float* __restrict__ lapo) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  const int d_out = a.lay.d_out, c0 = a.jb * HW;
  const int cols = d_out - c0 < HW ? d_out - c0 : HW;          // outputs of this block
  const int64_t total = a.n * cols;
  for (int64_t e = (int64_t)blockIdx.x * OUT_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * OUT_THREADS) {
    const int64_t pt = e / cols;
    const int o = (int)(e % cols);
    const float* src = raw + ((size_t)(pt / P) * N + (size_t)(pt % P)) * HW + o;
    const int64_t dst = pt * d_out + c0 + o;
    u[dst] = __ldcs(src) + __ldg(a.packed + a.lay.bl() + c0 + o);
#pragma unroll
    for (int k = 0; k < ND; ++k) du[dst * ND + k] = __ldcs(src + (size_t)(1 + k) * P * HW);
    if (LAP) lapo[dst] = __ldcs(src + (size_t)(S - 1) * P * HW);
  }
}

// reverse: the output adjoints of block a.jb as Zbar_L (every row of every tile, zeros for padding points and neurons >= d_out) for
// w_rev_kernel with l = L, and dbL.  One CTA = 128 threads = the 128 neuron columns, grid-stride over the points.
template <int ND, int LAP>
__global__ void __launch_bounds__(HW) w_out_seed_kernel(WArgs a, const float* __restrict__ gu, const float* __restrict__ gdu,
                                                         const float* __restrict__ glap) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  const int d_out = a.lay.d_out, o = threadIdx.x, c = a.jb * HW + o;
  const int64_t npad = (a.n + P - 1) / P * P;
  double acc_bl = 0.0;
  for (int64_t pt = blockIdx.x; pt < npad; pt += gridDim.x) {
    const bool on = pt < a.n && c < d_out;
    float* dst = a.out + ((size_t)(pt / P) * N + (size_t)(pt % P)) * HW + o;
    const float g0 = (on && gu) ? __ldg(gu + pt * d_out + c) : 0.0f;
    acc_bl += (double)g0;
    __stcs(dst, g0);
#pragma unroll
    for (int k = 0; k < ND; ++k) __stcs(dst + (size_t)(1 + k) * P * HW, (on && gdu) ? __ldg(gdu + (pt * d_out + c) * ND + k) : 0.0f);
    if (LAP) __stcs(dst + (size_t)(S - 1) * P * HW, (on && glap) ? __ldg(glap + pt * d_out + c) : 0.0f);
  }
  if (c < d_out && a.kb == 0) atomicAdd(a.gacc + a.lay.bl() + c, acc_bl);
}

// the same two kernels for output-major streams (JetArgs::out_major: u (d_out, n), du (nd, d_out, n), lap (d_out, n), row stride
// out_ld): one CTA per tile of 16 points, transposed through shared memory so that the raw rows are read / written as 512-byte rows
// and the caller's arrays as 64-byte runs of 16 consecutive points
template <int ND, int LAP>
__global__ void __launch_bounds__(OUT_THREADS) w_out_gather_t_kernel(WArgs a, const float* __restrict__ raw, float* __restrict__ u,
                                                                     float* __restrict__ du, float* __restrict__ lapo) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  __shared__ float t[N][HW + 1];
  const int d_out = a.lay.d_out, c0 = a.jb * HW;
  const int cols = d_out - c0 < HW ? d_out - c0 : HW;
  const int64_t ntiles = (a.n + P - 1) / P, ld = a.jet.out_ld;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float* src = raw + (size_t)tile * N * HW;
    for (int v = threadIdx.x; v < N * HW; v += OUT_THREADS) t[v / HW][v % HW] = __ldcs(src + v);
    __syncthreads();
    const int p = threadIdx.x % P;
    const int64_t pt = tile * P + p;
    if (pt < a.n) {
      for (int o = threadIdx.x / P; o < cols; o += OUT_THREADS / P) {
        const int64_t row = c0 + o;
        u[row * ld + pt] = t[p][o] + __ldg(a.packed + a.lay.bl() + row);
#pragma unroll
        for (int k = 0; k < ND; ++k) du[((int64_t)k * d_out + row) * ld + pt] = t[(1 + k) * P + p][o];
        if (LAP) lapo[row * ld + pt] = t[(S - 1) * P + p][o];
      }
    }
    __syncthreads();
  }
}

template <int ND, int LAP>
__global__ void __launch_bounds__(OUT_THREADS) w_out_seed_t_kernel(WArgs a, const float* __restrict__ gu, const float* __restrict__ gdu,
                                                                   const float* __restrict__ glap) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  __shared__ float t[N][HW + 1];
  __shared__ float sdb[HW];
  const int d_out = a.lay.d_out, c0 = a.jb * HW;
  const int64_t ntiles = (a.n + P - 1) / P, ld = a.jet.out_ld;
  for (int v = threadIdx.x; v < HW; v += OUT_THREADS) sdb[v] = 0.0f;
  float acc[HW / 32];                                        // value-stream adjoints of outputs lane + 32 q, summed over this CTA's points
#pragma unroll
  for (int q = 0; q < HW / 32; ++q) acc[q] = 0.0f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p = threadIdx.x % P;
    const int64_t pt = tile * P + p;
    for (int o = threadIdx.x / P; o < HW; o += OUT_THREADS / P) {
      const int64_t row = c0 + o;
      const bool on = pt < a.n && row < d_out;
      t[p][o] = (on && gu) ? __ldg(gu + row * ld + pt) : 0.0f;
#pragma unroll
      for (int k = 0; k < ND; ++k) t[(1 + k) * P + p][o] = (on && gdu) ? __ldg(gdu + ((int64_t)k * d_out + row) * ld + pt) : 0.0f;
      if (LAP) t[(S - 1) * P + p][o] = (on && glap) ? __ldg(glap + row * ld + pt) : 0.0f;
    }
    __syncthreads();
    float* dst = a.out + (size_t)tile * N * HW;
    for (int r = warp; r < N; r += OUT_THREADS / 32) {
#pragma unroll
      for (int q = 0; q < HW / 32; ++q) {
        const float v = t[r][q * 32 + lane];
        __stcs(dst + (size_t)r * HW + q * 32 + lane, v);
        if (r < P) acc[q] += v;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < HW / 32; ++q) atomicAdd(&sdb[q * 32 + lane], acc[q]);
  __syncthreads();
  for (int o = threadIdx.x; o < HW; o += OUT_THREADS)
    if (c0 + o < d_out && a.kb == 0) atomicAdd(a.gacc + a.lay.bl() + c0 + o, (double)sdb[o]);
}

// ---- d_in = 5..8 (TPX_MAX_DIN): the layer kernels keep layer 0 in registers for at most MAXD inputs, so a wider input layer gets
// kernels of its own: its checkpoints (keep(z0), d_k z0 = W0[:, dcol_k], L z0 = 0) are written once per sweep and matrix 1 reads them
// like any other layer's; the reverse sweep leaves Zbar_0 in memory and w_in_bwd_kernel turns it into dW0 (db0 comes from w_rev_kernel).
// One CTA = the 128 neurons of block a.kb, grid-stride over the tiles.
template <int ND, int LAP>
__global__ void __launch_bounds__(HW) w_in_fwd_kernel(WArgs a) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  const Layout lay = a.lay;
  const int d_in = lay.d_in, j = a.kb * HW + threadIdx.x;
  float w[TPX_MAX_DIN];
#pragma unroll
  for (int c = 0; c < TPX_MAX_DIN; ++c) w[c] = (c < d_in) ? __ldg(a.packed + lay.w0() + (size_t)c * lay.HP + j) : 0.0f;
  const float b = __ldg(a.packed + lay.b0() + j);
  float dz[ND > 0 ? ND : 1];
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    float r = w[0];
#pragma unroll
    for (int c = 1; c < TPX_MAX_DIN; ++c) r = (a.jet.dcol[k] == c) ? w[c] : r;
    dz[k] = r;
  }
  const int64_t ntiles = (a.n + P - 1) / P;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    float* dst = a.out + (size_t)tile * N * HW + threadIdx.x;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      const int64_t gp = tile * P + p;
      float z = b;
      if (gp < a.n) {
#pragma unroll
        for (int c = 0; c < TPX_MAX_DIN; ++c)
          if (c < d_in) z = fmaf(w[c], __ldg(a.x + gp * d_in + c), z);
      }
      __stcs(dst + (size_t)p * HW, keep_of(z));
#pragma unroll
      for (int k = 0; k < ND; ++k) __stcs(dst + (size_t)((1 + k) * P + p) * HW, dz[k]);
      if (LAP) __stcs(dst + (size_t)((S - 1) * P + p) * HW, 0.0f);
    }
  }
}

template <int ND, int LAP>
__global__ void __launch_bounds__(HW) w_in_bwd_kernel(WArgs a) {
  constexpr int S = 1 + ND + LAP, N = S * P;
  const Layout lay = a.lay;
  const int d_in = lay.d_in, j = a.kb * HW + threadIdx.x;
  double acc[TPX_MAX_DIN];
#pragma unroll
  for (int c = 0; c < TPX_MAX_DIN; ++c) acc[c] = 0.0;
  const int64_t ntiles = (a.n + P - 1) / P;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float* src = a.in + (size_t)tile * N * HW + threadIdx.x;      // Zbar_0 of this block
    float part[TPX_MAX_DIN];
#pragma unroll
    for (int c = 0; c < TPX_MAX_DIN; ++c) part[c] = 0.0f;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      const int64_t gp = tile * P + p;
      if (gp >= a.n) break;
      const float zb = __ldcs(src + (size_t)p * HW);
#pragma unroll
      for (int c = 0; c < TPX_MAX_DIN; ++c)
        if (c < d_in) part[c] = fmaf(zb, __ldg(a.x + gp * d_in + c), part[c]);
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        const float zk = __ldcs(src + (size_t)((1 + k) * P + p) * HW);
#pragma unroll
        for (int c = 0; c < TPX_MAX_DIN; ++c) part[c] += (a.jet.dcol[k] == c) ? zk : 0.0f;
      }
    }
#pragma unroll
    for (int c = 0; c < TPX_MAX_DIN; ++c) acc[c] += (double)part[c];
  }
#pragma unroll
  for (int c = 0; c < TPX_MAX_DIN; ++c)
    if (c < d_in) atomicAdd(a.gacc + lay.w0() + (size_t)c * lay.HP + j, acc[c]);
}

template <int ND, int LAP>
struct Launch {
  static constexpr int S = 1 + ND + LAP, N = S * P;
  static size_t layer_smem() { return (size_t)2 * 2 * N * HW * sizeof(float) + 1024; }
  static size_t rev_smem() {
    constexpr int NS = 8 * S, NP = (NS + 15) / 16 * 16, KA = (NS + 31) / 32;
    return (size_t)((S <= 4 ? 2 : 1) * 2 * NP * HW + 4 * KA * HW * 32 + 2 * NS * HW) * sizeof(float) + 1024;   // Zn, 2 x At, raw ring
  }
  static size_t dw_smem() { return (size_t)2 * 2 * ((N + 31) / 32) * HW * 32 * sizeof(float) + 1024; }

  template <class K>
  static int set_smem(K kernel, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(TPX_E_CUDA, "%s: shared memory attribute (%zu bytes): %s", what, bytes, cudaGetErrorString(e));
    return 0;
  }
  static int check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TPX_E_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
    return 0;
  }
  // dW block (a.jb, a.kb) of matrix a.l: the per-CTA partial sums of the w_rev_kernel launch that just went out -> gradient image
  static int reduce_dw(const WArgs& a, unsigned nctas, cudaStream_t stream) {
    double* g = a.gacc + a.lay.w(a.l) + (size_t)(a.jb * HW) * a.lay.HP + a.kb * HW;
    w_dw_reduce_kernel<<<HW * HW / 256, 256, 0, stream>>>(a.dwpart, (int)nctas, g, a.lay.HP);
    return check("w_dw_reduce_kernel");
  }
  static unsigned grid(const WArgs& a, int sm_count) {
    const int64_t ntiles = (a.n + P - 1) / P;
    return (unsigned)(sm_count < ntiles ? sm_count : ntiles);
  }

  // Storage (floats): per = one half-array = ntiles x N x 128; a layer is nblk = HP / 128 consecutive half-arrays.
  //   [checkpoints of hidden layers 1..L-1][Zbar A][Zbar B][partial products (nblk = 2 only)]
  static size_t per_floats(int64_t n) { return (size_t)((n + P - 1) / P) * N * HW; }
  // the partial-product array exists for 256-wide layers and for more than one block of outputs
  static bool has_part(const Layout& lay) { return lay.HP > HW || (lay.wide_out() && lay.dop() > HW); }

  static int forward(WArgs a, float* saved, float* u, float* du, float* lapo, int sm_count, cudaStream_t stream) {
    const int L = a.lay.L, nblk = a.lay.HP / HW;
    const size_t per = per_floats(a.n), layer = (size_t)nblk * per;
    float* part = saved + (size_t)(L + 1) * layer;
    float* ck0 = part + (has_part(a.lay) ? per : 0);             // checkpoints of layer 0 (d_in > MAXD only)
    a.in0 = a.lay.d_in > MAXD;
    int rc = set_smem(w_layer_kernel<ND, LAP, FWD>, layer_smem(), "w_layer_kernel<FWD>");
    if (rc) return rc;
    if (a.in0) {
      const int64_t ntiles = (a.n + P - 1) / P;
      const int64_t blocks = ntiles < (int64_t)sm_count * 16 ? ntiles : (int64_t)sm_count * 16;
      for (int hb = 0; hb < nblk; ++hb) {
        a.kb = hb; a.out = ck0 + (size_t)hb * per;
        w_in_fwd_kernel<ND, LAP><<<(unsigned)blocks, HW, 0, stream>>>(a);
        if ((rc = check("w_in_fwd_kernel"))) return rc;
      }
    }
    for (int l = 1; l < L; ++l)
      for (int jb = 0; jb < nblk; ++jb)
        for (int kb = 0; kb < nblk; ++kb) {      // z_l[jb] = sum_kb W_l[jb][kb] a_{l-1}[kb]: the K blocks meet in `part`
          a.l = l; a.jb = jb; a.kb = kb;
          a.in = (l == 1) ? (a.in0 ? ck0 + (size_t)kb * per : nullptr) : saved + (size_t)(l - 2) * layer + (size_t)kb * per;
          a.out = saved + (size_t)(l - 1) * layer + (size_t)jb * per;
          a.part = part; a.part_out = kb < nblk - 1; a.part_in = kb > 0;
          w_layer_kernel<ND, LAP, FWD><<<grid(a, sm_count), THREADS, layer_smem(), stream>>>(a);
          if ((rc = check("w_layer_kernel<FWD>"))) return rc;
        }
    if (u && a.lay.wide_out()) {
      // raw products of matrix L, one block of 128 outputs at a time, then u / du / lap.  One 128-neuron block per layer: into
      // the (idle) Zbar array; two: the K blocks meet in the partial-product array, which the second launch updates in place
      float* raw = nblk == 1 ? saved + (size_t)(L - 1) * layer : part;
      const int nob = (int)(a.lay.dop() / HW);
      for (int ob = 0; ob < nob; ++ob) {
        for (int kb = 0; kb < nblk; ++kb) {
          a.l = L; a.jb = ob; a.kb = kb;
          a.in = saved + (size_t)(L - 2) * layer + (size_t)kb * per;
          a.out = nullptr; a.part = raw; a.part_out = 1; a.part_in = kb > 0;
          w_layer_kernel<ND, LAP, FWD><<<grid(a, sm_count), THREADS, layer_smem(), stream>>>(a);
          if ((rc = check("w_layer_kernel<FWD> (output layer)"))) return rc;
        }
        const int cols = a.lay.d_out - ob * HW < HW ? a.lay.d_out - ob * HW : HW;
        int64_t blocks = (a.n * cols + OUT_THREADS - 1) / OUT_THREADS;
        if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
        if (a.jet.out_major) {
          const int64_t ntiles = (a.n + P - 1) / P;
          const int64_t tb = ntiles < (int64_t)sm_count * 8 ? ntiles : (int64_t)sm_count * 8;
          w_out_gather_t_kernel<ND, LAP><<<(unsigned)tb, OUT_THREADS, 0, stream>>>(a, raw, u, du, lapo);
        } else {
          w_out_gather_kernel<ND, LAP><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, raw, u, du, lapo);
        }
        if ((rc = check("w_out_gather_kernel"))) return rc;
      }
      return 0;
    }
    if (u) {
      a.in = saved + (size_t)(L - 2) * layer;
      a.half_stride = (int64_t)per;
      int64_t blocks = (a.n + OUT_THREADS / 32 - 1) / (OUT_THREADS / 32);
      if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
      w_out_fwd_kernel<ND, LAP><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, u, du, lapo);
      if ((rc = check("w_out_fwd_kernel"))) return rc;
    }
    return 0;
  }

  static int backward(WArgs a, float* saved, const float* gu, const float* gdu, const float* glap, int sm_count, cudaStream_t stream) {
    const int L = a.lay.L, nblk = a.lay.HP / HW;
    const size_t per = per_floats(a.n), layer = (size_t)nblk * per;
    float* zb[2] = {saved + (size_t)(L - 1) * layer, saved + (size_t)L * layer};
    float* part = saved + (size_t)(L + 1) * layer;
    float* ck0 = part + (has_part(a.lay) ? per : 0);             // checkpoints of layer 0 (d_in > MAXD only)
    a.in0 = a.lay.d_in > MAXD;
    a.dwpart = ck0 + (a.in0 ? layer : 0);
    static const int split = env_flag("TPX_TCW_SPLIT");      // development: separate dW / adjoint launches (width <= 128)
    int rc;
    const bool wide_out = a.lay.wide_out();
    const int nob = wide_out ? (int)(a.lay.dop() / HW) : 0;      // 128-output blocks of the output layer
    for (int hb = 0; hb < nblk && !wide_out; ++hb) {
      a.kb = hb;
      a.in = saved + (size_t)(L - 2) * layer + (size_t)hb * per;
      a.out = zb[0] + (size_t)hb * per;
      const int64_t npad = (a.n + P - 1) / P * P;
      int64_t blocks = (npad + OUT_THREADS / 32 - 1) / (OUT_THREADS / 32);
      if (blocks > (int64_t)sm_count * 6) blocks = (int64_t)sm_count * 6;
      const int d_out = a.lay.d_out;
      if (d_out <= 1) w_out_bwd_kernel<ND, LAP, 1><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, gu, gdu, glap);
      else if (d_out <= 2) w_out_bwd_kernel<ND, LAP, 2><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, gu, gdu, glap);
      else if (d_out <= 4) w_out_bwd_kernel<ND, LAP, 4><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, gu, gdu, glap);
      else w_out_bwd_kernel<ND, LAP, 8><<<(unsigned)blocks, OUT_THREADS, 0, stream>>>(a, gu, gdu, glap);
      if ((rc = check("w_out_bwd_kernel"))) return rc;
    }
    if (split && nblk == 1 && !wide_out && !a.in0) {
      if ((rc = set_smem(w_layer_kernel<ND, LAP, ADJ>, layer_smem(), "w_layer_kernel<ADJ>"))) return rc;
      if ((rc = set_smem(w_dw_kernel<ND, LAP>, dw_smem(), "w_dw_kernel"))) return rc;
    } else {
      if ((rc = set_smem(w_rev_kernel<ND, LAP, false>, rev_smem(), "w_rev_kernel"))) return rc;
      if ((rc = set_smem(w_rev_kernel<ND, LAP, true>, rev_smem(), "w_rev_kernel"))) return rc;
    }
    int cur = 0;
    if (wide_out) {
      // matrix L block by block: Zbar_L[ob] = the output adjoints of the block (seed kernel); dW_L[ob] is complete per launch,
      // Abar_{L-1} = sum_ob W_L[ob]^T Zbar_L[ob] meets in `part`; the last launch finishes Zbar_{L-1} and db_{L-1}
      a.l = L;
      const int64_t npad = (a.n + P - 1) / P * P;
      const int64_t sblocks = npad < (int64_t)sm_count * 16 ? npad : (int64_t)sm_count * 16;
      const int64_t nsub = (a.n + P - 1) / P * 2;
      unsigned g = (unsigned)(sm_count < nsub ? sm_count : nsub);
      if (g > MAX_REV_CTAS) g = MAX_REV_CTAS;
      for (int ib = 0; ib < nblk; ++ib)            // input-neuron block of the last hidden layer
        for (int ob = 0; ob < nob; ++ob) {
          a.jb = ob; a.kb = ib;                      // (the seed kernel adds dbL for kb == 0 only)
          a.out = zb[0];
          if (a.jet.out_major) {
            const int64_t ntiles = (a.n + P - 1) / P;
            const int64_t tb = ntiles < (int64_t)sm_count * 8 ? ntiles : (int64_t)sm_count * 8;
            w_out_seed_t_kernel<ND, LAP><<<(unsigned)tb, OUT_THREADS, 0, stream>>>(a, gu, gdu, glap);
          } else {
            w_out_seed_kernel<ND, LAP><<<(unsigned)sblocks, HW, 0, stream>>>(a, gu, gdu, glap);
          }
          if ((rc = check("w_out_seed_kernel"))) return rc;
          a.in = zb[0];
          a.ck = saved + (size_t)(L - 2) * layer + (size_t)ib * per;
          a.out = zb[1] + (size_t)ib * per;
          a.part = part; a.part_out = ob < nob - 1; a.part_in = ob > 0;
          w_rev_kernel<ND, LAP, false><<<g, THREADS, rev_smem(), stream>>>(a);
          if ((rc = check("w_rev_kernel (output layer)"))) return rc;
          if ((rc = reduce_dw(a, g, stream))) return rc;
        }
      cur = 1;
    }
    for (int l = L - 1; l >= 1; --l) {
      a.l = l;
      if (split && nblk == 1 && !wide_out && !a.in0) {
        a.jb = a.kb = 0; a.part_out = a.part_in = 0;
        a.in = zb[cur];
        a.ck = (l == 1) ? nullptr : saved + (size_t)(l - 2) * layer;
        a.out = zb[cur ^ 1];
        w_dw_kernel<ND, LAP><<<grid(a, sm_count), THREADS, dw_smem(), stream>>>(a);
        if ((rc = check("w_dw_kernel"))) return rc;
        w_layer_kernel<ND, LAP, ADJ><<<grid(a, sm_count), THREADS, layer_smem(), stream>>>(a);
        if ((rc = check("w_layer_kernel<ADJ>"))) return rc;
      } else {
        // dW_l[jb][ib] is complete per launch; Abar_{l-1}[ib] = sum_jb W_l[jb][ib]^T Zbar_l[jb] meets in `part`
        for (int ib = 0; ib < nblk; ++ib)
          for (int jb = 0; jb < nblk; ++jb) {
            a.jb = jb; a.kb = ib;
            a.in = zb[cur] + (size_t)jb * per;
            a.ck = (l == 1) ? (a.in0 ? ck0 + (size_t)ib * per : nullptr) : saved + (size_t)(l - 2) * layer + (size_t)ib * per;
            a.out = zb[cur ^ 1] + (size_t)ib * per;
            a.part = part; a.part_out = jb < nblk - 1; a.part_in = jb > 0;
            const int64_t nsub = (a.n + P - 1) / P * 2;
            unsigned g = (unsigned)(sm_count < nsub ? sm_count : nsub);
            if (g > MAX_REV_CTAS) g = MAX_REV_CTAS;
            if (l == 1 && !a.in0) w_rev_kernel<ND, LAP, true><<<g, THREADS, rev_smem(), stream>>>(a);
            else w_rev_kernel<ND, LAP, false><<<g, THREADS, rev_smem(), stream>>>(a);
            if ((rc = check("w_rev_kernel"))) return rc;
            if ((rc = reduce_dw(a, g, stream))) return rc;
          }
      }
      cur ^= 1;
    }
    if (a.in0) {                                 // Zbar_0 (the last sweep's output) -> dW0
      const int64_t ntiles = (a.n + P - 1) / P;
      const int64_t blocks = ntiles < (int64_t)sm_count * 8 ? ntiles : (int64_t)sm_count * 8;
      for (int hb = 0; hb < nblk; ++hb) {
        a.kb = hb; a.in = zb[cur] + (size_t)hb * per;
        w_in_bwd_kernel<ND, LAP><<<(unsigned)blocks, HW, 0, stream>>>(a);
        if ((rc = check("w_in_bwd_kernel"))) return rc;
      }
    }
    return 0;
  }
};

}  // namespace TCW_NS

using namespace TCW_NS;

#ifdef TCW_ACT_SIN
#define tcw_forward tcws_forward
#define tcw_backward tcws_backward
#else
bool tcw_takes_narrow(const tpx_fcn_desc* net) {
  if (net->activation != TPX_ACT_TANH && net->activation != TPX_ACT_SIN) return false;
  if (net->n_hidden < 2 || net->d_in > TPX_MAX_DIN || net->d_out > MAXO_WIDE) return false;
  for (int l = 0; l < net->n_hidden; ++l)
    if (net->widths[l] > 64) return false;
  return !tc_net_fits(net);                       // more than 5 hidden layers, more than 8 outputs, or more than 4 inputs
}

bool tcw_supported(const tpx_fcn_desc* net, int nd, int lap) {
  static const int disabled = tcw::env_flag("TPX_DISABLE_TCW");
  if (disabled) return false;
  if (net->activation != TPX_ACT_TANH && net->activation != TPX_ACT_SIN) return false;
  if (net->n_hidden < 2) return false;
  int wmax = 0;
  for (int l = 0; l < net->n_hidden; ++l) wmax = net->widths[l] > wmax ? net->widths[l] : wmax;
  if (wmax > 2 * HW) return false;                   // 129..256: two 128-neuron blocks per layer
  if (wmax <= 64 && !tcw_takes_narrow(net)) return false;
  if (net->d_in > TPX_MAX_DIN || net->d_out > MAXO_WIDE) return false;   // 9 or more outputs: Layout::wide_out; 5..8 inputs: w_in_*_kernel
  if (lap && nd == 0) return false;
  if (nd < 0 || nd > 3) return false;
  return true;
}

// checkpoints of hidden layers 1..L-1 plus the two Zbar buffers of the reverse sweep (+ one half-array of partial
// products for 256-wide layers); a layer of width <= 128 is one half-array, a wider one two
size_t tcw_saved_bytes(const tpx_fcn_desc* net, int nd, int lap, int64_t n) {
  int wmax = 0;
  for (int l = 0; l < net->n_hidden; ++l) wmax = net->widths[l] > wmax ? net->widths[l] : wmax;
  const size_t nblk = wmax > HW ? 2 : 1;
  const size_t per = (size_t)((n + P - 1) / P) * (size_t)((1 + nd + lap) * P) * HW * sizeof(float);
  const bool out_chain = nblk == 1 && net->d_out > HW;       // more than one block of outputs: their adjoint products meet in `part`
  const size_t in0 = net->d_in > MAXD ? nblk : 0;            // checkpoints of layer 0 when it has kernels of its own
  return ((size_t)(net->n_hidden + 1) * nblk + ((nblk > 1 || out_chain) ? 1 : 0) + in0) * per + (size_t)MAX_REV_CTAS * HW * HW * sizeof(float);
}

#endif  // !TCW_ACT_SIN

#define TCW_DISPATCH(CALL)                                     \
  switch (jet.nd * 2 + lap) {                                  \
    case 0: return Launch<0, 0>::CALL;                         \
    case 2: return Launch<1, 0>::CALL;                         \
    case 3: return Launch<1, 1>::CALL;                         \
    case 4: return Launch<2, 0>::CALL;                         \
    case 5: return Launch<2, 1>::CALL;                         \
    case 6: return Launch<3, 0>::CALL;                         \
    case 7: return Launch<3, 1>::CALL;                         \
    default: return fail(TPX_E_UNSUPPORTED, "wide tensor-core path: nd=%d lap=%d", jet.nd, lap); \
  }

int tcw_forward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream) {
  WArgs a;
  memset(&a, 0, sizeof(a));
  a.lay = lay; a.jet = jet; a.packed = packed; a.x = x; a.n = n;
  TCW_DISPATCH(forward(a, saved, u, du, lapo, sm_count, stream))
}

int tcw_backward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                 const float* gu, const float* gdu, const float* glap, double* gacc, float* saved, int sm_count,
                 cudaStream_t stream) {
  WArgs a;
  memset(&a, 0, sizeof(a));
  a.lay = lay; a.jet = jet; a.packed = packed; a.x = x; a.n = n; a.gacc = gacc;
  a.dbg = g_tc_debug;
  TCW_DISPATCH(backward(a, saved, gu, gdu, glap, sm_count, stream))
}

}  // namespace tpx
