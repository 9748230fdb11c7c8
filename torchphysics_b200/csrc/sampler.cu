// Collocation-point sampler kernels: Philox4x32-10 counter-based uniform sampling of
// Interval / Parallelogram / Circle, membership of boolean regions (Cut / Union /
// Intersection as a postfix program over shapes), rejection sampling with warp-ballot
// stable compaction, deterministic grids, Latin-hypercube strata and the Union pick.
//
// Membership tests use the float32 operation order of the reference's `_contains`
// (domain1D/interval.py:37-43, domain2D/parallelogram.py:83-103, domain2D/circle.py:34-40)
// with explicit round-to-nearest intrinsics so that ptxas cannot contract them into FMAs:
// the produced points then satisfy the reference's own membership test bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tpx.h"
#include "tpx_internal.h"

namespace tpx {

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t sub) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = sub, c3 = 0;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// 24 random bits -> [0, 1), the mapping torch.rand uses for float32
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r & 0x00FFFFFFu) * 5.9604644775390625e-8f; }

// ---------------------------------------------------------------- shapes
// torch.isclose(a, b), rtol = 1e-5, atol = 1e-8, float32
__device__ __forceinline__ bool isclose32(float a, float b) {
  return fabsf(__fsub_rn(a, b)) <= __fadd_rn(1e-8f, __fmul_rn(1e-5f, fabsf(b)));
}
// torch.linalg.norm of a float32 pair as torch's CPU kernel evaluates it: sqrt(fma(y, y, x * x)) (pinned on the
// borderline candidates of tests/golden/cut_boundary.npz)
__device__ __forceinline__ float norm2(float x, float y) { return __fsqrt_rn(__fmaf_rn(y, y, __fmul_rn(x, x))); }
// barycentric coordinates of a point in a parallelogram (parallelogram.py:96-103, Cramer's rule)
__device__ __forceinline__ void bary(const tpx_shape& s, const float* q, float& bx, float& by) {
  const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
  const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
  const float px = __fsub_rn(q[0], s.p[0]), py = __fsub_rn(q[1], s.p[1]);
  const float det = __fsub_rn(__fmul_rn(d1x, d2y), __fmul_rn(d1y, d2x));
  bx = __fdiv_rn(__fsub_rn(__fmul_rn(d2y, px), __fmul_rn(d2x, py)), det);
  by = __fdiv_rn(__fsub_rn(__fmul_rn(d1x, py), __fmul_rn(d1y, px)), det);
}
// point at arc length `loc` along the four sides of a parallelogram (parallelogram.py:220-246)
__device__ __forceinline__ void para_walk(const tpx_shape& s, float loc, float* q) {
  const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
  const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
  const float s1 = norm2(d1x, d1y), s2 = norm2(d2x, d2y);
  float x = 0.0f, y = 0.0f;
  const float dx[4] = {d1x, d2x, -d1x, -d2x}, dy[4] = {d1y, d2y, -d1y, -d2y}, len[4] = {s1, s2, s1, s2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float sc = fminf(fmaxf(__fdiv_rn(loc, len[k]), 0.0f), 1.0f);
    x = __fadd_rn(x, __fmul_rn(sc, dx[k]));
    y = __fadd_rn(y, __fmul_rn(sc, dy[k]));
    loc = __fsub_rn(loc, len[k]);
  }
  q[0] = __fadd_rn(x, s.p[0]);
  q[1] = __fadd_rn(y, s.p[1]);
}
__device__ __forceinline__ float para_perimeter(const tpx_shape& s) {
  const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
  const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
  return __fmul_rn(2.0f, __fadd_rn(norm2(d1x, d1y), norm2(d2x, d2y)));
}
// element j of torch.linspace(0, 1, steps) in float32
__device__ __forceinline__ float lin01(int64_t j, int64_t steps) {
  const float step = 1.0f / (float)(steps - 1);
  return (j < steps / 2) ? (float)j * step : 1.0f - step * (float)(steps - 1 - j);
}

__device__ __forceinline__ bool shape_contains(const tpx_shape& s, const float* row) {
  const float* q = row + s.col;
  switch (s.kind) {
    case TPX_SHAPE_INTERVAL:
      return q[0] >= s.p[0] && q[0] <= s.p[1];
    case TPX_SHAPE_PARALLELOGRAM: {
      const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
      const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
      const float px = __fsub_rn(q[0], s.p[0]), py = __fsub_rn(q[1], s.p[1]);
      const float det = __fsub_rn(__fmul_rn(d1x, d2y), __fmul_rn(d1y, d2x));
      const float xd = __fsub_rn(__fmul_rn(d2y, px), __fmul_rn(d2x, py));
      const float yd = __fsub_rn(__fmul_rn(d1x, py), __fmul_rn(d1y, px));
      const float bx = __fdiv_rn(xd, det), by = __fdiv_rn(yd, det);
      return bx >= 0.0f && bx <= 1.0f && by >= 0.0f && by <= 1.0f;
    }
    case TPX_SHAPE_CIRCLE: {
      const float dx = __fsub_rn(q[0], s.p[0]), dy = __fsub_rn(q[1], s.p[1]);
      return norm2(dx, dy) <= s.p[2];
    }
    case TPX_SHAPE_INTERVAL_BOUNDARY:
      return isclose32(q[0], s.p[0]) || isclose32(q[0], s.p[1]);
    case TPX_SHAPE_INTERVAL_POINT:
      return isclose32(q[0], s.p[0]);
    case TPX_SHAPE_PARALLELOGRAM_BOUNDARY: {
      float bx, by;
      bary(s, q, bx, by);
      const bool xe = (isclose32(bx, 0.0f) || isclose32(bx, 1.0f)) && by >= 0.0f && by <= 1.0f;
      const bool ye = (isclose32(by, 0.0f) || isclose32(by, 1.0f)) && bx >= 0.0f && bx <= 1.0f;
      return xe || ye;
    }
    case TPX_SHAPE_CIRCLE_BOUNDARY: {
      const float dx = __fsub_rn(q[0], s.p[0]), dy = __fsub_rn(q[1], s.p[1]);
      return isclose32(norm2(dx, dy), s.p[2]);
    }
    default:
      return false;
  }
}

__device__ __forceinline__ bool region_contains(const tpx_region& rg, const float* row) {
  uint32_t stack = 0;  // bit stack, top = bit 0
  for (int i = 0; i < rg.n_ops; ++i) {
    const int op = rg.ops[i] & 0xFF, arg = rg.ops[i] >> 8;
    switch (op) {
      case TPX_OP_SHAPE: stack = (stack << 1) | (shape_contains(rg.shapes[arg], row) ? 1u : 0u); break;
      case TPX_OP_NOT: stack ^= 1u; break;
      case TPX_OP_AND: { const uint32_t b = stack & 1u; stack >>= 1; stack = (stack & ~1u) | ((stack & 1u) & b); break; }
      case TPX_OP_OR: { const uint32_t b = stack & 1u; stack >>= 1; stack |= b; break; }
      default: break;
    }
  }
  return (stack & 1u) != 0;
}

__device__ __host__ __forceinline__ int shape_dim(int kind) {
  return (kind == TPX_SHAPE_INTERVAL || kind == TPX_SHAPE_INTERVAL_BOUNDARY || kind == TPX_SHAPE_INTERVAL_POINT) ? 1 : 2;
}

// uniform sample of a shape from one Philox block (reference sample_random_uniform:
// interval.py:45-55, parallelogram.py:105-117, circle.py:52-68)
__device__ __forceinline__ void shape_sample(const tpx_shape& s, uint4 r, float* q) {
  switch (s.kind) {
    case TPX_SHAPE_INTERVAL:
      q[0] = __fadd_rn(__fmul_rn(u01(r.x), __fsub_rn(s.p[1], s.p[0])), s.p[0]);
      break;
    case TPX_SHAPE_PARALLELOGRAM: {
      const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
      const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
      const float b0 = u01(r.x), b1 = u01(r.y);
      q[0] = __fadd_rn(__fadd_rn(__fmul_rn(b0, d1x), __fmul_rn(b1, d2x)), s.p[0]);
      q[1] = __fadd_rn(__fadd_rn(__fmul_rn(b0, d1y), __fmul_rn(b1, d2y)), s.p[1]);
      break;
    }
    case TPX_SHAPE_CIRCLE: {
      const float rr = __fmul_rn(__fsqrt_rn(u01(r.x)), s.p[2]);
      const float phi = __fmul_rn(6.283185307179586f, u01(r.y));
      float sn, cs;
      sincosf(phi, &sn, &cs);
      q[0] = __fadd_rn(__fmul_rn(rr, cs), s.p[0]);
      q[1] = __fadd_rn(__fmul_rn(rr, sn), s.p[1]);
      break;
    }
    case TPX_SHAPE_INTERVAL_BOUNDARY:       // interval.py:115-126
      q[0] = u01(r.x) < 0.5f ? s.p[0] : s.p[1];
      break;
    case TPX_SHAPE_INTERVAL_POINT:          // interval.py:167-176
      q[0] = s.p[0];
      break;
    case TPX_SHAPE_PARALLELOGRAM_BOUNDARY:  // parallelogram.py:183-200
      para_walk(s, __fmul_rn(u01(r.x), para_perimeter(s)), q);
      break;
    case TPX_SHAPE_CIRCLE_BOUNDARY: {       // circle.py:126-141
      const float phi = __fmul_rn(6.283185307179586f, u01(r.x));
      float sn, cs;
      sincosf(phi, &sn, &cs);
      q[0] = __fadd_rn(__fmul_rn(s.p[2], cs), s.p[0]);
      q[1] = __fadd_rn(__fmul_rn(s.p[2], sn), s.p[1]);
      break;
    }
    default: break;
  }
}

// ---------------------------------------------------------------- kernels
__global__ void contains_kernel(tpx_region rg, const float* __restrict__ pts, int64_t n, int stride,
                                uint8_t* __restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = region_contains(rg, pts + i * stride) ? 1 : 0;
}

__global__ void sample_kernel(tpx_shape s, uint64_t seed, uint64_t offset, int64_t n, float* __restrict__ out, int stride) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float q[2];
    shape_sample(s, ph(offset + (uint64_t)i, 0), q);
    float* row = out + i * stride + s.col;
    row[0] = q[0];
    if (shape_dim(s.kind) > 1) row[1] = q[1];
  }
}

// the same with the Philox key in device memory (a CUDA-graph replay cannot receive a fresh key as a launch argument)
__global__ void sample_dkey_kernel(tpx_shape s, const uint64_t* __restrict__ key, uint64_t offset, int64_t n, float* __restrict__ out,
                                   int stride) {
  const Philox ph(*key);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float q[2];
    shape_sample(s, ph(offset + (uint64_t)i, 0), q);
    float* row = out + i * stride + s.col;
    row[0] = q[0];
    if (shape_dim(s.kind) > 1) row[1] = q[1];
  }
}
// key <- splitmix64(key): one fresh key per sampling call of a captured step
__global__ void key_advance_kernel(uint64_t* key) {
  uint64_t x = *key + 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  *key = x ^ (x >> 31);
}

// Rejection sampling = three launches, no host round trip:
//   count   : every CTA generates its CPB candidates, tests them, stores #accepted
//   scan    : one CTA turns the per-CTA counts into exclusive offsets (+ running total)
//   scatter : candidates are regenerated (Philox is counter based, so this is free of
//             state) and the accepted ones are written in candidate order -- a STABLE
//             compaction, so the result is a deterministic function of (seed, offset).
constexpr int RJ_THREADS = 256;
constexpr int RJ_PER_THREAD = 8;
constexpr int RJ_CPB = RJ_THREADS * RJ_PER_THREAD;

__device__ __forceinline__ bool candidate(const tpx_shape& gen, const tpx_region& acc, const Philox& ph,
                                          uint64_t ctr, int d_total, float* row) {
  (void)d_total;
  row[0] = row[1] = row[2] = row[3] = 0.0f;
  shape_sample(gen, ph(ctr, 0), row + gen.col);
  return region_contains(acc, row);
}

__global__ void __launch_bounds__(RJ_THREADS) reject_count_kernel(tpx_shape gen, tpx_region acc, uint64_t seed,
                                                                   uint64_t offset, int64_t n_cand,
                                                                   uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t warp_cnt[RJ_THREADS / 32];
  const Philox ph(seed);
  const int64_t base = (int64_t)blockIdx.x * RJ_CPB;
  uint32_t cnt = 0;
  float row[4];
#pragma unroll 1
  for (int k = 0; k < RJ_PER_THREAD; ++k) {
    const int64_t i = base + k * RJ_THREADS + threadIdx.x;
    const bool ok = i < n_cand && candidate(gen, acc, ph, offset + (uint64_t)i, shape_dim(gen.kind) + gen.col, row);
    cnt += __popc(__ballot_sync(0xFFFFFFFFu, ok));
  }
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < RJ_THREADS / 32; ++w) t += warp_cnt[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of n_blocks counts (in place, 64-bit offsets written to offs), total
// appended to counters[0] (accepted so far, clamped to n_wanted) ; counters[1] = offered.
__global__ void __launch_bounds__(1024) reject_scan_kernel(const uint32_t* __restrict__ block_counts, int n_blocks,
                                                            int64_t* __restrict__ offs, int64_t* __restrict__ counters,
                                                            int64_t n_wanted, int64_t n_cand) {
  __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int per = (n_blocks + 1023) / 1024;
  const int lo = t * per, hi = min(lo + per, n_blocks);
  int64_t s = 0;
  for (int i = lo; i < hi; ++i) s += block_counts[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    int64_t v = (t >= d) ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int64_t run = counters[0] + ((t > 0) ? part[t - 1] : 0);
  for (int i = lo; i < hi; ++i) {
    offs[i] = run;
    run += block_counts[i];
  }
  __syncthreads();
  if (t == 1023) {
    const int64_t total = counters[0] + part[1023];
    counters[2] = counters[0];           // base of this round (read by nobody after this launch)
    counters[0] = total < n_wanted ? total : n_wanted;
    counters[1] += n_cand;
    counters[3] += part[1023];           // accepted candidates over all rounds (acceptance estimate)
  }
}

__global__ void __launch_bounds__(RJ_THREADS) reject_scatter_kernel(tpx_shape gen, tpx_region acc, uint64_t seed,
                                                                     uint64_t offset, int64_t n_cand,
                                                                     const int64_t* __restrict__ offs, int64_t n_wanted,
                                                                     float* __restrict__ out, int stride) {
  __shared__ uint32_t warp_base[RJ_PER_THREAD][RJ_THREADS / 32];
  const Philox ph(seed);
  const int64_t base = (int64_t)blockIdx.x * RJ_CPB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = shape_dim(gen.kind);
  float q[RJ_PER_THREAD][2];
  uint32_t okbits = 0, rank[RJ_PER_THREAD];
#pragma unroll
  for (int k = 0; k < RJ_PER_THREAD; ++k) {
    const int64_t i = base + k * RJ_THREADS + threadIdx.x;
    float row[4];
    const bool ok = i < n_cand && candidate(gen, acc, ph, offset + (uint64_t)i, d + gen.col, row);
    q[k][0] = row[gen.col];
    q[k][1] = d > 1 ? row[gen.col + 1] : 0.0f;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
    rank[k] = __popc(bal & ((1u << lane) - 1u));
    if (ok) okbits |= 1u << k;
    if (lane == 0) warp_base[k][warp] = __popc(bal);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive scan over (k, warp) in candidate order
    uint32_t run = 0;
    for (int k = 0; k < RJ_PER_THREAD; ++k)
      for (int w = 0; w < RJ_THREADS / 32; ++w) {
        const uint32_t c = warp_base[k][w];
        warp_base[k][w] = run;
        run += c;
      }
  }
  __syncthreads();
  const int64_t blk = offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < RJ_PER_THREAD; ++k) {
    if (!((okbits >> k) & 1u)) continue;
    const int64_t pos = blk + warp_base[k][warp] + rank[k];
    if (pos >= n_wanted) continue;
    float* row = out + pos * stride + gen.col;
    row[0] = q[k][0];
    if (d > 1) row[1] = q[k][1];
  }
}

// deterministic grids (interval.py:57-65, parallelogram.py:119-152, circle.py:70-99)
__global__ void grid_kernel(tpx_shape s, int64_t n, int n1, int n2, float* __restrict__ out, int stride) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float* row = out + i * stride + s.col;
    if (s.kind == TPX_SHAPE_INTERVAL) {
      // torch.linspace(0, 1, n+2)[1:-1]: step = 1/(n+1); element j<half: j*step, else 1-(steps-1-j)*step
      const int64_t steps = n + 2, j = i + 1;
      const float step = 1.0f / (float)(steps - 1);
      const float t = (j < steps / 2) ? (float)j * step : 1.0f - step * (float)(steps - 1 - j);
      row[0] = __fadd_rn(__fmul_rn(__fsub_rn(s.p[1], s.p[0]), t), s.p[0]);
    } else if (s.kind == TPX_SHAPE_PARALLELOGRAM) {
      // meshgrid(x (n1), y (n2)) permuted so that x runs fastest: index i -> (ix = i % n1, iy = i / n1)
      const int ix = (int)(i % n1), iy = (int)(i / n1);
      auto lin = [](int j, int steps) {
        const float step = 1.0f / (float)(steps - 1);
        return (j < steps / 2) ? (float)j * step : 1.0f - step * (float)(steps - 1 - j);
      };
      const float b0 = lin(ix + 1, n1 + 2), b1 = lin(iy + 1, n2 + 2);
      const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
      const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
      row[0] = __fadd_rn(__fadd_rn(__fmul_rn(b0, d1x), __fmul_rn(b1, d2x)), s.p[0]);
      row[1] = __fadd_rn(__fadd_rn(__fmul_rn(b0, d1y), __fmul_rn(b1, d2y)), s.p[1]);
    } else if (s.kind == TPX_SHAPE_INTERVAL_BOUNDARY) {
      row[0] = (i & 1) ? s.p[0] : s.p[1];                    // interval.py:128-137: upper, lower, upper, ...
    } else if (s.kind == TPX_SHAPE_INTERVAL_POINT) {
      row[0] = s.p[0];
    } else if (s.kind == TPX_SHAPE_PARALLELOGRAM_BOUNDARY) {
      para_walk(s, __fmul_rn(lin01(i, n + 1), para_perimeter(s)), row);   // parallelogram.py:202-218
    } else if (s.kind == TPX_SHAPE_CIRCLE_BOUNDARY) {
      // torch.linspace(0, 2 pi, n + 1)[:-1] (circle.py:143-160)
      const int64_t steps = n + 1;
      const float end = 6.283185307179586f, step = end / (float)(steps - 1);
      const float phi = (i < steps / 2) ? (float)i * step : end - step * (float)(steps - 1 - i);
      float sn, cs;
      sincosf(phi, &sn, &cs);
      row[0] = __fadd_rn(__fmul_rn(s.p[2], cs), s.p[0]);
      row[1] = __fadd_rn(__fmul_rn(s.p[2], sn), s.p[1]);
    } else {
      // sunflower seed arrangement, float64 angle reduced before the float32 sin/cos
      const double gr = (2.23606797749979 + 1.0) / 2.0;
      const double k = (double)(i + 1);
      const float phi = __fmul_rn((float)(2.0 * 3.141592653589793 / gr), (float)k);  // float32 product, as torch does
      const float rad = sqrtf((float)(k - 0.5)) / (float)sqrt((double)n + 0.5);
      float sn, cs;
      sincosf(phi, &sn, &cs);
      row[0] = __fadd_rn(__fmul_rn(s.p[2], __fmul_rn(rad, cs)), s.p[0]);
      row[1] = __fadd_rn(__fmul_rn(s.p[2], __fmul_rn(rad, sn)), s.p[1]);
    }
  }
}

// Latin hypercube strata (random_samplers.py:203-221): column c of row i receives
// lo + len * (perm[i] + rand) / n  where perm is a permutation supplied per axis.
__global__ void lhs_kernel(int d, const float* __restrict__ box, const int64_t* __restrict__ perm, uint64_t seed,
                           uint64_t offset, int64_t n, float* __restrict__ out, int stride, int col) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = ph(offset + (uint64_t)i, 1);
    const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
    for (int c = 0; c < d; ++c) {
      const float lo = box[2 * c], len = __fsub_rn(box[2 * c + 1], lo);
      const int64_t k = perm[(int64_t)c * n + i];
      const float cell = __fadd_rn(lo, __fmul_rn(len, (float)k / (float)n));
      out[i * stride + col + c] = __fadd_rn(cell, __fmul_rn(__fdiv_rn(len, (float)n), u01(rv[c & 3])));
    }
  }
}

// random sort keys for the LHS permutations (argsort of iid keys = uniform permutation)
__global__ void keys_kernel(uint64_t seed, uint64_t offset, int64_t n, int64_t* __restrict__ keys) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = ph(offset + (uint64_t)i, 2);
    keys[i] = (int64_t)((((uint64_t)r.x << 32) | r.y) >> 1);
  }
}

// Union pick (union.py:72-93): row i = a[i] if (b[i] in A) or rand <= vol_A/(vol_A+vol_B) else b[i]
__global__ void union_pick_kernel(tpx_region in_a, const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                  int d, int stride, float ratio, uint64_t seed, uint64_t offset, float* __restrict__ out) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool pick_a = region_contains(in_a, b + i * stride) || u01(ph(offset + (uint64_t)i, 3).x) <= ratio;
    const float* src = (pick_a ? a : b) + i * stride;
    for (int c = 0; c < d; ++c) out[i * stride + c] = src[c];
  }
}

// outward unit normals (interval.py:139-144,181-186, parallelogram.py:248-286, circle.py:162-170)
__global__ void normal_kernel(tpx_shape s, const float* __restrict__ pts, int64_t n, int stride, float* __restrict__ out) {
  const int d = shape_dim(s.kind);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* q = pts + i * stride + s.col;
    float* o = out + i * d;
    if (s.kind == TPX_SHAPE_INTERVAL_BOUNDARY) {
      o[0] = isclose32(q[0], s.p[0]) ? -1.0f : 1.0f;
    } else if (s.kind == TPX_SHAPE_INTERVAL_POINT) {
      o[0] = s.p[1];
    } else if (s.kind == TPX_SHAPE_CIRCLE_BOUNDARY) {
      o[0] = __fdiv_rn(__fsub_rn(q[0], s.p[0]), s.p[2]);
      o[1] = __fdiv_rn(__fsub_rn(q[1], s.p[1]), s.p[2]);
    } else {
      float bx, by;
      bary(s, q, bx, by);
      const float d1x = __fsub_rn(s.p[2], s.p[0]), d1y = __fsub_rn(s.p[3], s.p[1]);
      const float d2x = __fsub_rn(s.p[4], s.p[0]), d2y = __fsub_rn(s.p[5], s.p[1]);
      // unit normals of the two directions: (-dy, dx) / |d|, the second one negated
      const float l1 = norm2(-d1y, d1x), l2 = norm2(-d2y, d2x);
      const float n1x = __fdiv_rn(-d1y, l1), n1y = __fdiv_rn(d1x, l1);
      const float n2x = -__fdiv_rn(-d2y, l2), n2y = -__fdiv_rn(d2x, l2);
      float x = 0.0f, y = 0.0f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float lvl = (float)k, sgn = 2.0f * lvl - 1.0f;
        const float wy = isclose32(by, lvl) ? sgn : 0.0f, wx = isclose32(bx, lvl) ? sgn : 0.0f;
        x = __fadd_rn(x, __fmul_rn(n1x, wy)); y = __fadd_rn(y, __fmul_rn(n1y, wy));
        x = __fadd_rn(x, __fmul_rn(n2x, wx)); y = __fadd_rn(y, __fmul_rn(n2y, wx));
      }
      const float nrm = norm2(x, y);
      o[0] = __fdiv_rn(x, nrm);
      o[1] = __fdiv_rn(y, nrm);
    }
  }
}

static int check_shape(const tpx_shape* s, int stride) {
  if (!s) return fail(TPX_E_ARG, "shape is NULL");
  if (s->kind < TPX_SHAPE_INTERVAL || s->kind > TPX_SHAPE_CIRCLE_BOUNDARY) return fail(TPX_E_UNSUPPORTED, "shape kind %d", s->kind);
  const int d = shape_dim(s->kind);
  if (s->col < 0 || s->col + d > stride || s->col + d > 4) return fail(TPX_E_ARG, "shape columns %d..%d outside row of %d (max 4)", s->col, s->col + d, stride);
  return 0;
}
static int check_region(const tpx_region* r, int stride) {
  if (!r) return fail(TPX_E_ARG, "region is NULL");
  if (r->n_shapes < 1 || r->n_shapes > TPX_MAX_SHAPES || r->n_ops < 1 || r->n_ops > TPX_MAX_OPS)
    return fail(TPX_E_ARG, "region with %d shapes / %d ops", r->n_shapes, r->n_ops);
  int depth = 0;
  for (int i = 0; i < r->n_ops; ++i) {
    const int op = r->ops[i] & 0xFF, arg = r->ops[i] >> 8;
    if (op == TPX_OP_SHAPE) {
      if (arg < 0 || arg >= r->n_shapes) return fail(TPX_E_ARG, "op %d: shape index %d", i, arg);
      int rc = check_shape(&r->shapes[arg], stride);
      if (rc) return rc;
      ++depth;
    } else if (op == TPX_OP_NOT) {
      if (depth < 1) return fail(TPX_E_ARG, "op %d: NOT on empty stack", i);
    } else if (op == TPX_OP_AND || op == TPX_OP_OR) {
      if (depth < 2) return fail(TPX_E_ARG, "op %d: binary op needs two operands", i);
      --depth;
    } else {
      return fail(TPX_E_ARG, "op %d: unknown opcode %d", i, op);
    }
    if (depth > 30) return fail(TPX_E_UNSUPPORTED, "region program too deep");
  }
  if (depth != 1) return fail(TPX_E_ARG, "region program leaves %d values", depth);
  return 0;
}

static int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace tpx

using namespace tpx;

extern "C" {

int tpx_region_contains(const tpx_region* region, const float* points, int64_t n, int32_t row_stride,
                        uint8_t* mask, void* stream) {
  int rc = check_region(region, row_stride);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!points || !mask) return fail(TPX_E_ARG, "NULL pointer");
  contains_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*region, points, n, row_stride, mask);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("contains_kernel");
}

int tpx_sample_uniform(const tpx_shape* shape, uint64_t seed, uint64_t offset, int64_t n, float* out,
                       int32_t row_stride, void* stream) {
  int rc = check_shape(shape, row_stride);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!out) return fail(TPX_E_ARG, "NULL pointer");
  sample_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*shape, seed, offset, n, out, row_stride);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("sample_kernel");
}

int tpx_key_advance(uint64_t* key_dev, void* stream) {
  if (!key_dev) return fail(TPX_E_ARG, "NULL pointer");
  key_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(key_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("key_advance_kernel");
}

int tpx_sample_uniform_dkey(const tpx_shape* shape, const uint64_t* key_dev, uint64_t offset, int64_t n, float* out,
                            int32_t row_stride, void* stream) {
  int rc = check_shape(shape, row_stride);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!out || !key_dev) return fail(TPX_E_ARG, "NULL pointer");
  sample_dkey_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*shape, key_dev, offset, n, out, row_stride);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("sample_dkey_kernel");
}

int tpx_sample_reject_workspace_bytes(int64_t n_candidates, size_t* bytes) {
  if (!bytes || n_candidates < 0) return fail(TPX_E_ARG, "bad argument");
  const int64_t nb = (n_candidates + RJ_CPB - 1) / RJ_CPB;
  *bytes = (size_t)nb * (sizeof(int64_t) + sizeof(uint32_t)) + 64;
  return 0;
}

int tpx_sample_uniform_reject(const tpx_shape* generator, const tpx_region* accept, uint64_t seed,
                              uint64_t offset, int64_t n_candidates, int64_t n_wanted, float* out,
                              int32_t row_stride, int64_t* counters, void* workspace,
                              size_t workspace_bytes, void* stream) {
  int rc = check_shape(generator, row_stride);
  if (rc) return rc;
  rc = check_region(accept, row_stride);
  if (rc) return rc;
  if (n_candidates < 0 || n_wanted < 0) return fail(TPX_E_ARG, "negative count");
  if (n_candidates == 0) return 0;
  if (!out || !counters || !workspace) return fail(TPX_E_ARG, "NULL pointer");
  const int64_t nb = (n_candidates + RJ_CPB - 1) / RJ_CPB;
  if (nb > 0x7FFFFFFF) return fail(TPX_E_UNSUPPORTED, "too many candidates in one round");
  size_t need;
  tpx_sample_reject_workspace_bytes(n_candidates, &need);
  if (workspace_bytes < need) return fail(TPX_E_WORKSPACE, "workspace of %zu bytes < %zu", workspace_bytes, need);
  int64_t* offs = (int64_t*)workspace;
  uint32_t* counts = (uint32_t*)(offs + nb);
  cudaStream_t st = (cudaStream_t)stream;
  reject_count_kernel<<<(unsigned)nb, RJ_THREADS, 0, st>>>(*generator, *accept, seed, offset, n_candidates, counts);
  reject_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)nb, offs, counters, n_wanted, n_candidates);
  reject_scatter_kernel<<<(unsigned)nb, RJ_THREADS, 0, st>>>(*generator, *accept, seed, offset, n_candidates, offs, n_wanted, out, row_stride);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("reject kernels");
}

int tpx_sample_grid(const tpx_shape* shape, int64_t n, int32_t n1, int32_t n2, float* out, int32_t row_stride,
                    void* stream) {
  int rc = check_shape(shape, row_stride);
  if (rc) return rc;
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!out) return fail(TPX_E_ARG, "NULL pointer");
  if (shape->kind == TPX_SHAPE_PARALLELOGRAM && ((int64_t)n1 * n2 != n || n1 < 1)) return fail(TPX_E_ARG, "grid %d x %d != %lld", n1, n2, (long long)n);
  grid_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*shape, n, n1, n2, out, row_stride);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("grid_kernel");
}

int tpx_boundary_normal(const tpx_shape* shape, const float* points, int64_t n, int32_t row_stride,
                        float* normals, void* stream) {
  int rc = check_shape(shape, row_stride);
  if (rc) return rc;
  if (shape->kind < TPX_SHAPE_INTERVAL_BOUNDARY) return fail(TPX_E_ARG, "shape kind %d is not a boundary", shape->kind);
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!points || !normals) return fail(TPX_E_ARG, "NULL pointer");
  normal_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*shape, points, n, row_stride, normals);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("normal_kernel");
}

int tpx_sample_keys(uint64_t seed, uint64_t offset, int64_t n, int64_t* keys, void* stream) {
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!keys) return fail(TPX_E_ARG, "NULL pointer");
  keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, n, keys);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("keys_kernel");
}

int tpx_sample_lhs(int32_t d, const float* box, const int64_t* perm, uint64_t seed, uint64_t offset, int64_t n,
                   float* out, int32_t row_stride, int32_t col, void* stream) {
  if (d < 1 || d > 4 || col < 0 || col + d > row_stride) return fail(TPX_E_ARG, "d=%d col=%d stride=%d", d, col, row_stride);
  if (n < 0) return fail(TPX_E_ARG, "n < 0");
  if (n == 0) return 0;
  if (!box || !perm || !out) return fail(TPX_E_ARG, "NULL pointer");
  lhs_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d, box, perm, seed, offset, n, out, row_stride, col);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("lhs_kernel");
}

int tpx_union_pick(const tpx_region* in_a, const float* a, const float* b, int64_t n, int32_t d, int32_t row_stride,
                   float ratio, uint64_t seed, uint64_t offset, float* out, void* stream) {
  int rc = check_region(in_a, row_stride);
  if (rc) return rc;
  if (n < 0 || d < 1 || d > row_stride) return fail(TPX_E_ARG, "bad sizes");
  if (n == 0) return 0;
  if (!a || !b || !out) return fail(TPX_E_ARG, "NULL pointer");
  union_pick_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*in_a, a, b, n, d, row_stride, ratio, seed, offset, out);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("union_pick_kernel");
}

}  // extern "C"
