// SquaredError + mean of a residual batch (reference problem/conditions/condition.py:11-27 and 488-489:
// loss = mean over rows of the sum of squares over every other axis) and its adjoint, as two element-wise /
// reduction kernels: float4 grid-stride reads, warp-shuffle reduction, one float64 atomic per CTA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tpx.h"
#include "tpx_internal.h"

namespace tpx {

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS) sq_err_mean_kernel(const float* __restrict__ r, int64_t total, double inv_rows,
                                                                   double* __restrict__ acc) {
  float s = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * LOSS_THREADS;
  const int64_t t = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(r) & 15) == 0) {
    const int64_t n4 = total >> 2;
    const float4* r4 = reinterpret_cast<const float4*>(r);
    for (int64_t i = t; i < n4; i += stride) {
      const float4 v = __ldg(r4 + i);
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    for (int64_t i = (n4 << 2) + t; i < total; i += stride) s = fmaf(r[i], r[i], s);
  } else {
    for (int64_t i = t; i < total; i += stride) s = fmaf(r[i], r[i], s);
  }
  double d = (double)s;                          // per-thread fp32 partial (<= a few thousand terms), float64 from here on
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, m);
  __shared__ double part[LOSS_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x < 32) {
    d = (threadIdx.x < LOSS_THREADS / 32) ? part[threadIdx.x] : 0.0;
#pragma unroll
    for (int m = 4; m > 0; m >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, m);
    if (threadIdx.x == 0) atomicAdd(acc, d * inv_rows);
  }
}

// gr = (upstream scalar gradient) * 2 r / rows
__global__ void __launch_bounds__(LOSS_THREADS) sq_err_mean_bwd_kernel(const float* __restrict__ r, int64_t total, float coef,
                                                                       const float* __restrict__ gscale, float* __restrict__ gr) {
  const float c = coef * (gscale ? __ldg(gscale) : 1.0f);
  const int64_t stride = (int64_t)gridDim.x * LOSS_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < total; i += stride) gr[i] = c * r[i];
}

static unsigned loss_grid(int64_t total, int per_thread) {
  int64_t blocks = (total + (int64_t)LOSS_THREADS * per_thread - 1) / ((int64_t)LOSS_THREADS * per_thread);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return (unsigned)blocks;
}

}  // namespace tpx

using namespace tpx;

extern "C" {

int tpx_squared_error_mean(const float* r, int64_t total, int64_t rows, double* loss_acc, void* stream) {
  if (total < 0 || rows <= 0) return fail(TPX_E_ARG, "total=%lld rows=%lld", (long long)total, (long long)rows);
  if (total == 0) return 0;
  if (!r || !loss_acc) return fail(TPX_E_ARG, "NULL pointer");
  sq_err_mean_kernel<<<loss_grid(total, 16), LOSS_THREADS, 0, (cudaStream_t)stream>>>(r, total, 1.0 / (double)rows, loss_acc);
  if (cudaGetLastError() != cudaSuccess) return cuda_fail("sq_err_mean_kernel");
  return 0;
}

int tpx_squared_error_mean_backward(const float* r, int64_t total, int64_t rows, const float* grad_loss, float* grad_r,
                                    void* stream) {
  if (total < 0 || rows <= 0) return fail(TPX_E_ARG, "total=%lld rows=%lld", (long long)total, (long long)rows);
  if (total == 0) return 0;
  if (!r || !grad_r) return fail(TPX_E_ARG, "NULL pointer");
  sq_err_mean_bwd_kernel<<<loss_grid(total, 4), LOSS_THREADS, 0, (cudaStream_t)stream>>>(r, total, (float)(2.0 / (double)rows),
                                                                                         grad_loss, grad_r);
  if (cudaGetLastError() != cudaSuccess) return cuda_fail("sq_err_mean_bwd_kernel");
  return 0;
}

}  // extern "C"
