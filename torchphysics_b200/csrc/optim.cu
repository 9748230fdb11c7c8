// Multi-tensor Adam step in ONE launch: the reference leaves the update to torch.optim.Adam
// behind Lightning (solver.py:111-136), i.e. 4-5 elementwise launches per parameter tensor and step; a 4x64
// PINN has 10 tensors of 64..4096 elements, so the update is pure launch latency.  Same arithmetic as
// torch.optim.Adam (torch/optim/adam.py _single_tensor_adam, amsgrad = False, maximize = False):
//   g += wd * p;  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// A tensor whose gradient pointer is NULL is skipped, as torch skips parameters with .grad = None.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tpx.h"
#include "tpx_internal.h"

namespace tpx {

constexpr int AD_THREADS = 256;
constexpr int AD_PER_BLOCK = AD_THREADS * 4;

__global__ void __launch_bounds__(AD_THREADS) adam_kernel(tpx_tensor_list tl, float step_size, float bc2_sqrt, float w1,
                                                         float beta2, float w2, float eps, float weight_decay) {
  const int t = blockIdx.y;
  const float* __restrict__ g = (const float*)tl.grad[t];
  if (g == nullptr) return;
  float* __restrict__ p = (float*)tl.param[t];
  float* __restrict__ m = (float*)tl.exp_avg[t];
  float* __restrict__ v = (float*)tl.exp_avg_sq[t];
  const int64_t n = tl.numel[t];
  for (int64_t i = (int64_t)blockIdx.x * AD_PER_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * AD_PER_BLOCK) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = i + k * AD_THREADS;
      if (j < n) {
        const float pj = p[j];
        float gj = g[j];
        if (weight_decay != 0.0f) gj = __fadd_rn(gj, __fmul_rn(weight_decay, pj));
        const float mj = __fadd_rn(m[j], __fmul_rn(w1, __fsub_rn(gj, m[j])));            // lerp, weight < 0.5
        const float vj = __fadd_rn(__fmul_rn(beta2, v[j]), __fmul_rn(w2, __fmul_rn(gj, gj)));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vj), bc2_sqrt), eps);
        m[j] = mj;
        v[j] = vj;
        p[j] = __fadd_rn(pj, __fmul_rn(-step_size, __fdiv_rn(mj, denom)));
      }
    }
  }
}

// the same update with the step count in device memory (CUDA-graph capture): bias corrections are computed on the device
__global__ void __launch_bounds__(AD_THREADS) adam_dstep_kernel(tpx_tensor_list tl, float lr, float beta1, float beta2, float eps,
                                                               float weight_decay, const int32_t* __restrict__ step_dev) {
  const int t = blockIdx.y;
  const float* __restrict__ g = (const float*)tl.grad[t];
  if (g == nullptr) return;
  const double step = (double)*step_dev;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
  float* __restrict__ p = (float*)tl.param[t];
  float* __restrict__ m = (float*)tl.exp_avg[t];
  float* __restrict__ v = (float*)tl.exp_avg_sq[t];
  const int64_t n = tl.numel[t];
  for (int64_t i = (int64_t)blockIdx.x * AD_PER_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * AD_PER_BLOCK) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = i + k * AD_THREADS;
      if (j < n) {
        const float pj = p[j];
        float gj = g[j];
        if (weight_decay != 0.0f) gj = __fadd_rn(gj, __fmul_rn(weight_decay, pj));
        const float mj = __fadd_rn(m[j], __fmul_rn(w1, __fsub_rn(gj, m[j])));
        const float vj = __fadd_rn(__fmul_rn(beta2, v[j]), __fmul_rn(w2, __fmul_rn(gj, gj)));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vj), bc2_sqrt), eps);
        m[j] = mj;
        v[j] = vj;
        p[j] = __fadd_rn(pj, __fmul_rn(-step_size, __fdiv_rn(mj, denom)));
      }
    }
  }
}
__global__ void step_increment_kernel(int32_t* step_dev) { *step_dev += 1; }

}  // namespace tpx

using namespace tpx;

extern "C" int tpx_adam_step_dstep(const tpx_tensor_list* list, float lr, float beta1, float beta2, float eps, float weight_decay,
                                   int32_t* step_dev, void* stream) {
  if (!list || !step_dev) return fail(TPX_E_ARG, "NULL pointer");
  if (list->n < 0 || list->n > TPX_MAX_TENSORS) return fail(TPX_E_ARG, "%d tensors (max %d)", list->n, TPX_MAX_TENSORS);
  if (!(beta1 >= 0.5f && beta1 < 1.0f) || !(beta2 >= 0.0f && beta2 < 1.0f))
    return fail(TPX_E_UNSUPPORTED, "betas (%g, %g): beta1 in [0.5, 1), beta2 in [0, 1) supported", beta1, beta2);
  if (list->n == 0) return 0;
  int64_t nmax = 0;
  for (int t = 0; t < list->n; ++t) {
    if (list->numel[t] < 0) return fail(TPX_E_ARG, "tensor %d: numel < 0", t);
    if (list->grad[t] && (!list->param[t] || !list->exp_avg[t] || !list->exp_avg_sq[t])) return fail(TPX_E_ARG, "tensor %d: NULL pointer", t);
    if (list->numel[t] > nmax) nmax = list->numel[t];
  }
  if (nmax == 0) return 0;
  int64_t bx = (nmax + AD_PER_BLOCK - 1) / AD_PER_BLOCK;
  if (bx > 4096) bx = 4096;
  dim3 grid((unsigned)bx, (unsigned)list->n);
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  adam_dstep_kernel<<<grid, AD_THREADS, 0, (cudaStream_t)stream>>>(*list, lr, beta1, beta2, eps, weight_decay, step_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("adam_dstep_kernel");
}

extern "C" int tpx_adam_step(const tpx_tensor_list* list, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int32_t step, void* stream) {
  if (!list) return fail(TPX_E_ARG, "tensor list is NULL");
  if (list->n < 0 || list->n > TPX_MAX_TENSORS) return fail(TPX_E_ARG, "%d tensors (max %d)", list->n, TPX_MAX_TENSORS);
  if (step < 1) return fail(TPX_E_ARG, "step %d < 1", step);
  if (!(beta1 >= 0.5f && beta1 < 1.0f) || !(beta2 >= 0.0f && beta2 < 1.0f))
    return fail(TPX_E_UNSUPPORTED, "betas (%g, %g): beta1 in [0.5, 1), beta2 in [0, 1) supported", beta1, beta2);
  if (list->n == 0) return 0;
  int64_t nmax = 0;
  for (int t = 0; t < list->n; ++t) {
    if (list->numel[t] < 0) return fail(TPX_E_ARG, "tensor %d: numel < 0", t);
    if (list->grad[t] && (!list->param[t] || !list->exp_avg[t] || !list->exp_avg_sq[t])) return fail(TPX_E_ARG, "tensor %d: NULL pointer", t);
    if (list->numel[t] > nmax) nmax = list->numel[t];
  }
  if (nmax == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  int64_t bx = (nmax + AD_PER_BLOCK - 1) / AD_PER_BLOCK;
  if (bx > 4096) bx = 4096;
  dim3 grid((unsigned)bx, (unsigned)list->n);
  adam_kernel<<<grid, AD_THREADS, 0, (cudaStream_t)stream>>>(*list, step_size, bc2_sqrt, (float)(1.0 - (double)beta1), beta2,
                                                             (float)(1.0 - (double)beta2), eps, weight_decay);
  return cudaGetLastError() == cudaSuccess ? 0 : cuda_fail("adam_kernel");
}
