// Development probe: shared-memory store throughput on one SM (16 warps), scalar vs 128-bit, conflict-free.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/sts_probe sts_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(512) sts32(float* out, long long* cyc, int iters) {
  extern __shared__ float sm[];
  const int t = threadIdx.x;
  float v = (float)t;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k)   // one 128-byte row per warp store
      asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(sm + ((t >> 5) * 16 + k) * 32 + (t & 31))), "f"(v + k) : "memory");
    v += 1.0f;
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) *cyc = t1 - t0;
  out[t] = sm[t];
}
__global__ void __launch_bounds__(512) sts128(float* out, long long* cyc, int iters) {
  extern __shared__ float sm[];
  const int t = threadIdx.x;
  float v = (float)t;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(sm + ((t >> 5) * 4 + k) * 128 + (t & 31) * 4)), "f"(v), "f"(v + 1), "f"(v + 2), "f"(v + k) : "memory");
    v += 1.0f;
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) *cyc = t1 - t0;
  out[t] = sm[t];
}
__global__ void __launch_bounds__(512) lds32(float* out, long long* cyc, int iters) {
  extern __shared__ float sm[];
  const int t = threadIdx.x;
  for (int i = t; i < 16 * 16 * 32; i += 512) sm[i] = (float)i;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sm[((t >> 5) * 16 + k) * 32 + ((t + i) & 31)];
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) *cyc = t1 - t0;
  out[t] = acc;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int iters = 2000, smem = 16 * 16 * 32 * 4;
  long long c;
  sts32<<<1, 512, smem>>>(out, cyc, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("STS.32 : %.2f cycles per warp store (128 B), %.1f B/clk\n", (double)c / (iters * 16.0 * 16), 128.0 * iters * 16 * 16 / c);
  sts128<<<1, 512, smem>>>(out, cyc, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("STS.128: %.2f cycles per warp store (512 B), %.1f B/clk\n", (double)c / (iters * 4.0 * 16), 512.0 * iters * 4 * 16 / c);
  lds32<<<1, 512, smem>>>(out, cyc, iters); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.32 : %.2f cycles per warp load (128 B), %.1f B/clk\n", (double)c / (iters * 16.0 * 16), 128.0 * iters * 16 * 16 / c);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
