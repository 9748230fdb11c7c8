// Development probe 3 (not part of the product): kind::f16 (fp16 operands, fp32 accumulate) layouts used by the
// points-as-lanes kernels (csrc/tcp_*.cu):
//   T1  A from TMEM, two fp16 per 32-bit column (K even in the low half), B K-major SWIZZLE_128B
//   T2  A K-major SWIZZLE_128B from shared memory, B MN-major SWIZZLE_128B   (adjoint product: B = W as stored)
//   T3  A and B MN-major SWIZZLE_128B, M = 128 as two 64-wide blocks LBO apart (gradient product: K = points)
//   T4  as T3 with N = 128 (B two blocks LBO apart)
// and the issue / execute rate of each form.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {   // fp16 x fp16 -> fp32
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
               ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
                 "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// byte offset of fp16 element (row r, element e < 64) of a SWIZZLE_128B image with 128-byte rows
__host__ __device__ inline int swz16(int r, int e) { return r * 128 + ((((e >> 3) ^ (r & 7))) << 4) + (e & 7) * 2; }

// A: [128][128] fp16 (row-major, generic source), B: [128][128] fp16.  D: [128][128] fp32 (N columns used)
//  mode 1: D[p][n] = sum_{k<64} A[p][k] B[n][k]        A in TMEM (packed), B K-major, N = 64
//  mode 2: D[p][n] = sum_{k<64} A[p][k] B[k][n]        A K-major smem, B MN-major (rows k), N = 64
//  mode 3: D[m][n] = sum_{k<128} A[k][m] B[k][n]       A MN-major (m < 64 block 0, m >= 64 block 1), B MN-major, N = 64
//  mode 4: as 3 with N = 128 (n >= 64 in block 1)
__global__ void __launch_bounds__(128) probe(int mode, const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;             // 32 KB
  uint8_t* sB = smem + 32768;     // 32 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (mode == 1) {
    for (int e = tid; e < 64 * 64; e += 128) *(__half*)(sB + swz16(e / 64, e % 64)) = B[(e / 64) * 128 + e % 64];   // rows n, K contiguous
  } else if (mode == 2) {
    for (int e = tid; e < 128 * 64; e += 128) *(__half*)(sA + swz16(e / 64, e % 64)) = A[(e / 64) * 128 + e % 64];  // rows p, K contiguous
    for (int e = tid; e < 64 * 64; e += 128) *(__half*)(sB + swz16(e / 64, e % 64)) = B[(e / 64) * 128 + e % 64];   // rows k, n contiguous
  } else {
    for (int e = tid; e < 128 * 128; e += 128) {   // rows k; block (m / 64) 16 KB apart
      const int k = e / 128, m = e % 128;
      *(__half*)(sA + (m / 64) * 16384 + swz16(k, m % 64)) = A[k * 128 + m];
      *(__half*)(sB + (m / 64) * 16384 + swz16(k, m % 64)) = B[k * 128 + m];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tmem_base_s, tA = tmem_base_s + 128;
  if (mode == 1) {   // A -> TMEM, lane = row, column c = (A[r][2c] low, A[r][2c+1] high)
    uint32_t r[32];
    const int row = warp * 32 + lane;
    for (int c = 0; c < 32; ++c) {
      const __half2 h = __halves2half2(A[row * 128 + 2 * c], A[row * 128 + 2 * c + 1]);
      r[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    st32(tA + ((uint32_t)(warp * 32) << 16), r);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (elect_one()) {
      if (mode == 1) {
        const uint32_t idesc = make_idesc(128, 64, 0, 0);
        for (int ks = 0; ks < 4; ++ks) mma_ts(tD, tA + ks * 8, make_desc(b0 + ks * 32, 16, 1024), idesc, ks > 0);
      } else if (mode == 2) {
        const uint32_t idesc = make_idesc(128, 64, 0, 1);
        for (int ks = 0; ks < 4; ++ks) mma_ss(tD, make_desc(a0 + ks * 32, 16, 1024), make_desc(b0 + ks * 2048, 16384, 1024), idesc, ks > 0);
      } else if (mode == 3) {
        const uint32_t idesc = make_idesc(128, 64, 1, 1);
        for (int ks = 0; ks < 8; ++ks) mma_ss(tD, make_desc(a0 + ks * 2048, 16384, 1024), make_desc(b0 + ks * 2048, 16384, 1024), idesc, ks > 0);
      } else {
        const uint32_t idesc = make_idesc(128, 128, 1, 1);
        for (int ks = 0; ks < 8; ++ks) mma_ss(tD, make_desc(a0 + ks * 2048, 16384, 1024), make_desc(b0 + ks * 2048, 16384, 1024), idesc, ks > 0);
      }
      commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 4; ++h) {
    float v[32];
    ld32(tD + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 128 + h * 32 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(256));
}

// rate: FORM 0 = TS (A TMEM, B K-major), 1 = SS (A K-major, B MN-major), 2 = SS MN / MN
template <int N, int FORM, int NACC>
__global__ void __launch_bounds__(128) rate(int trips, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 131072 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;   // fp16 ones
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    constexpr uint32_t idesc = make_idesc(128, N, FORM == 2, FORM >= 1);
    const uint32_t tA = tmem + 448;
    long long t0 = 0;
    if (elect_one()) {
      uint64_t da[4], db[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        da[k] = FORM == 2 ? make_desc(a0 + k * 2048, 16384, 1024) : make_desc(a0 + k * 32, 16, 1024);
        db[k] = FORM >= 1 ? make_desc(b0 + k * 2048, 16384, 1024) : make_desc(b0 + k * 32, 16, 1024);
      }
      t0 = clock64();
      for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint32_t d = tmem + (uint32_t)((u % NACC) * N);
          if constexpr (FORM == 0) mma_ts(d, tA + (u % 4) * 8, db[u % 4], idesc, 1);
          else mma_ss(d, da[u % 4], db[u % 4], idesc, 1);
        }
      }
      commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      if (cycles) *cycles = t1 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int FORM, int NACC>
void run_rate(long long* dC, const char* name) {
  CK(cudaFuncSetAttribute(rate<N, FORM, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096));
  rate<N, FORM, NACC><<<1, 128, 132096>>>(256, dC);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<N, FORM, NACC><<<148, 128, 132096>>>(2048, nullptr);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate<N, FORM, NACC><<<148, 128, 132096>>>(2048, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 148.0 * 2048 * 16 * 2.0 * 128 * N * 16;
  printf("rate %-34s: 1 CTA %.1f cycles/MMA (floor %d); 148 CTAs %.1f TFLOP/s, %.1f ns/MMA\n", name, (double)c / (256 * 16), N / 2,
         flops / ms / 1e9, ms * 1e6 / (2048 * 16));
}

int main() {
  const int SMEM = 65536 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  std::vector<__half> hA(128 * 128), hB(128 * 128);
  std::vector<float> fA(128 * 128), fB(128 * 128), hD(128 * 128);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2half((float)rand() / RAND_MAX * 2 - 1); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2half((float)rand() / RAND_MAX * 2 - 1); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, hD.size() * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  for (int mode = 1; mode <= 4; ++mode) {
    CK(cudaMemset(dD, 0, hD.size() * 4));
    probe<<<1, 128, SMEM>>>(mode, dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    const int N = mode == 4 ? 128 : 64;
    double err = 0, err_swapped = 0, mx = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double tr = 0, sw = 0;
        if (mode == 1) {
          for (int k = 0; k < 64; ++k) { tr += (double)fA[m * 128 + k] * fB[n * 128 + k]; sw += (double)fA[m * 128 + (k ^ 1)] * fB[n * 128 + k]; }
        } else if (mode == 2) {
          for (int k = 0; k < 64; ++k) tr += (double)fA[m * 128 + k] * fB[k * 128 + n];
        } else {
          for (int k = 0; k < 128; ++k) tr += (double)fA[k * 128 + m] * fB[k * 128 + n];
        }
        err = fmax(err, fabs(hD[m * 128 + n] - tr)); err_swapped = fmax(err_swapped, fabs(hD[m * 128 + n] - sw)); mx = fmax(mx, fabs(tr));
      }
    printf("mode %d: max|D|=%.3f err %.3e (half-swapped packing: %.3e)  D[0][0..2]=%.4f %.4f %.4f  D[127][N-1]=%.4f\n", mode, mx, err, err_swapped,
           hD[0], hD[1], hD[2], hD[127 * 128 + N - 1]);
  }
  run_rate<64, 0, 1>(dC, "f16 TS N=64");
  run_rate<64, 0, 4>(dC, "f16 TS N=64 nacc=4");
  run_rate<64, 1, 1>(dC, "f16 SS A K-major B MN N=64");
  run_rate<64, 2, 1>(dC, "f16 SS MN/MN N=64");
  run_rate<128, 2, 1>(dC, "f16 SS MN/MN N=128");
  run_rate<80, 2, 1>(dC, "f16 SS MN/MN N=80");
  run_rate<16, 0, 4>(dC, "f16 TS N=16 nacc=4");
  run_rate<32, 0, 4>(dC, "f16 TS N=32 nacc=4");
  printf("done\n");
  return 0;
}
