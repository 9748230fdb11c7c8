// Development probe 2 (not part of the product): 128-byte-swizzled operand buffers that serve BOTH
// as K-major operand (rows = M, neurons = K) and as MN-major operand (neurons = M/N, rows = K),
// tf32, plus a clean measurement of the tcgen05.mma issue/execute rate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// layout_type: 0 none, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn, int fmt /*2 tf32, 1 bf16*/) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
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

// swizzled offset (floats) of element (row r, col j) in a buffer of R rows: [j/32][r/8][r%8][ (chunk ^ r%8) ][4]
__host__ __device__ inline int swz(int r, int j, int R) {
  return (j / 32) * (R * 32) + (r / 8) * 256 + (r % 8) * 32 + ((((j % 32) / 4) ^ (r % 8)) * 4) + (j % 4);
}

// mode 0: K-major SW128: D[r, n] = sum_j A[r, j] B[n, j], A 128x64, B 64x64
// mode 1: MN-major SW128: D[j, i] = sum_r Z[r, j] Y[r, i],  Z 128 rows x 128, Y 128 rows x 64
// mode 2: A MN-major, B K-major:  D[j, n] = sum_r Z[r, j] Bk[n, r],  Bk 64 x 128
// mode 3: A K-major, B MN-major:  D[r, i] = sum_j A[r, j] Y[j, i]  (Y rows 0..63 as K)
__global__ void __launch_bounds__(128) probe(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + 65536);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // stage
  if (mode == 0 || mode == 3) {
    for (int e = tid; e < 128 * 64; e += 128) sA[swz(e / 64, e % 64, 128)] = A[(e / 64) * 128 + e % 64];   // A = first 64 cols of hA (ld 128)
  } else {
    for (int e = tid; e < 128 * 128; e += 128) sA[swz(e / 128, e % 128, 128)] = A[e];
  }
  if (mode == 0) {
    for (int e = tid; e < 64 * 64; e += 128) sB[swz(e / 64, e % 64, 64)] = B[(e % 64) * 64 + e / 64];      // B[n][j] = hB[j][n]
  } else if (mode == 2) {
    for (int e = tid; e < 64 * 128; e += 128) sB[swz(e / 128, e % 128, 64)] = B[(e % 128) * 64 + e / 128];  // Bk[n][r] = hB[r][n], 64 rows x 128 K
  } else {
    for (int e = tid; e < 128 * 64; e += 128) sB[swz(e / 64, e % 64, 128)] = B[e];                         // Y[r][i], 128 rows
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tmem_base_s;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (elect_one()) {
      if (mode == 0) {
        const uint32_t idesc = make_idesc(128, 64, 0, 0, 2);
        for (int ks = 0; ks < 8; ++ks) {   // K-major: k-atom (ks/4) of 32 cols, 32 B per step inside the atom
          const uint64_t da = make_desc(a0 + (ks / 4) * (128 * 128) + (ks % 4) * 32, 16, 1024, 2);
          const uint64_t db = make_desc(b0 + (ks / 4) * (64 * 128) + (ks % 4) * 32, 16, 1024, 2);
          mma_ss<0>(tD, da, db, idesc, ks > 0);
        }
      } else if (mode == 1) {
        const uint32_t idesc = make_idesc(128, 64, 1, 1, 2);
        for (int ks = 0; ks < 16; ++ks) {  // MN-major: K = rows, 8 rows (one 1 KB group) per step; MN blocks of 32 at LBO
          const uint64_t da = make_desc(a0 + ks * 1024, 128 * 128, 1024, 2);
          const uint64_t db = make_desc(b0 + ks * 1024, 128 * 128, 1024, 2);
          mma_ss<0>(tD, da, db, idesc, ks > 0);
        }
      } else if (mode == 2) {
        const uint32_t idesc = make_idesc(128, 64, 1, 0, 2);
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t da = make_desc(a0 + ks * 1024, 128 * 128, 1024, 2);
          const uint64_t db = make_desc(b0 + (ks / 4) * (64 * 128) + (ks % 4) * 32, 16, 1024, 2);
          mma_ss<0>(tD, da, db, idesc, ks > 0);
        }
      } else {
        const uint32_t idesc = make_idesc(128, 64, 0, 1, 2);
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = make_desc(a0 + (ks / 4) * (128 * 128) + (ks % 4) * 32, 16, 1024, 2);
          const uint64_t db = make_desc(b0 + ks * 1024, 128 * 128, 1024, 2);
          mma_ss<0>(tD, da, db, idesc, ks > 0);
        }
      }
      commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    float v[32];
    ld32(tD + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + h * 32 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(64));
}

// clean rate measurement: warp 0 (one elected lane) issues UNROLL straight-line MMAs per loop trip
template <int N, int KIND, int TS, int NACC>
__global__ void __launch_bounds__(128) rate(int trips, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 131072 / 4; e += 128) reinterpret_cast<float*>(smem)[e] = 1.0f;
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
    constexpr uint32_t idesc = make_idesc(128, N, 0, 0, KIND == 0 ? 2 : 1);
    const uint32_t tA = tmem + 448;
    long long t0 = 0;
    if (elect_one()) {
      uint64_t da[4], db[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        da[k] = make_desc(a0 + k * 32, 16, 1024, 2);
        db[k] = make_desc(b0 + k * 32, 16, 1024, 2);
      }
      t0 = clock64();
      for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint32_t d = tmem + (uint32_t)((u % NACC) * N);
          if constexpr (TS) mma_ts(d, tA + (u % 4) * 8, db[u % 4], idesc, 1);
          else mma_ss<KIND>(d, da[u % 4], db[u % 4], idesc, 1);
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

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int N, int KIND, int TS, int NACC>
void run_rate(long long* dC, const char* name) {
  CK(cudaFuncSetAttribute(rate<N, KIND, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096));
  rate<N, KIND, TS, NACC><<<1, 128, 132096>>>(256, dC);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate<N, KIND, TS, NACC><<<148, 128, 132096>>>(2048, nullptr);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  rate<N, KIND, TS, NACC><<<148, 128, 132096>>>(2048, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double k = KIND == 0 ? 8 : 16;
  const double flops = 148.0 * 2048 * 16 * 2.0 * 128 * N * k;
  printf("rate %-28s: 1 CTA %.1f cycles/MMA (floor %d); 148 CTAs %.1f TFLOP/s, %.1f ns/MMA\n", name, (double)c / (256 * 16), N / 2,
         flops / ms / 1e9, ms * 1e6 / (2048 * 16));
}

int main() {
  const int SMEM = 65536 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  std::vector<float> hA(128 * 128), hB(128 * 64), hD(128 * 64);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& v : hB) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  for (int mode = 0; mode < 4; ++mode) {
    CK(cudaMemset(dD, 0, hD.size() * 4));
    probe<<<1, 128, SMEM>>>(mode, dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double e_tr = 0, mx = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double tr = 0;
        if (mode == 0) for (int j = 0; j < 64; ++j) tr += (double)tf32_trunc(hA[m * 128 + j]) * tf32_trunc(hB[j * 64 + n]);
        else if (mode == 1) for (int r = 0; r < 128; ++r) tr += (double)tf32_trunc(hA[r * 128 + m]) * tf32_trunc(hB[r * 64 + n]);
        else if (mode == 2) for (int r = 0; r < 128; ++r) tr += (double)tf32_trunc(hA[r * 128 + m]) * tf32_trunc(hB[r * 64 + n]);
        else for (int j = 0; j < 64; ++j) tr += (double)tf32_trunc(hA[m * 128 + j]) * tf32_trunc(hB[j * 64 + n]);
        e_tr = fmax(e_tr, fabs(hD[m * 64 + n] - tr)); mx = fmax(mx, fabs(tr));
      }
    printf("mode %d: max|D|=%.3f err vs tf32-trunc %.3e   D[0][0..2]=%.3f %.3f %.3f\n", mode, mx, e_tr, hD[0], hD[1], hD[2]);
  }
  run_rate<64, 0, 0, 1>(dC, "tf32 SS N=64 nacc=1");
  run_rate<64, 0, 0, 4>(dC, "tf32 SS N=64 nacc=4");
  run_rate<64, 0, 1, 1>(dC, "tf32 TS N=64 nacc=1");
  run_rate<64, 0, 1, 4>(dC, "tf32 TS N=64 nacc=4");
  run_rate<128, 0, 0, 1>(dC, "tf32 SS N=128 nacc=1");
  run_rate<128, 0, 1, 1>(dC, "tf32 TS N=128 nacc=1");
  run_rate<256, 0, 0, 1>(dC, "tf32 SS N=256 nacc=1");
  run_rate<256, 0, 1, 1>(dC, "tf32 TS N=256 nacc=1");
  run_rate<64, 1, 0, 1>(dC, "bf16 SS N=64 nacc=1");
  run_rate<128, 1, 0, 1>(dC, "bf16 SS N=128 nacc=1");
  run_rate<256, 1, 0, 1>(dC, "bf16 SS N=256 nacc=1");
  printf("done\n");
  return 0;
}
