// Development probe (not part of the product): validates the tcgen05 primitives the jet kernels
// rely on -- smem descriptors for the chunk-major no-swizzle layout in K-major and MN-major form,
// A operand from TMEM, tf32 kind, M=128 N=64 -- and times the MMA issue rate in each mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
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
__device__ __forceinline__ void st32(uint32_t taddr, const float (&v)[32]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
               : "memory");
}

// mode 0: SS K-major A (128x64) x K-major B (64x64) -> D (128x64)
// mode 1: TS  A from TMEM
// mode 2: SS MN-major A' (M=128 x K=128) x MN-major B' (N=64 x K=128): D'[j,i] = sum_r Z[r,j] Y[r,i]
// mode 3/4: timing of `iters` back-to-back MMAs, SS / TS
__global__ void __launch_bounds__(128) probe(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                             float* __restrict__ D, int iters, long long* cycles, int lbo_sel) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float* sA = reinterpret_cast<float*>(smem);             // up to 64 KB
  float* sB = reinterpret_cast<float*>(smem + 65536);     // up to 32 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem, tA = tmem + 64;               // D: cols 0..63, A (TS): cols 64..127

  if (mode == 0 || mode == 1 || mode == 3 || mode == 4) {
    // A global row-major (128 x 64), B global row-major (64 x 64) [n][k]
    for (int e = tid; e < 128 * 64; e += 128) {
      const int r = e / 64, k = e % 64;
      sA[(k / 4) * (128 * 4) + r * 4 + (k % 4)] = A[e];
    }
    for (int e = tid; e < 64 * 64; e += 128) {
      const int n = e / 64, k = e % 64;
      sB[(k / 4) * (64 * 4) + n * 4 + (k % 4)] = B[e];
    }
    if (mode == 1 || mode == 4) {
      float v[32];
      for (int h = 0; h < 2; ++h) {
        for (int i = 0; i < 32; ++i) v[i] = A[(warp * 32 + lane) * 64 + h * 32 + i];
        st32(tA + ((uint32_t)(warp * 32) << 16) + h * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  } else {
    // A: global (128 rows r) x (128 cols j); logical A'[m=j][k=r].  B: global (128 rows r) x (64 cols i); B'[n=i][k=r]
    for (int e = tid; e < 128 * 128; e += 128) {
      const int r = e / 128, j = e % 128;
      if (mode == 5) {   // A K-major from the first 64 columns: A[m=r][k=j<64]
        if (j < 64) sA[(j / 4) * (128 * 4) + r * 4 + (j % 4)] = A[r * 128 + j];
      } else {           // MN-major: [j/4][r][j%4]
        sA[(j / 4) * (128 * 4) + r * 4 + (j % 4)] = A[e];
      }
    }
    for (int e = tid; e < 128 * 64; e += 128) {
      const int r = e / 64, i = e % 64;
      if (mode == 6) sB[(r / 4) * (64 * 4) + i * 4 + (r % 4)] = B[e];          // K-major B[n=i][k=r]
      else sB[(i / 4) * (128 * 4) + r * 4 + (i % 4)] = B[e];                    // MN-major [i/4][r][i%4]
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    long long t0 = clock64();
    if (mode == 0 || mode == 1) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t db = make_desc(b0 + ks * 2 * (64 * 16), 64 * 16, 128);
        if (mode == 0) {
          const uint64_t da = make_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
          mma_ss(tD, da, db, idesc, ks > 0);
        } else {
          mma_ts(tD, tA + ks * 8, db, idesc, ks > 0);
        }
      }
    } else if (mode == 2) {
      const uint32_t idesc = make_idesc(128, 64, 1, 1);
      for (int ks = 0; ks < 16; ++ks) {      // 8 rows (K) per step
        const uint64_t da = make_desc(a0 + ks * 128, lbo_sel ? 128 * 16 : 128, lbo_sel ? 128 : 128 * 16);
        const uint64_t db = make_desc(b0 + ks * 128, lbo_sel ? 128 * 16 : 128, lbo_sel ? 128 : 128 * 16);
        mma_ss(tD, da, db, idesc, ks > 0);
      }
    } else if (mode == 5) {                  // A K-major (K = 64 cols of A), B MN-major (k = rows 0..63 of B)
      const uint32_t idesc = make_idesc(128, 64, 0, 1);
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = make_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
        const uint64_t db = make_desc(b0 + ks * 128, lbo_sel ? 128 * 16 : 128, lbo_sel ? 128 : 128 * 16);
        mma_ss(tD, da, db, idesc, ks > 0);
      }
    } else if (mode == 6) {                  // A MN-major (M = 128 cols j, K = rows), B K-major (k = rows)
      const uint32_t idesc = make_idesc(128, 64, 1, 0);
      for (int ks = 0; ks < 16; ++ks) {
        const uint64_t da = make_desc(a0 + ks * 128, lbo_sel ? 128 * 16 : 128, lbo_sel ? 128 : 128 * 16);
        const uint64_t db = make_desc(b0 + ks * 2 * (64 * 16), 64 * 16, 128);
        mma_ss(tD, da, db, idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int ks = it & 7;
        const uint64_t db = make_desc(b0 + ks * 2 * (64 * 16), 64 * 16, 128);
        if (mode == 3) {
          const uint64_t da = make_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
          mma_ss(tD, da, db, idesc, it > 0);
        } else {
          mma_ts(tD, tA + ks * 8, db, idesc, it > 0);
        }
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read D: warp w reads lanes 32w.., 64 columns
  for (int h = 0; h < 2; ++h) {
    float v[32];
    ld32(tD + ((uint32_t)(warp * 32) << 16) + h * 32, v);
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + h * 32 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// timing only (operands are whatever is in smem): N, kind (0 tf32, 1 bf16), nacc D tiles in rotation, TS or SS
__global__ void __launch_bounds__(128) rate(int N, int kind, int nacc, int ts, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (65536 + 65536) / 4; e += 128) reinterpret_cast<float*>(smem)[e] = 0.0f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    uint32_t idesc = make_idesc(128, N, 0, 0);
    if (kind == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(a0, 128 * 16, 128), db = make_desc(b0, N * 16, 128);
    const uint32_t tA = tmem + 448;   // TS: 8..16 columns of A at the end (garbage values)
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)((it % nacc) * N);
      if (ts) mma_ts(d, tA, db, idesc, 1);
      else if (kind == 0) mma_ss(d, da, db, idesc, 1);
      else mma_ss_f16(d, da, db, idesc, 1);
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int SMEM = 65536 + 32768;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  std::vector<float> hA(128 * 128), hB(128 * 64), hD(128 * 64);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& v : hB) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  for (int lbo_sel = 0; lbo_sel < 2; ++lbo_sel)
  for (int mode : {0, 1, 2, 5, 6}) {
    if (lbo_sel && mode < 2) continue;
    CK(cudaMemset(dD, 0, hD.size() * 4));
    probe<<<1, 128, SMEM>>>(mode, dA, dB, dD, 0, dC, lbo_sel);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double e_exact = 0, e_trunc = 0, e_rn = 0, mx = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ex = 0, tr = 0, rn = 0;
        if (mode < 2) {
          for (int k = 0; k < 64; ++k) {
            float a = hA[m * 64 + k], b = hB[n * 64 + k];
            ex += (double)a * b; tr += (double)tf32_trunc(a) * tf32_trunc(b); rn += (double)tf32_rn(a) * tf32_rn(b);
          }
        } else if (mode == 5) {
          for (int k = 0; k < 64; ++k) {
            float a = hA[m * 128 + k], b = hB[k * 64 + n];
            ex += (double)a * b; tr += (double)tf32_trunc(a) * tf32_trunc(b); rn += (double)tf32_rn(a) * tf32_rn(b);
          }
        } else {
          for (int r = 0; r < 128; ++r) {
            float a = hA[r * 128 + m], b = hB[r * 64 + n];
            ex += (double)a * b; tr += (double)tf32_trunc(a) * tf32_trunc(b); rn += (double)tf32_rn(a) * tf32_rn(b);
          }
        }
        const double got = hD[m * 64 + n];
        e_exact = fmax(e_exact, fabs(got - ex)); e_trunc = fmax(e_trunc, fabs(got - tr)); e_rn = fmax(e_rn, fabs(got - rn));
        mx = fmax(mx, fabs(ex));
      }
    printf("mode %d lbo_sel %d: max|D|=%.3f  err vs exact %.3e  vs tf32-trunc %.3e  vs tf32-rn %.3e   D[0][0..2]=%.3f %.3f %.3f D[1][0]=%.3f\n", mode, lbo_sel, mx, e_exact, e_trunc, e_rn, hD[0], hD[1], hD[2], hD[64]);
  }
  for (int mode = 3; mode <= 4; ++mode)
    for (int iters : {8, 64, 512, 4096}) {
      probe<<<1, 128, SMEM>>>(mode, dA, dB, dD, iters, dC, 0);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
      printf("mode %d (%s) iters %d: %lld cycles, %.1f cycles/MMA\n", mode, mode == 3 ? "SS" : "TS", iters, c, (double)c / iters);
    }
  // all SMs busy: does the per-SM rate hold?
  for (int mode = 3; mode <= 4; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    probe<<<148, 128, SMEM>>>(mode, dA, dB, dD, iters, nullptr, 0);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    probe<<<148, 128, SMEM>>>(mode, dA, dB, dD, iters, nullptr, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 148.0 * iters * 2.0 * 128 * 64 * 8;
    printf("mode %d full chip: %.3f ms, %.1f TFLOP/s tf32, %.1f ns/MMA\n", mode, ms, flops / ms / 1e9, ms * 1e6 / iters);
  }
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  for (int kind = 0; kind < 2; ++kind)
    for (int ts = 0; ts < 2; ++ts) {
      if (kind == 1 && ts) continue;
      for (int N : {64, 128, 256})
        for (int nacc : {1, 2}) {
          if (N * nacc > 448) continue;
          rate<<<1, 128, 131072>>>(N, kind, nacc, ts, 2048, dC);
          CK(cudaDeviceSynchronize());
          long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
          printf("rate kind=%s %s N=%d nacc=%d: %.1f cycles/MMA (floor %d)\n", kind ? "bf16" : "tf32", ts ? "TS" : "SS", N, nacc, (double)c / 2048, N / 2);
        }
    }
  printf("done\n");
  return 0;
}
