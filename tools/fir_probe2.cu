// fir_probe2.cu — second probe for the tensor-core FIR stage:
//  P1. kind::f16 with MIXED operand formats: A = bf16 (MN-major x tile), B = fp16 (Toeplitz taps, K-major).
//  P2. A operand in TMEM (written with tcgen05.st from registers, two bf16 K elements per 32-bit column),
//      B = fp16 K-major in smem, N = 16, K = 48 (the down-sampling FIR), plus its cadence.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fir_probe2 fir_probe2.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// afmt / bfmt: 0 = f16, 1 = bf16
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MODE 0: SS, A = bf16 MN-major [16 groups][ROWS][8] (shift rows), B = fp16 [KT/8][16][8], KT = 16
// MODE 1: TS, A[m][k] bf16 written to TMEM columns (two K elements per column) from Ag[m][KTOT], window at
//         element offset `shift` (even multiple of 8), B = fp16, KT = 48
template <int MODE, int KT>
__global__ void __launch_bounds__(128) k_mixed(const __nv_bfloat16* Ag, const __half* Bg, float* D, int shift, int ROWS,
                                               long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int N = 16;
  const int a_elems = MODE == 0 ? 16 * ROWS * 8 : 0;
  __nv_bfloat16* sA = (__nv_bfloat16*)smem;
  __half* sB = (__half*)(smem + ((a_elems * 2 + 1023) / 1024) * 1024);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < a_elems; i += 128) sA[i] = Ag[i];
  for (int i = tid; i < N * KT; i += 128) { int n = i / KT, k = i % KT; sB[((k / 8) * N + n) * 8 + (k % 8)] = Bg[i]; }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  const uint32_t A_COL = 64;     // TMEM columns of the A operand (MODE 1): ROWS = K elements per row (<= 256)
  if (MODE == 1) {
    // lane L = tid writes its row: columns A_COL + k/2 hold elements (k, k+1), low half = even k
    const uint32_t row = tid;
    for (int c0 = 0; c0 < ROWS / 2; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const int k = 2 * (c0 + j);
        const uint32_t lo = k < ROWS ? (uint32_t)__half_as_ushort(__float2half(__bfloat162float(Ag[row * ROWS + k]))) : 0u;
        const uint32_t hi = k + 1 < ROWS ? (uint32_t)__half_as_ushort(__float2half(__bfloat162float(Ag[row * ROWS + k + 1]))) : 0u;
        v[j] = lo | (hi << 16);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + A_COL + c0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (tid == 0) {
    if (MODE == 0) {
      const uint32_t idesc = make_idesc(128, N, 1, 1, 0);   // A bf16 MN-major, B fp16
      mma_ss(tm, make_desc(smem_u32(sA) + shift * 16, 128, ROWS * 16), make_desc(smem_u32(sB), N * 16, 128), idesc, 0);
    } else {
      const uint32_t idesc = make_idesc(128, N, 0, 0, 0);   // A fp16 (TMEM), B fp16
      for (int ks = 0; ks < KT / 16; ++ks)
        mma_ts(tm, tm + A_COL + shift / 2 + ks * 8, make_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 128), idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[16];
  ld16(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
  // cadence: the same MMAs in a tight loop into a second accumulator
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0 && iters > 0) {
    long long t0 = clock64();
    if (MODE == 0) {
      const uint32_t idesc = make_idesc(128, N, 1, 1, 0);
      const uint64_t bd = make_desc(smem_u32(sB), N * 16, 128);
      uint64_t ad[4];
      for (int i = 0; i < 4; ++i) ad[i] = make_desc(smem_u32(sA) + i * 128, 128, ROWS * 16);
      for (int it = 0; it < iters; it += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_ss(tm + 32 + (i & 1) * 16, ad[i], bd, idesc, 1);
      }
    } else {
      const uint32_t idesc = make_idesc(128, N, 0, 0, 0);
      uint64_t bd[3];
      for (int ks = 0; ks < 3; ++ks) bd[ks] = make_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 128);
      for (int it = 0; it < iters; it += 6) {
#pragma unroll
        for (int i = 0; i < 6; ++i) mma_ts(tm + 32 + (i / 3) * 16, tm + A_COL + (i % 3) * 8 + (i / 3) * 16, bd[i % 3], idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 1);
    cyc[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

template <int MODE, int KT>
static bool run_mixed(int shift, int ROWS) {
  constexpr int N = 16;
  const int a_elems = MODE == 0 ? 16 * ROWS * 8 : 128 * ROWS;
  std::vector<__nv_bfloat16> hA(a_elems);
  std::vector<__half> hB(N * KT);
  std::vector<float> fA(a_elems), fB(N * KT);
  srand(7 + shift);
  for (int i = 0; i < a_elems; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * KT; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  __nv_bfloat16* dA; __half* dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8 * 148));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  size_t smem = (size_t)(MODE == 0 ? 16 * ROWS * 16 : 0) + 2048 + N * KT * 2;
  auto kern = k_mixed<MODE, KT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, 128, smem>>>(dA, dB, dD, shift, ROWS, dC, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  mixed mode %d shift=%d: CUDA error %s\n", MODE, shift, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) {
        double a = MODE == 0 ? fA[((m / 8) * ROWS + shift + k) * 8 + (m % 8)] : fA[m * ROWS + shift + k];
        s += a * fB[n * KT + k];
      }
      double err = fabs(s - hD[m * N + n]);
      if (!(err <= 1e30)) err = 1e30;
      if (err > maxerr) maxerr = err;
    }
  printf("  %s shift=%-3d max|err|=%.3e  %s\n", MODE == 0 ? "SS A=bf16 MN-major, B=fp16, K=16 N=16" : "TS A=fp16 in TMEM (2 per column, low half = even k), B=fp16, K=48 N=16", shift,
         maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  if (maxerr < 1e-3) {
    const int iters = 6000;
    kern<<<148, 128, smem>>>(dA, dB, dD, 0, ROWS, dC, iters);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), dC, 8 * 148, cudaMemcpyDeviceToHost));
    double mean = 0; for (auto c : h) mean += c; mean /= 148;
    printf("    cadence: %.1f cyc/MMA\n", mean / iters);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return maxerr < 1e-3;
}

int main() {
  bool ok = true;
  printf("[P2] A in TMEM (bf16 pairs per column), B fp16, TS\n");
  for (int shift : {0, 8, 16, 56}) ok &= run_mixed<1, 48>(shift, 176);
  printf("probe2: %s\n", ok ? "ALL PASS" : "SOME FAIL");
  return ok ? 0 : 1;
}
