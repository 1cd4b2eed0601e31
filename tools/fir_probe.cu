// fir_probe.cu — hardware probe for the tensor-core FIR stage of the fused AMP kernel:
//  A. tcgen05.mma kind::f16 with an MN-major SWIZZLE_NONE A operand: the TMA-staged x tile
//     [row group][time row][8 channels] (16-byte rows) read as A[M = (group, channel), K = time] with a
//     start-address offset of 16*shift bytes selecting the K window, B = banded Toeplitz taps (K-major).
//  B. K-major A with LBO = 2048 B (the s tile [K/8][128 lanes][8]), K = 48, N = 16 (down-sampling FIR).
//  C. tcgen05.ld throughput with 16 warps (4 per lane quarter), x16 and x32 shapes.
//  D. cadence of small-N MMAs (N = 16 / 32, K = 16) issued by one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fir_probe fir_probe.cu
#include <cuda_bf16.h>
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
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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

// ---------------------------------------------------------------- A / B: descriptor semantics
// mode 0 (MN-major A): sA = [16 groups][ROWS][8] bf16, A[m = g*8+c][k] = sA[g][shift + k][c], KT = 16
// mode 1 (K-major A, LBO 2048): sA = [KCH][128][8], A[m][k] = sA[kc0 + k/8][m][k%8], KT = 48
// B: K-major [KT/8][16][8];  D[m][n] = sum_k A[m][k] * B[n][k]
template <int MODE, int KT>
__global__ void __launch_bounds__(128) k_desc(const __nv_bfloat16* Ag, const __nv_bfloat16* Bg, float* D, int shift, int ROWS, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int N = 16;
  const int a_elems = MODE == 0 ? 16 * ROWS * 8 : ROWS * 128 * 8;     // mode 1: ROWS = number of K chunks
  __nv_bfloat16* sA = (__nv_bfloat16*)smem;
  __nv_bfloat16* sB = sA + a_elems;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < a_elems; i += 128) sA[i] = Ag[i];
  for (int i = tid; i < N * KT; i += 128) { int n = i / KT, k = i % KT; sB[((k / 8) * N + n) * 8 + (k % 8)] = Bg[i]; }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    if (MODE == 0) {
      const uint32_t idesc = make_idesc(128, N, 1);
      // MN-major no-swizzle: SBO = stride between 8-element M groups, LBO = stride between 8-row K groups
      uint32_t lbo = 128, sbo = ROWS * 16;
      if (variant) { uint32_t t = lbo; lbo = sbo; sbo = t; }
      mma_bf16(tm, make_desc(smem_u32(sA) + shift * 16, lbo, sbo), make_desc(smem_u32(sB), N * 16, 128), idesc, 0);
    } else {
      const uint32_t idesc = make_idesc(128, N, 0);
      for (int ks = 0; ks < KT / 16; ++ks)
        mma_bf16(tm, make_desc(smem_u32(sA) + (shift + 2 * ks) * 2048, 2048, 128),
                 make_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 128), idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = warp * 32 + (tid & 31);
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(tm + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j = 0; j < 16; ++j) D[row * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(32));
}

template <int MODE, int KT>
static bool run_desc(int shift, int ROWS, int variant) {
  constexpr int N = 16;
  const int a_elems = MODE == 0 ? 16 * ROWS * 8 : ROWS * 128 * 8;
  std::vector<__nv_bfloat16> hA(a_elems), hB(N * KT);
  std::vector<float> fA(a_elems), fB(N * KT);
  srand(99 + shift);
  for (int i = 0; i < a_elems; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * KT; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  size_t smem = (size_t)a_elems * 2 + N * KT * 2;
  auto kern = k_desc<MODE, KT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, 128, smem>>>(dA, dB, dD, shift, ROWS, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  desc mode %d shift=%d: CUDA error %s\n", MODE, shift, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) {
        double a = MODE == 0 ? fA[((m / 8) * ROWS + shift + k) * 8 + (m % 8)] : fA[((shift + k / 8) * 128 + m) * 8 + (k % 8)];
        s += a * fB[n * KT + k];
      }
      double err = fabs(s - hD[m * N + n]);
      if (!(err <= 1e30)) err = 1e30;
      if (err > maxerr) maxerr = err;
    }
  printf("  %s  shift=%-3d rows=%-3d variant=%d  max|err|=%.3e  %s\n",
         MODE == 0 ? "MN-major A (x tile), K=16, N=16" : "K-major A LBO=2048 (s tile), K=48, N=16", shift, ROWS, variant, maxerr,
         maxerr < 1e-3 ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3;
}

// ---------------------------------------------------------------- C: tcgen05.ld throughput
template <int X>
__global__ void __launch_bounds__(512) k_ldtm(long long* cycles, float* sink, int iters, int nwarps) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid / 32;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (uint32_t)(((it * 32) + (warp >> 2) * 64) & 255);
      if (X == 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tm + col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
      } else {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tm + col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 512 + tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

template <int X>
static void run_ldtm(int nwarps) {
  long long* d; float* s;
  CK(cudaMalloc(&d, 8 * 148)); CK(cudaMalloc(&s, 4 * 148 * 512));
  const int iters = 2000;
  k_ldtm<X><<<148, 512>>>(d, s, iters, nwarps);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 8 * 148, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto c : h) mean += c; mean /= 148;
  const double bytes = (double)iters * nwarps * 32 * X * 4;
  printf("  tcgen05.ld 32x32b.x%-2d  %2d warps: %.1f cyc per warp-load, %.1f B/clk/SM\n", X, nwarps, mean / iters, bytes / mean);
  cudaFree(d); cudaFree(s);
}

// ---------------------------------------------------------------- D: small-N MMA cadence
template <int N>
__global__ void __launch_bounds__(128) k_smalln(long long* cycles, int iters, int a_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (int)(64 * 1024 / 4); i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 255);
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, a_mn);
    const uint64_t hiA = a_mn ? make_desc(0, 128, 96 * 16) : make_desc(0, 2048, 128);
    const uint64_t hiB = make_desc(0, N * 16, 128);
    const uint32_t aU = smem_u32(smem) >> 4, bU = (smem_u32(smem) + 48 * 1024) >> 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a0 = aU + (uint32_t)((it % 10) * 8);
      mma_bf16(tm + (uint32_t)((it & 7) * N), hiA | a0, hiB | bU, idesc, it >= 8);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

template <int N>
static void run_smalln(int a_mn) {
  long long* d; CK(cudaMalloc(&d, 8 * 148));
  auto kern = k_smalln<N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int iters = 8000;
  kern<<<148, 128, 64 * 1024>>>(d, iters, a_mn);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 8 * 148, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto c : h) mean += c; mean /= 148;
  printf("  MMA M=128 N=%-3d K=16 A %s: %.1f cyc/MMA (floor %.0f)\n", N, a_mn ? "MN-major" : "K-major ", mean / iters, 128.0 * N / 256);
  cudaFree(d);
}


// tight variant: descriptors precomputed, 8 MMAs per loop iteration, optionally issued by `nthr` threads
// (one per warp) into disjoint accumulators
template <int N>
__global__ void __launch_bounds__(128) k_smalln_tight(long long* cycles, int iters, int a_mn, int nthr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (int)(64 * 1024 / 4); i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 255);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  long long t0 = clock64();
  if ((tid & 31) == 0 && warp < nthr) {
    const uint32_t idesc = make_idesc(128, N, a_mn);
    const uint64_t hiA = a_mn ? make_desc(0, 128, 96 * 16) : make_desc(0, 2048, 128);
    const uint64_t bd = make_desc(smem_u32(smem) + 48 * 1024, N * 16, 128);
    uint64_t ad[8];
    for (int i = 0; i < 8; ++i) ad[i] = hiA | (uint64_t)((smem_u32(smem) >> 4) + i * 8);
    const uint32_t d0 = tm + warp * 128;
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_bf16(d0 + (i & 1) * N, ad[i], bd, idesc, 1);
    }
    mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

template <int N>
static void run_smalln_tight(int a_mn, int nthr) {
  long long* d; CK(cudaMalloc(&d, 8 * 148));
  auto kern = k_smalln_tight<N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int iters = 8000;
  kern<<<148, 128, 64 * 1024>>>(d, iters, a_mn, nthr);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(148);
  CK(cudaMemcpy(h.data(), d, 8 * 148, cudaMemcpyDeviceToHost));
  double mean = 0; for (auto c : h) mean += c; mean /= 148;
  printf("  tight: M=128 N=%-3d K=16 A %s, %d issuing thread(s): %.1f cyc per MMA overall (floor %.0f)\n", N,
         a_mn ? "MN-major" : "K-major ", nthr, mean / (iters * nthr), 128.0 * N / 256);
  cudaFree(d);
}

// ---------------------------------------------------------------- E: MUFU.COS + FFMA2 snake throughput from registers
__global__ void __launch_bounds__(512) k_snake(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f;
  const float a2 = 1.9f, nhb = -0.45f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(nhb, __cosf(a2 * a[i]), a[i]);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  printf("[A] MN-major A operand, start-address K shifts (variant 0: LBO=128 (K groups), SBO=group stride)\n");
  bool ok = true;
  for (int shift : {0, 1, 8, 13, 80}) ok &= run_desc<0, 16>(shift, 96, 0);
  ok &= run_desc<0, 16>(5, 89, 0);
  if (!ok) {
    printf("  variant 0 failed; trying swapped LBO/SBO\n");
    for (int shift : {0, 8}) run_desc<0, 16>(shift, 96, 1);
  }
  printf("[B] K-major A with LBO = 2048\n");
  bool okb = true;
  for (int shift : {0, 1, 4, 14}) okb &= run_desc<1, 48>(shift, 22, 0);
  printf("descriptor probe: %s\n", (ok && okb) ? "ALL PASS" : "SOME FAIL");
  printf("[C] tcgen05.ld throughput\n");
  for (int nw : {4, 8, 16}) { run_ldtm<16>(nw); run_ldtm<32>(nw); }
  printf("[D] small-N MMA cadence, one issuing thread\n");
  run_smalln<16>(0); run_smalln<16>(1); run_smalln<32>(0); run_smalln<32>(1); run_smalln<64>(1);
  for (int nt : {1, 2, 4}) { run_smalln_tight<16>(1, nt); run_smalln_tight<16>(0, nt); run_smalln_tight<32>(1, nt); run_smalln_tight<64>(1, nt); run_smalln_tight<128>(0, nt); }
  printf("[E] snake from registers (FMUL + MUFU.COS + FFMA per value)\n");
  {
    float* d; CK(cudaMalloc(&d, 4 * 148 * 4 * 512));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_snake<<<148 * 4, 512>>>(d, 64); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_snake<<<148 * 4, 512>>>(d, 2048); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * 4 * 512 * 2048 * 16;
    printf("  snake: %.3f ms, %.1f values/clk/SM @1.9 GHz\n", ms, n / (ms * 1e-3) / 148 / 1.9e9);
  }
  return (ok && okb) ? 0 : 1;
}
