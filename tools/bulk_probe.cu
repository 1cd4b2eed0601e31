// bulk_probe.cu — how fast can one CTA per SM stream HBM into shared memory with cp.async.bulk (1-D) copies?
// 148 CTAs, each walks its own slice of a 4 GiB buffer with copies of S bytes, D of them in flight (ring of D slots,
// one mbarrier each, one issuing thread).  Prints GB/s for a sweep of (S, D) — the x ring of k_amp_tc is S = 5 KB,
// D = 9-12 (3 slots x 3-4 channel groups).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k_bulk(const uint8_t* src, size_t slice, int S, int D, int iters, long long* cyc, int NT) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  // NT issuing threads (one per warp), each with its own D / NT slots and every NT-th copy
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < NT) {
    const int me = threadIdx.x >> 5, Dm = D / NT;
    const uint8_t* p = src + (size_t)blockIdx.x * slice;
    long long t0 = clock64();
    for (int j = 0; j < iters / NT + Dm; ++j) {
      const int i = j * NT + me;
      const int sl = me * Dm + (j % Dm);
      if (j >= Dm) {                      // wait for the copy that used this slot
        const uint32_t ph = ((j / Dm) - 1) & 1;
        asm volatile("{\n\t.reg .pred P;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t@P bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(s32(&bar[sl])), "r"(ph) : "memory");
      }
      if (j < iters / NT) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[sl])), "r"(S) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(smem + (size_t)sl * S)), "l"((uint64_t)(p + ((size_t)i * S) % slice)), "r"(S), "r"(s32(&bar[sl])) : "memory");
      }
    }
    if (me == 0) cyc[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const size_t total = 4ull << 30;
  uint8_t* d; CK(cudaMalloc(&d, total)); CK(cudaMemset(d, 1, total));
  long long* dc; CK(cudaMalloc(&dc, 148 * 8));
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t slice = total / 148 / 65536 * 65536;
  int Ss[] = {2560, 5120, 16384}, Ds[] = {12, 24}, NTs[] = {1, 2, 4};
  for (int S : Ss)
    for (int D : Ds)
     for (int NT : NTs) {
      if ((size_t)S * D > 196 * 1024) continue;
      const int iters = (int)(slice / S);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_bulk<<<148, 128, (size_t)S * D>>>(d, slice, S, D, iters / 4, dc, NT);       // warm
      cudaEventRecord(e0);
      k_bulk<<<148, 128, (size_t)S * D>>>(d, slice, S, D, iters, dc, NT);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("S = %5d B  D = %2d in flight (%6.1f KB/SM), %d issuing thread(s): %7.1f GB/s  (%.2f us per copy per SM)\n", S, D, S * D / 1024.0, NT,
             (double)iters * S * 148 / ms / 1e6, ms * 1e3 / iters);
    }
  return 0;
}
