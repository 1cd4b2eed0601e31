// act_blk.cuh — Activation1d (2x kaiser-sinc FIR up -> SnakeBeta -> FIR down, alias_free_torch/act.py:24-29) as a
// stand-alone streaming kernel on the blocked bf16 layout, for the layers whose fused form would repeat it.
//
// Why it exists: k_amp_tc activates a 32-channel chunk once per (time tile, COLUMN tile).  C = 768 has three column
// tiles and C = 384 two, so the fused kernel runs every FIR / snake of those stages three / two times; those launches
// are bound by exactly that arithmetic (profiles/r01_launch_table_v5: the k = 3 layers of stage 0 reach 300 TFLOP/s).
// For such layers the plan writes z = Activation1d(x) ONCE to a scratch buffer (23 MB per layer at 16 x 10 s: it
// stays in the 126 MB L2) and k_amp_tc<ACT = false> consumes it through the same TMA / UMMA path as conv_pre.
//
// The arithmetic is act_run<9> of amp_tc.cuh, i.e. bit-identical z rows to the fused kernel.  Work unit = one warp:
// one 8-channel group x 66 time rows of one utterance (8 runs of 9 rows, 6 rows of overlap with the neighbouring
// units because the down FIR needs +-5/6 up-sampled neighbours).  Each warp stages its own 78 x rows with cp.async
// (double buffered: the next unit's rows are in flight while this one computes), writes z rows to its own staging
// tile and stores them as 16-byte rows (512 B per warp store).  No inter-warp synchronisation at all.
#pragma once
#include "amp_tc.cuh"

namespace bvg {
namespace tc {

constexpr int AB_L = 9;                     // rows per run (odd: the 4-byte smem accesses of a warp hit 32 banks)
constexpr int AB_V = 8 * AB_L - 6;          // 66 z rows per warp unit
constexpr int AB_XROWS = 8 * AB_L + 6;      // 78 x rows per unit: [vlo - 6, vlo + 72)
constexpr int AB_XALLOC = 96;               // 3 x 32 rows (one 16-byte cp.async per lane and pass)
constexpr int AB_ZROWS = 8 * AB_L;          // 72 staged z rows, rows [3, 69) are the unit's
constexpr int AB_WARPS = 8;
constexpr int AB_SMEM_PER_WARP = (2 * AB_XALLOC + AB_ZROWS) * 16;   // 4224 B

struct AbArgs {
  const __nv_bfloat16* x;      // blocked [B][groups][Tstride][8]
  __nv_bfloat16* z;            // same geometry
  const float* a2;             // [C] 2*exp(alpha)
  const float* nhb;            // [C] -0.5/(exp(beta)+1e-9)
  const int* lengths;          // frames per utterance, or null
  int B, groups, Tstride, rate, Tmax;
  int nslices;                 // ceil(Tstride / AB_V)
  float up2[12], dn[12];
};


// rows >= T of an utterance are the zero padding the consuming conv reads; they are written up to the end of the
// consumer's last 256-row tile plus its 32-row read-ahead
__device__ __forceinline__ int ab_zero_bound(int T, int Tstride) {
  return min(Tstride, (T + M_TILE - 1) / M_TILE * M_TILE + 64);
}

__global__ void __launch_bounds__(AB_WARPS * 32, 2) k_act_blk(const __grid_constant__ AbArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, p = lane & 3;
  uint8_t* my = smem + warp * AB_SMEM_PER_WARP;
  const uint32_t my_u = smem_u32(my);
  uint32_t* zs = reinterpret_cast<uint32_t*>(my + 2 * AB_XALLOC * 16);

  ActCtx k;
  load_taps(k, a);
  const int rowS = g * AB_L;                        // first staged z row of this lane's run
  uint32_t smask = 0;                               // bit r: staged row rowS + r is one of the unit's 66
#pragma unroll
  for (int r = 0; r < AB_L; ++r)
    if (rowS + r >= 3 && rowS + r < 3 + AB_V) smask |= 1u << r;

  const uint32_t per_b = (uint32_t)a.groups * (uint32_t)a.nslices;
  const uint32_t total = per_b * (uint32_t)a.B;          // host checks that this fits 31 bits
  const uint32_t stride = gridDim.x * AB_WARPS;

  // unit u -> (b, group, slice), slice fastest; decoded once, when its x rows are requested
  struct Unit { int b, grp, sl, T; };
  auto prefetch = [&](uint32_t u, int buf) -> Unit {
    Unit n{0, 0, 0, 0};
    if (u < total) {
      n.b = (int)(u / per_b);
      const uint32_t r = u - (uint32_t)n.b * per_b;
      n.grp = (int)(r / (uint32_t)a.nslices);
      n.sl = (int)(r - (uint32_t)n.grp * (uint32_t)a.nslices);
      n.T = a.lengths ? __ldg(a.lengths + n.b) * a.rate : a.Tmax;
      const int vlo = n.sl * AB_V;
      if (vlo < n.T) {                              // units past the end only write zeros
        const __nv_bfloat16* src = a.x + ((size_t)n.b * a.groups + n.grp) * (size_t)a.Tstride * 8;
        const int xlo = vlo - 6;
#pragma unroll
        for (int i = 0; i < AB_XALLOC / 32; ++i) {
          const int r2 = i * 32 + lane, t = xlo + r2;
          if (r2 < AB_XROWS && t >= 0 && t < a.Tstride)
            cp_async16(my_u + (uint32_t)(buf * AB_XALLOC + r2) * 16, src + (size_t)t * 8);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return n;
  };

  uint32_t u = blockIdx.x * AB_WARPS + warp;
  int buf = 0;
  Unit nxt = prefetch(u, 0);
  for (; u < total; u += stride, buf ^= 1) {
    const Unit cur = nxt;
    nxt = prefetch(u + stride, buf ^ 1);
    const int b = cur.b, grp = cur.grp, sl = cur.sl, T = cur.T;
    const int vlo = sl * AB_V;
    const int zend = ab_zero_bound(T, a.Tstride);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    if (vlo >= zend) continue;                      // warp-uniform
    __nv_bfloat16* dst = a.z + ((size_t)b * a.groups + grp) * (size_t)a.Tstride * 8;
    if (vlo < T) {
      const int ch = grp * 8 + 2 * p;
      const float2 a2v = __ldg(reinterpret_cast<const float2*>(a.a2 + ch));
      const float2 nhbv = __ldg(reinterpret_cast<const float2*>(a.nhb + ch));
      k.a2 = pk(a2v.x, a2v.y);
      k.nhb = pk(nhbv.x, nhbv.y);
      k.hb = pk(-nhbv.x, -nhbv.y);
      const uint32_t* xk = reinterpret_cast<const uint32_t*>(my + buf * AB_XALLOC * 16) + p;
      uint32_t* zk = zs + p;
      const int m0 = vlo - 3 + rowS;                // global time of this lane's first row
      const int xlo = vlo - 6;
      const bool edge = __any_sync(0xffffffffu, (m0 - 3 < 0) || (m0 + AB_L + 2 > T - 1));
      if (edge) act_run_edge<AB_L>(xk, zk, rowS, smask, m0, xlo, T, k.a2, k.nhb, a, lane);
      else act_run<AB_L, false>(xk, zk, rowS, smask, m0, xlo, T, k, lane);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int r = i * 32 + lane, t = vlo + r;   // unit row r is staged row r + 3
        if (r < AB_V && t < zend)
          *reinterpret_cast<uint4*>(dst + (size_t)t * 8) = *reinterpret_cast<const uint4*>(zs + (r + 3) * 4);
      }
      __syncwarp();                                 // staged rows are re-written by the next unit
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int r = i * 32 + lane, t = vlo + r;
        if (r < AB_V && t < zend) *reinterpret_cast<uint4*>(dst + (size_t)t * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace tc
}  // namespace bvg
