// amp_tc.cuh — the fused AMPBlock1 layer on tcgen05 / TMEM (BVG_PREC_BF16 path).
//
// One launch = one `xt = conv_k,d(Activation1d(x)) [+ resid]` step of AMPBlock1.forward
// (indextts/BigVGAN/models.py:65-74), or a plain dilated conv (ACT=false: conv_pre, :226-228).
//
// Data layout in HBM ("blocked channels"):  X[b][C/8][T][8] bf16 — groups of 8 channels
// innermost, time next.  Chosen because (1) a TMA box {8ch, rows} lands in shared memory as
// [rows][8] = 16-byte rows, which IS the SWIZZLE_NONE K-major UMMA core-matrix layout with
// SBO = 128 B, so rows are linear at 16 B and a conv tap shift of s samples is a start-address
// offset of 16*s bytes in the A descriptor (verified on B200 by tools/umma_probe.cu); (2) the
// epilogue thread that owns TMEM lane = time row writes 16-byte groups that are contiguous
// across the 32 lanes of a warp (512 B runs).
//
// CTA tile: 256 time rows (two M=128 accumulators) x n_tile output channels (<= 256), looping
// over C_in in chunks of 32 channels:
//   warp 8      TMA producer: x rows [t0-32, t0+288) of the chunk -> xbuf[2]   (8 boxes of {8,160})
//   warp 9      weight producer: per-(chunk,tap) bf16 tiles [4][n_tile][8] via cp.async.bulk -> ring[4]
//   warps 0-7   activation: x -> 2x kaiser-sinc FIR up -> SnakeBeta -> FIR down -> bf16 z tile in the
//               UMMA A layout (fp32 math on packed f32x2 registers, two channels per thread, one
//               run of L consecutive rows per thread); afterwards the same warps run the epilogue
//   warp 10     MMA issuer: for every tap j, M block, K step: tcgen05.mma.kind::f16
//               A = z tile rows (mb*128 + j*dil ..), B = weight tile, D = TMEM[mb*n_tile ..]
// Synchronisation is mbarrier-only inside the main loop.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace bvg {
namespace tc {

constexpr int M_TILE = 256;
constexpr int KC = 32;        // input channels per chunk (4 groups of 8)
constexpr int XR = 320;       // x rows staged per chunk: t0-32 .. t0+287
constexpr int X_LEAD = 32;
constexpr int BOXR = 160;     // TMA box rows (two boxes per channel group)
constexpr int ZR = 336;       // z rows allocated (16 runs x 21)
constexpr int NW_ACT = 8;
constexpr int W_STAGES = 4;
constexpr int W_STAGE_BYTES = 16384;
constexpr int NTHREADS = (NW_ACT + 3) * 32;

constexpr int X_BUF_BYTES = 4 * XR * 16;   // 20480
constexpr int Z_BUF_BYTES = 4 * ZR * 16;   // 21504
constexpr int OFF_X = 0;
constexpr int OFF_Z = OFF_X + 2 * X_BUF_BYTES;
constexpr int OFF_W = OFF_Z + 2 * Z_BUF_BYTES;
constexpr int OFF_BIAS = OFF_W + W_STAGES * W_STAGE_BYTES;
constexpr int OFF_BAR = OFF_BIAS + 256 * 4;
constexpr int NUM_BARS = 2 + 2 + 2 + 2 + 2 * W_STAGES + 1;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;

struct TcArgs {
  const __nv_bfloat16* wt;     // [ntile][chunk][tap][4][n_tile][8] bf16
  const float* bias;           // [Cout]
  const float* bias_b;         // [B][bias_b_stride] speaker-conditioning add, or null
  int bias_b_stride;
  const __nv_bfloat16* resid;  // blocked, or null
  const __nv_bfloat16* acc_in; // blocked, or null
  __nv_bfloat16* out;          // blocked [B][Cout/8][Tstride][8]
  float div;
  int Cin, Cout, K, dil, n_tile, taps_per_stage, tmem_cols;
  int Tstride;                 // rows per (b, channel group) in every activation buffer of this stage
  const int* lengths;
  int rate, Tmax;
  const float* a2;             // [Cin padded to 32]  2*exp(alpha)
  const float* nhb;            // [Cin padded to 32]  -0.5/(exp(beta)+1e-9)
  float up2[12];               // 2*f[k]  (the x2 gain of resample.py:30 folded in)
  float dn[12];
};

// ------------------------------------------------------------------------------ PTX helpers
typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ u64 make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // SWIZZLE_NONE K-major: SBO = stride between 8-row groups, LBO = stride between the two
  // 8-element K halves of a K=16 step (cute/arch/mma_sm100_desc.hpp:96-118; version bit 46).
  return (u64)((saddr & 0x3FFFF) >> 4) | ((u64)((lbo >> 4) & 0x3FFF) << 16) |
         ((u64)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
  // kind::f16: D=f32 (bit4), A=B=bf16 (bits 7,10), both K-major, N>>3 at 17, M>>4 at 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// packed fp32x2 arithmetic (sm_100: one FFMA2 issues two FMAs)
__device__ __forceinline__ u64 pk(float x, float y) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 bf2_to_f2(uint32_t w) {   // two bf16 -> two fp32 (exact)
  return pk(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// ------------------------------------------------------------------------------ activation
// s'(n) = y - hb*cos(2a*y) with y the 2x-upsampled sample n; the constant +hb of
// sin^2 = (1 - cos 2t)/2 is added once per output because the down taps sum to 1:
//   z[m] = hb + sum_k dn[k] * s'[clamp(2m+k-5)]      (activations.py:109-122, filter.py:87-96)
struct ActCtx {
  u64 upE[6], upO[6], dn[12];   // taps, broadcast to both halves
  u64 a2, nhb, hb;
};

__device__ __forceinline__ u64 snake_s(u64 y, const ActCtx& k) {
  float t0, t1;
  upk(mul2(k.a2, y), t0, t1);
  return fma2(k.nhb, pk(__cosf(t0), __cosf(t1)), y);
}

// s' at an absolute up-sampled index n in [0, 2T) straight from the staged x rows (edge tiles only)
__device__ __forceinline__ u64 snake_s_at(const uint32_t* xk, int n, int xlo, int T, const ActCtx& k) {
  const int q = n >> 1;
  const int base = (n & 1) ? q - 2 : q - 3;
  u64 y = 0ull;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    int t = base + i;
    t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    int r = t - xlo;
    r = r < 0 ? 0 : (r > XR - 1 ? XR - 1 : r);
    y = fma2((n & 1) ? k.upO[i] : k.upE[i], bf2_to_f2(xk[r * 4]), y);
  }
  return snake_s(y, k);
}

// One thread: channel pair, L consecutive z rows starting at local row `row0` (global time m0).
template <int L, bool EDGE>
__device__ __forceinline__ void act_run(const uint32_t* __restrict__ xk, uint32_t* __restrict__ zk, int row0,
                                        int m0, int xlo, int T, const ActCtx& k) {
  u64 xw[L + 10];
  u64 sv[2 * L + 10];
  const int nb = 2 * m0 - 5;
  u64 s_first = 0ull, s_last = 0ull;
  if (EDGE) {
    s_first = snake_s_at(xk, 0, xlo, T, k);
    s_last = snake_s_at(xk, 2 * T - 1, xlo, T, k);
  }
  auto load_x = [&](int i) {
    int t = m0 - 5 + i;
    if (EDGE) t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    int r = t - xlo;
    r = r < 0 ? 0 : (r > XR - 1 ? XR - 1 : r);
    return bf2_to_f2(xk[r * 4]);
  };
  auto make_s = [&](int u) {          // u = n - nb; compile-time after unrolling
    u64 y;
    if (u & 1) {                       // n even: taps f[11-2i] on x[q-3+i]
      const int i0 = (u - 1) / 2;
      y = mul2(k.upE[0], xw[i0]);
#pragma unroll
      for (int i = 1; i < 6; ++i) y = fma2(k.upE[i], xw[i0 + i], y);
    } else {                           // n odd: taps f[10-2i] on x[q-2+i]
      const int i0 = u / 2;
      y = mul2(k.upO[0], xw[i0]);
#pragma unroll
      for (int i = 1; i < 6; ++i) y = fma2(k.upO[i], xw[i0 + i], y);
    }
    u64 s = snake_s(y, k);
    if (EDGE) {
      const int n = nb + u;
      if (n < 0) s = s_first;
      else if (n > 2 * T - 1) s = s_last;
    }
    return s;
  };
#pragma unroll
  for (int i = 0; i < 10; ++i) xw[i] = load_x(i);
#pragma unroll
  for (int u = 0; u < 10; ++u) {
    sv[u] = make_s(u);
  }
#pragma unroll
  for (int r = 0; r < L; ++r) {
    xw[r + 10] = load_x(r + 10);
    sv[2 * r + 10] = make_s(2 * r + 10);
    sv[2 * r + 11] = make_s(2 * r + 11);
    u64 z = k.hb;
#pragma unroll
    for (int j = 0; j < 12; ++j) z = fma2(k.dn[j], sv[2 * r + j], z);
    float z0, z1;
    upk(z, z0, z1);
    const int m = m0 + r;
    if (m < 0 || m >= T) { z0 = 0.f; z1 = 0.f; }      // conv zero padding (utils.py:59)
    __nv_bfloat162 o = __floats2bfloat162_rn(z0, z1);
    zk[(row0 + r) * 4] = *reinterpret_cast<uint32_t*>(&o);
  }
}

// ------------------------------------------------------------------------------ the kernel
template <int L, bool ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_amp_tc(const __grid_constant__ CUtensorMap tmx, const TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, nt = blockIdx.y, t0 = blockIdx.x * M_TILE;
  const int T = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
  const int n_tile = a.n_tile;
  const int cg_total = a.Cout >> 3;

  if (t0 >= T) {
    // one tile of zeros past the end so that un-activated consumers (ConvTranspose1d, plain
    // convs) see the conv zero padding; tiles further out are never read.
    if (t0 < T + M_TILE) {
      const int rows = min(M_TILE, a.Tmax - t0);
      const int ngrp = min(n_tile >> 3, cg_total - nt * (n_tile >> 3));
      for (int i = threadIdx.x; i < rows * ngrp; i += NTHREADS) {
        const int g = i / rows, r = i % rows;
        uint4* o = reinterpret_cast<uint4*>(
            a.out + (((size_t)b * cg_total + nt * (n_tile >> 3) + g) * a.Tstride + t0 + r) * 8);
        *o = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + OFF_BAR;
  auto BAR_XFULL = [&](int i) { return bar0 + 8 * (0 + i); };
  auto BAR_XEMPTY = [&](int i) { return bar0 + 8 * (2 + i); };
  auto BAR_ZFULL = [&](int i) { return bar0 + 8 * (4 + i); };
  auto BAR_ZEMPTY = [&](int i) { return bar0 + 8 * (6 + i); };
  auto BAR_WFULL = [&](int i) { return bar0 + 8 * (8 + i); };
  auto BAR_WEMPTY = [&](int i) { return bar0 + 8 * (8 + W_STAGES + i); };
  const uint32_t BAR_ACC = bar0 + 8 * (8 + 2 * W_STAGES);
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int hc = a.dil * (a.K - 1) / 2;
  const int NCH = (a.Cin + KC - 1) / KC;
  const int tile_bytes = n_tile * 64;
  const int tps = a.taps_per_stage;
  const int spc = (a.K + tps - 1) / tps;

  if (warp == NW_ACT && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR_XFULL(i), 1);
      mbar_init(BAR_XEMPTY(i), ACT ? NW_ACT : 1);
      mbar_init(BAR_ZFULL(i), NW_ACT);
      mbar_init(BAR_ZEMPTY(i), 1);
    }
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(BAR_WFULL(i), 1); mbar_init(BAR_WEMPTY(i), 1); }
    mbar_init(BAR_ACC, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
  }
  if (warp == NW_ACT + 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(s_base + OFF_TMEM), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < NW_ACT) {
    for (int i = threadIdx.x; i < n_tile; i += NW_ACT * 32) {
      const int co = nt * n_tile + i;
      float v = 0.f;
      if (co < a.Cout) {
        v = __ldg(a.bias + co);
        if (a.bias_b) v += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
      }
      bias_s[i] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == NW_ACT) {
    // ===================== x producer (TMA) =====================
    if (lane == 0) {
      for (int c = 0; c < NCH; ++c) {
        const int buf = c & 1, use = c >> 1;
        mbar_wait(BAR_XEMPTY(buf), (use & 1) ^ 1);
        mbar_expect_tx(BAR_XFULL(buf), X_BUF_BYTES);
        const uint32_t dst = s_base + OFF_X + buf * X_BUF_BYTES;
#pragma unroll
        for (int kg = 0; kg < 4; ++kg)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_4d(dst + kg * (XR * 16) + h * (BOXR * 16), &tmx, 0, t0 - X_LEAD + h * BOXR, c * 4 + kg, b,
                        BAR_XFULL(buf));
      }
    }
  } else if (warp == NW_ACT + 1) {
    // ===================== weight producer (bulk copies) =====================
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wt) + (size_t)nt * NCH * a.K * tile_bytes;
      int stage = 0, phase = 0;
      for (int c = 0; c < NCH; ++c)
        for (int s = 0; s < spc; ++s) {
          const int taps = min(tps, a.K - s * tps);
          const uint32_t bytes = (uint32_t)(taps * tile_bytes);
          mbar_wait(BAR_WEMPTY(stage), phase ^ 1);
          mbar_expect_tx(BAR_WFULL(stage), bytes);
          bulk_load(s_base + OFF_W + stage * W_STAGE_BYTES, src, bytes, BAR_WFULL(stage));
          src += bytes;
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == NW_ACT + 2) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, n_tile);
      const uint32_t lboA = (ACT ? ZR : XR) * 16;
      const uint32_t lboB = (uint32_t)n_tile * 16;
      int stage = 0, phase = 0;
      for (int c = 0; c < NCH; ++c) {
        const int buf = c & 1, use = c >> 1;
        mbar_wait(ACT ? BAR_ZFULL(buf) : BAR_XFULL(buf), use & 1);
        tc_fence_after();
        const uint32_t abase = ACT ? (s_base + OFF_Z + buf * Z_BUF_BYTES)
                                   : (s_base + OFF_X + buf * X_BUF_BYTES + (X_LEAD - hc) * 16);
        for (int s = 0; s < spc; ++s) {
          const int taps = min(tps, a.K - s * tps);
          mbar_wait(BAR_WFULL(stage), phase);
          tc_fence_after();
          const uint32_t wbase = s_base + OFF_W + stage * W_STAGE_BYTES;
          for (int tj = 0; tj < taps; ++tj) {
            const int j = s * tps + tj;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const u64 ad = make_sdesc(abase + (mb * 128 + j * a.dil) * 16 + ks * 2 * lboA, lboA, 128);
                const u64 bd = make_sdesc(wbase + tj * tile_bytes + ks * 2 * lboB, lboB, 128);
                umma_bf16(tmem + mb * n_tile, ad, bd, idesc, (c | j | ks) != 0);
              }
          }
          umma_commit(BAR_WEMPTY(stage));          // weight stage free once these MMAs retire
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(ACT ? BAR_ZEMPTY(buf) : BAR_XEMPTY(buf));
      }
      umma_commit(BAR_ACC);
    }
  } else {
    // ===================== activation warps, then epilogue =====================
    if (ACT) {
      ActCtx k;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        k.upE[i] = pk(a.up2[11 - 2 * i], a.up2[11 - 2 * i]);
        k.upO[i] = pk(a.up2[10 - 2 * i], a.up2[10 - 2 * i]);
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) k.dn[i] = pk(a.dn[i], a.dn[i]);
      const int kg = warp & 3, half = warp >> 2, g = lane >> 2, p = lane & 3;
      const int row0 = (half * 8 + g) * L;
      const int ZW = M_TILE + 2 * hc;
      const int m0 = t0 - hc + row0;
      const int xlo = t0 - X_LEAD;
      // runs whose x window [m0-5, m0+L+4] leaves [0, T) need the replicate clamps
      const bool edge = (m0 - 5 < 0) || (m0 + L + 4 > T - 1);
      for (int c = 0; c < NCH; ++c) {
        const int buf = c & 1, use = c >> 1;
        const int ch = c * KC + kg * 8 + 2 * p;
        const float2 a2v = __ldg(reinterpret_cast<const float2*>(a.a2 + ch));
        const float2 nhbv = __ldg(reinterpret_cast<const float2*>(a.nhb + ch));
        k.a2 = pk(a2v.x, a2v.y);
        k.nhb = pk(nhbv.x, nhbv.y);
        k.hb = pk(-nhbv.x, -nhbv.y);
        mbar_wait(BAR_XFULL(buf), use & 1);
        mbar_wait(BAR_ZEMPTY(buf), (use & 1) ^ 1);
        if (row0 < ZW) {
          const uint32_t* xk = reinterpret_cast<const uint32_t*>(smem + OFF_X + buf * X_BUF_BYTES) + kg * (XR * 4) + p;
          uint32_t* zk = reinterpret_cast<uint32_t*>(smem + OFF_Z + buf * Z_BUF_BYTES) + kg * (ZR * 4) + p;
          if (edge) act_run<L, true>(xk, zk, row0, m0, xlo, T, k);
          else act_run<L, false>(xk, zk, row0, m0, xlo, T, k);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // z stores -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR_ZFULL(buf));
          mbar_arrive(BAR_XEMPTY(buf));
        }
      }
    }
    // ---- epilogue: TMEM -> registers -> (+bias, +resid, +running sum, /div) -> bf16 -> HBM
    mbar_wait(BAR_ACC, 0);
    tc_fence_after();
    const int q = warp & 3, h = warp >> 2;
    const int t = t0 + h * 128 + q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * n_tile);
    for (int c0 = 0; c0 < n_tile; c0 += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int cg = nt * (n_tile >> 3) + (c0 >> 3) + kk;       // global channel group
        if (cg >= cg_total || t >= a.Tmax) continue;
        const size_t idx = (((size_t)b * cg_total + cg) * a.Tstride + t) * 8;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (t < T) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[kk * 8 + e]) + bias_s[c0 + kk * 8 + e];
          if (a.resid) {
            const uint4 r = *reinterpret_cast<const uint4*>(a.resid + idx);
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[2 * e] += __uint_as_float(rw[e] << 16);
              f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
            }
          }
          if (a.acc_in) {
            const uint4 r = *reinterpret_cast<const uint4*>(a.acc_in + idx);
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[2 * e] += __uint_as_float(rw[e] << 16);
              f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
            }
          }
          if (a.div != 1.0f) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = f[e] / a.div;
          }
          __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
          o = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                         *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
        }
        *reinterpret_cast<uint4*>(a.out + idx) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW_ACT + 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols) : "memory");
  }
}

}  // namespace tc
}  // namespace bvg
