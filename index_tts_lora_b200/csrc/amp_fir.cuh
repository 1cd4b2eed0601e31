// amp_fir.cuh — the fused AMPBlock1 layer for the narrow stages (C <= 96): the 2x kaiser-sinc
// up-sampling FIR of Activation1d runs on the tensor cores, SnakeBeta and the down-sampling FIR in
// registers, the dilated Conv1d as the same implicit GEMM as k_amp_tc (amp_tc.cuh).
//
// Same contract as k_amp_tc<L, true>:  xt = conv_{k,d}(Activation1d(x)) [+ resid] [+ sum] [/div]
// (indextts/BigVGAN/models.py:65-74, alias_free_torch/act.py:9-29, resample.py:10-49, filter.py:60-96).
//
// Why: with lane = (channel pair, row run) the FIR/snake stage of k_amp_tc costs ~35 issue slots per
// element and bounds the narrow stages.  Here a TMEM lane is ONE channel of one time
// segment and its columns are consecutive time samples, so
//   * the up-sampling FIR is a banded-Toeplitz GEMM  D1[lane, 16 u-samples] = X[lane, 16 x-rows] * UP:
//     A = the TMA-staged x tile read MN-major (channels contiguous, time = K; a block's K window is a
//     start-address offset), B = [UP_hi | UP_lo] bf16 splits of the fp32 taps (N = 32, one MMA per block);
//   * SnakeBeta needs no per-element parameter traffic (one channel per thread);
//   * the down-sampling FIR runs in registers with NO cross-lane traffic (the thread holds the time
//     series), as packed FFMA2 over naturally aligned sample pairs (two accumulator parities).
// Tile = 256 output rows x all C_out columns of one utterance; the 256 + 2*hc activated rows of a
// 32-channel chunk are 4 time segments of S rows x 4 channel groups = 16 row groups = 128 TMEM lanes.
// Segment q is handled by the warps with (warp & 3) == q; a set of 4 such warps owns a whole chunk and
// NSETS chunks are in flight.  A segment is walked in NUB = S/8 + 1 blocks of 8 x-rows; block bi
// completes z rows [8(bi-1), 8bi) of the segment.
//
//   warps 0-11  activation sets     tcgen05.ld D1 -> hi+lo -> snake -> down FIR -> bf16 z rows (UMMA A layout)
//   warp 12     TMA producer        16 boxes {8 ch, 96 rows} per chunk -> x ring
//   warp 13     weight producer     (as k_amp_tc)
//   warp 14     FIR MMA issuer      one N=32 MMA per (set, block), D1 slots handed over by mbarriers
//   warp 15     conv MMA issuer     (as k_amp_tc) on the z ring
//   warps 16-19 epilogue            (shared with k_amp_tc)
// Sequence edges (replicate clamps, activations.py / filter.py) are re-evaluated exactly for the <= 12
// affected rows per utterance by a scalar path; rows outside [0, T) are the conv's zero padding.
#pragma once
#include "amp_tc.cuh"

namespace bvg {
namespace fir {
using namespace tc;

constexpr int NSETS = 3;
constexpr int NW_ACT = 4 * NSETS;
constexpr int WARP_X = NW_ACT, WARP_W = NW_ACT + 1, WARP_FIR = NW_ACT + 2, WARP_CONV = NW_ACT + 3, WARP_EPI = NW_ACT + 4;
constexpr int NTHREADS_F = (NW_ACT + 8) * 32;   // 640
constexpr int NXF = 4, NZF = 4;          // x / z ring depths
constexpr int XB = 96;                   // TMA box rows per (segment, channel group): S + 16 <= 96
constexpr int X_SLOT = 16 * XB * 16;     // 24576
constexpr int ZRF = 322;                 // z rows per channel group: >= 4*80, = 2 (mod 8) -> conflict-free 2-byte stores
constexpr int Z_SLOT = 4 * ZRF * 16;     // 20608
constexpr int W_STAGES_F = 2;
constexpr int TM_D1 = 320;               // TMEM: [0, 320) conv accumulators, [320, 512) NSETS x 2 D1 slots of 32 columns
constexpr int MAX_NTILE_F = 160;

constexpr int FOFF_BIAS = 0;
constexpr int FOFF_PREFIX = FOFF_BIAS + 2 * 256 * 4;
constexpr int FOFF_BAR = FOFF_PREFIX + (MAX_B + 8) * 4;
constexpr int F_NUM_BARS = 2 * NXF + 4 * NSETS + 2 * NZF + 2 * W_STAGES_F + 4;
constexpr int FOFF_TMEM = FOFF_BAR + F_NUM_BARS * 8;
constexpr int FOFF_UPB = (FOFF_TMEM + 16 + 127) / 128 * 128;   // [2][32][8] bf16 Toeplitz taps [hi | lo]
constexpr int FOFF_X = FOFF_UPB + 1024;
constexpr int FOFF_Z = FOFF_X + NXF * X_SLOT;
constexpr int FOFF_W = FOFF_Z + NZF * Z_SLOT;
constexpr int F_SMEM = FOFF_W + W_STAGES_F * W_STAGE_BYTES;
static_assert(F_SMEM <= 227 * 1024, "k_amp_fir shared-memory plan exceeds 227 KB");
static_assert(FOFF_X % 128 == 0 && X_SLOT % 128 == 0 && Z_SLOT % 128 == 0, "slot alignment");

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// exact re-evaluation of one activated sample z[m] of one channel at a sequence edge (scalar fp32):
//   z[m] = hb + sum_j dn[j] * s'[clamp(2m+j-5, 0, 2T-1)],  s'[n] = u[n] + nhb*cos(a2*u[n]),
//   u[n] = sum_i up2[..] * x[clamp(., 0, T-1)]      (resample.py:27-31 replicate pad, filter.py:87-93)
// xrow0 = address of this channel's bf16 sample at box row 0, box row r <-> time tbox0 + r.
__device__ __noinline__ float fir_edge_z(const uint8_t* xrow0, int tbox0, int m, int T, float a2, float nhb,
                                         const TcArgs& a) {
  float z = -nhb;
#pragma unroll 1
  for (int j = 0; j < 12; ++j) {
    int n = 2 * m + j - 5;
    n = n < 0 ? 0 : (n > 2 * T - 1 ? 2 * T - 1 : n);
    const int qn = n >> 1;
    const bool odd = n & 1;
    const int base = odd ? qn - 2 : qn - 3;
    float u = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      int t = base + i;
      t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
      const float xv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(xrow0 + (t - tbox0) * 16));
      u = fmaf(odd ? a.up2[10 - 2 * i] : a.up2[11 - 2 * i], xv, u);
    }
    z = fmaf(a.dn[j], fmaf(nhb, __cosf(a2 * u), u), z);
  }
  return z;
}

template <int NUB>
__global__ void __launch_bounds__(NTHREADS_F, 1)
k_amp_fir(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int S = 8 * (NUB - 1);           // z rows per segment
  static_assert(S + 16 <= XB && 4 * S <= ZRF, "segment geometry");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = a.n_tile;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + FOFF_BAR;
  auto BAR_XFULL = [&](int i) { return bar0 + 8 * i; };
  auto BAR_XEMPTY = [&](int i) { return bar0 + 8 * (NXF + i); };
  auto BAR_DFULL = [&](int i) { return bar0 + 8 * (2 * NXF + i); };                 // i = set*2 + slot
  auto BAR_DEMPTY = [&](int i) { return bar0 + 8 * (2 * NXF + 2 * NSETS + i); };
  auto BAR_ZFULL = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + i); };
  auto BAR_ZEMPTY = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + NZF + i); };
  auto BAR_WFULL = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + 2 * NZF + i); };
  auto BAR_WEMPTY = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + 2 * NZF + W_STAGES_F + i); };
  auto BAR_ACCFULL = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + 2 * NZF + 2 * W_STAGES_F + i); };
  auto BAR_ACCEMPTY = [&](int i) { return bar0 + 8 * (2 * NXF + 4 * NSETS + 2 * NZF + 2 * W_STAGES_F + 2 + i); };
  float* bias_s = reinterpret_cast<float*>(smem + FOFF_BIAS);
  int* prefix = reinterpret_cast<int*>(smem + FOFF_PREFIX);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FOFF_TMEM);

  const int hc = a.dil * (a.K - 1) / 2;
  const int NCH = (a.Cin + KC - 1) / KC;
  const int tile_bytes = n_tile * 64;
  const int tps = a.taps_per_stage;
  const int spc = (a.K + tps - 1) / tps;
  const int nacc = (4 * n_tile <= TM_D1) ? 2 : 1;

  // ---- prologue: tile prefix table, barriers, Toeplitz taps, TMEM
  if (warp == 0) {
    int run = 0;
    for (int b0 = 0; b0 < a.B; b0 += 32) {
      const int b = b0 + lane;
      int inc = 0;
      if (b < a.B) {
        const int Tin = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
        inc = (Tin + M_TILE - 1) / M_TILE;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (b < a.B) prefix[b + 1] = run + inc;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) prefix[0] = 0;
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NXF; ++i) { mbar_init(BAR_XFULL(i), 1); mbar_init(BAR_XEMPTY(i), 5); }
    for (int i = 0; i < 2 * NSETS; ++i) { mbar_init(BAR_DFULL(i), 1); mbar_init(BAR_DEMPTY(i), 4); }
    for (int i = 0; i < NZF; ++i) { mbar_init(BAR_ZFULL(i), 4); mbar_init(BAR_ZEMPTY(i), 1); }
    for (int i = 0; i < W_STAGES_F; ++i) { mbar_init(BAR_WFULL(i), 1); mbar_init(BAR_WEMPTY(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR_ACCFULL(i), 1); mbar_init(BAR_ACCEMPTY(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
  }
  if (warp >= 2 && warp < 6) {
    // UP[k][cc]: u(block sample cc) = sum_k UP[k][cc] * x(box row 8*bi + k);  cc = 2i: taps up2[11-2m] at k = i+m,
    // cc = 2i+1: taps up2[10-2m] at k = i+1+m (resample.py:19-31 polyphase form, gain folded into up2).
    // Columns: n < 8 -> even sample cc = 2n, n >= 8 -> odd sample cc = 2(n-8)+1; +16 = low-order bf16 split.
    __nv_bfloat16* upb = reinterpret_cast<__nv_bfloat16*>(smem + FOFF_UPB);
    for (int idx = (warp - 2) * 32 + lane; idx < 16 * 16; idx += 128) {
      const int k = idx >> 4, n = idx & 15;
      const int i = n & 7;
      const int m = (n < 8) ? k - i : k - i - 1;
      float v = 0.f;
      if (m >= 0 && m < 6) v = (n < 8) ? a.up2[11 - 2 * m] : a.up2[10 - 2 * m];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      upb[((k >> 3) * 32 + n) * 8 + (k & 7)] = hi;
      upb[((k >> 3) * 32 + 16 + n) * 8 + (k & 7)] = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == WARP_CONV) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(s_base + FOFF_TMEM), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total_tiles = prefix[a.B];
  const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_chunks = my_tiles * NCH;

  if (warp < NW_ACT) {
    // ===================== activation sets =====================
    reg_inc<120>();
    const int set = warp >> 2, q = warp & 3, g = lane >> 3, c8 = lane & 7;
    u64 dnp[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) dnp[j] = pk(a.dn[j], a.dn[j]);
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + TM_D1 + (uint32_t)(set * 64);
    TileCursor cur{prefix};
    int ks = 0;                                    // blocks consumed by this set (slot = ks & 1)
    for (int n = set; n < total_chunks; n += NSETS) {
      const int it = n / NCH, c = n - it * NCH;
      int b, t0, nt;
      cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
      const int T = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
      const int ts = t0 - hc + q * S;              // time of this segment's z row 0
      const int ch = c * KC + g * 8 + c8;
      const float a2f = __ldg(a.a2 + ch), nhbf = __ldg(a.nhb + ch);
      const u64 a2p = pk(a2f, a2f), nhbp = pk(nhbf, nhbf), hbp = pk(-nhbf, -nhbf);
      const bool edge = (ts - 8 < 0) || (ts + S + 8 > T);          // warp-uniform
      const int xs = n & (NXF - 1), zs = n & (NZF - 1);
      uint8_t* zrow0 = smem + FOFF_Z + zs * Z_SLOT + g * (ZRF * 16) + (q * S) * 16 + c8 * 2;
      u64 EP[7], OP[8];
#pragma unroll
      for (int i = 0; i < 7; ++i) EP[i] = 0ull;
#pragma unroll
      for (int i = 0; i < 8; ++i) OP[i] = 0ull;
#pragma unroll
      for (int bi = 0; bi < NUB; ++bi) {
        const int slot = ks & 1;
        mbar_wait(BAR_DFULL(set * 2 + slot), (ks >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tlane + (uint32_t)(slot * 32)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR_DEMPTY(set * 2 + slot));
        ++ks;
        // window shift: EP[t] pairs E[8bi-10+2t ..], OP[t] pairs O[8bi-12+2t ..]
        EP[0] = EP[4]; EP[1] = EP[5]; EP[2] = EP[6];
        OP[0] = OP[4]; OP[1] = OP[5]; OP[2] = OP[6]; OP[3] = OP[7];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          // u = hi + lo, s' = u + nhb*cos(a2*u)  (two consecutive even (p < 4) or odd (p >= 4) samples)
          const u64 u = add2(pk(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1])),
                             pk(__uint_as_float(v[16 + 2 * p]), __uint_as_float(v[16 + 2 * p + 1])));
          float t0f, t1f;
          upk(mul2(a2p, u), t0f, t1f);
          const u64 sv = fma2(nhbp, pk(__cosf(t0f), __cosf(t1f)), u);
          if (p < 4) EP[3 + p] = sv; else OP[4 + (p - 4)] = sv;
        }
        if (bi >= 1) {
          if (bi == 1) mbar_wait(BAR_ZEMPTY(zs), ((n / NZF) & 1) ^ 1);      // conv MMAs of this slot's previous chunk retired
          // rows rho = 8(bi-1) + 2a (+1): acc1[a] pairs (rho, rho+1), acc2[a] pairs (rho-1, rho)
          u64 acc1[4], acc2[5];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            u64 s1 = hbp;
#pragma unroll
            for (int i = 0; i < 6; i += 2) s1 = fma2(dnp[2 * i + 1], EP[aa + i / 2], s1);
#pragma unroll
            for (int i = 1; i < 6; i += 2) s1 = fma2(dnp[2 * i], OP[aa + (i + 1) / 2], s1);
            acc1[aa] = s1;
          }
#pragma unroll
          for (int aa = 0; aa < 5; ++aa) {
            u64 s2 = mul2(dnp[3], EP[aa]);
#pragma unroll
            for (int i = 3; i < 6; i += 2) s2 = fma2(dnp[2 * i + 1], EP[aa + (i - 1) / 2], s2);
#pragma unroll
            for (int i = 0; i < 6; i += 2) s2 = fma2(dnp[2 * i], OP[aa + i / 2], s2);
            acc2[aa] = s2;
          }
          float z[8];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            float l1, h1, l2, h2, l3, h3;
            upk(acc1[aa], l1, h1);
            upk(acc2[aa], l2, h2);
            upk(acc2[aa + 1], l3, h3);
            z[2 * aa] = l1 + h2;
            z[2 * aa + 1] = h1 + l3;
          }
          if (edge) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const int tm_ = ts + 8 * (bi - 1) + r;
              if (tm_ < 0 || tm_ >= T) z[r] = 0.f;             // conv zero padding (utils.py:59)
            }
          }
#pragma unroll
          for (int r = 0; r < 8; r += 2) {
            __nv_bfloat162 o = __floats2bfloat162_rn(z[r], z[r + 1]);
            *reinterpret_cast<__nv_bfloat16*>(zrow0 + (8 * (bi - 1) + r) * 16) = o.x;
            *reinterpret_cast<__nv_bfloat16*>(zrow0 + (8 * (bi - 1) + r + 1) * 16) = o.y;
          }
        }
      }
      if (edge) {
        // rows within 6 samples of a sequence end see the replicate clamps of the two FIRs: exact scalar redo
        __syncwarp();
        const int nlo = T < 6 ? T : 6;
        const uint8_t* xrow0 = smem + FOFF_X + xs * X_SLOT + (q * 4 + g) * (XB * 16) + c8 * 2;
#pragma unroll 1
        for (int e = 0; e < 12; ++e) {
          const int m = e < 6 ? e : T - 12 + e;
          const bool valid = e < 6 ? (m < nlo) : (m >= nlo);
          const int rho = m - ts;
          if (valid && rho >= 0 && rho < S) {
            const float zf = fir_edge_z(xrow0, ts - 7, m, T, a2f, nhbf, a);
            *reinterpret_cast<__nv_bfloat16*>(zrow0 + rho * 16) = __float2bfloat16_rn(zf);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // z stores -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR_ZFULL(zs));
        mbar_arrive(BAR_XEMPTY(xs));
      }
    }
  } else if (warp < WARP_EPI) {
    reg_dec<40>();
    if (warp == WARP_X) {
      // ===================== x producer (TMA): 16 boxes (segment, channel group) per chunk =====================
      if (lane == 0) {
        TileCursor cur{prefix};
        int b = 0, t0 = 0, nt;
        for (int n = 0; n < total_chunks; ++n) {
          const int it = n / NCH, c = n - it * NCH;
          if (c == 0) cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
          const int xs = n & (NXF - 1);
          mbar_wait_relaxed(BAR_XEMPTY(xs), ((n / NXF) & 1) ^ 1, 200);
          mbar_expect_tx(BAR_XFULL(xs), X_SLOT);
          const uint32_t dst = s_base + FOFF_X + xs * X_SLOT;
#pragma unroll 1
          for (int j = 0; j < 16; ++j)
            tma_load_4d(dst + j * (XB * 16), &tmx, 0, t0 - hc + (j >> 2) * S - 7, c * 4 + (j & 3), b, BAR_XFULL(xs));
        }
      }
    } else if (warp == WARP_W) {
      // ===================== weight producer (bulk copies) =====================
      if (lane == 0) {
        int stage = 0, phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wt);
          for (int c = 0; c < NCH; ++c)
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              const uint32_t bytes = (uint32_t)(taps * tile_bytes);
              mbar_wait_relaxed(BAR_WEMPTY(stage), phase ^ 1, 200);
              mbar_expect_tx(BAR_WFULL(stage), bytes);
              bulk_load(s_base + FOFF_W + stage * W_STAGE_BYTES, src, bytes, BAR_WFULL(stage));
              src += bytes;
              if (++stage == W_STAGES_F) { stage = 0; phase ^= 1; }
            }
        }
      }
    } else if (warp == WARP_FIR) {
      // ===================== FIR MMA issuer: D1[set][slot] = X(block) * [UP_hi | UP_lo] =====================
      if (lane == 0) {
        // A: MN-major SWIZZLE_NONE (LBO = stride between 8-row K groups, SBO = stride between 8-channel M groups)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(32 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        const u64 hiA = make_sdesc(0, 128, XB * 16);
        const u64 bdesc = make_sdesc(s_base + FOFF_UPB, 32 * 16, 128);
        int nn[NSETS], bb[NSETS], kk[NSETS];
#pragma unroll
        for (int s = 0; s < NSETS; ++s) { nn[s] = s; bb[s] = 0; kk[s] = 0; }
        int live = 0;
#pragma unroll
        for (int s = 0; s < NSETS; ++s) live += nn[s] < total_chunks;
        while (live > 0) {
          bool progressed = false;
#pragma unroll
          for (int s = 0; s < NSETS; ++s) {
            if (nn[s] >= total_chunks) continue;
            const int slot = kk[s] & 1;
            if (!mbar_test(BAR_DEMPTY(s * 2 + slot), ((kk[s] >> 1) & 1) ^ 1)) continue;
            const int xs = nn[s] & (NXF - 1);
            if (bb[s] == 0 && !mbar_test(BAR_XFULL(xs), (nn[s] / NXF) & 1)) continue;
            tc_fence_after();
            const uint32_t a0 = (s_base + FOFF_X + xs * X_SLOT + bb[s] * 128) >> 4;
            umma_bf16(tmem + TM_D1 + (uint32_t)((s * 2 + slot) * 32), hiA | a0, bdesc, idesc, 0u);
            umma_commit(BAR_DFULL(s * 2 + slot));
            ++kk[s];
            if (++bb[s] == NUB) {
              umma_commit(BAR_XEMPTY(xs));
              bb[s] = 0;
              nn[s] += NSETS;
              if (nn[s] >= total_chunks) --live;
            }
            progressed = true;
          }
          if (!progressed) asm volatile("nanosleep.u32 32;");
        }
      }
    } else {
      // ===================== conv MMA issuer (as k_amp_tc, A = z ring) =====================
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n_tile);
        const uint32_t lboA = ZRF * 16, lboB = (uint32_t)n_tile * 16;
        const u64 hiA = make_sdesc(0, lboA, 128), hiB = make_sdesc(0, lboB, 128);
        const uint32_t ksA = 2 * lboA / 16, ksB = 2 * lboB / 16, tileU = (uint32_t)tile_bytes / 16;
        int stage = 0, phase = 0, n = 0;
        for (int it = 0; it < my_tiles; ++it) {
          const int as = (nacc == 2) ? (it & 1) : 0;
          const int ause = (nacc == 2) ? (it >> 1) : it;
          mbar_wait(BAR_ACCEMPTY(as), (ause & 1) ^ 1);
          tc_fence_after();
          const uint32_t tm = tmem + (uint32_t)(as * 2 * n_tile);
          uint32_t accflag = 0;
          for (int c = 0; c < NCH; ++c, ++n) {
            const int zs = n & (NZF - 1);
            mbar_wait(BAR_ZFULL(zs), (n / NZF) & 1);
            tc_fence_after();
            const uint32_t aU = (s_base + FOFF_Z + zs * Z_SLOT) >> 4;
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              mbar_wait(BAR_WFULL(stage), phase);
              tc_fence_after();
              const uint32_t wU = (s_base + FOFF_W + stage * W_STAGE_BYTES) >> 4;
              for (int tj = 0; tj < taps; ++tj) {
                const uint32_t a0 = aU + (uint32_t)((s * tps + tj) * a.dil);
                const uint32_t b0 = wU + (uint32_t)tj * tileU;
                umma_bf16(tm, hiA | a0, hiB | b0, idesc, accflag);
                umma_bf16(tm, hiA | (a0 + ksA), hiB | (b0 + ksB), idesc, 1u);
                umma_bf16(tm + n_tile, hiA | (a0 + 128), hiB | b0, idesc, accflag);
                umma_bf16(tm + n_tile, hiA | (a0 + 128 + ksA), hiB | (b0 + ksB), idesc, 1u);
                accflag = 1u;
              }
              umma_commit(BAR_WEMPTY(stage));
              if (++stage == W_STAGES_F) { stage = 0; phase ^= 1; }
            }
            umma_commit(BAR_ZEMPTY(zs));
          }
          umma_commit(BAR_ACCFULL(as));
        }
      }
    }
  } else {
    // ===================== epilogue warps (shared with k_amp_tc) =====================
    reg_dec<64>();
    epilogue_role(a, bias_s, prefix, BAR_ACCFULL(0), BAR_ACCEMPTY(0), tmem, nacc, total_tiles, 0, warp & 3, lane,
                  threadIdx.x - WARP_EPI * 32);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_CONV) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace fir
}  // namespace bvg
