// amp_fir.cuh — the fused AMPBlock1 layer for the narrow stages (C <= 96): BOTH kaiser-sinc FIRs of
// Activation1d run on the tensor cores, only SnakeBeta stays on the CUDA cores; the dilated Conv1d is the
// same implicit GEMM as k_amp_tc (amp_tc.cuh).
//
// Same contract as k_amp_tc<L, true>:  xt = conv_{k,d}(Activation1d(x)) [+ resid] [+ sum] [/div]
// (indextts/BigVGAN/models.py:65-74, alias_free_torch/act.py:9-29, resample.py:10-49, filter.py:60-96).
//
// Why: in k_amp_tc the two 12-tap FIRs cost ~24 FMA-pipe lane-ops per element and bound the narrow stages
// (DESIGN.md §6).  Here a TMEM lane is ONE channel of one time segment and TMEM columns are consecutive
// time samples, so both FIRs are banded-Toeplitz GEMMs with a tiny constant B operand:
//   up   D1[lane, 16 u] = X[lane, 16 x-rows] * UP     A = the TMA-staged x tile read MN-major from shared
//        memory (channels contiguous, time = K; a block's K window is a start-address offset), B = bf16 hi
//        and lo splits of the fp32 taps accumulated by two MMAs (exact to 2^-17 of a tap);
//   down D2[lane, 16 z] = S[lane, 48 s] * DN          A = the snake output s, written to TMEM by the
//        activation warps as fp16 pairs (tcgen05.st) and read by the MMA straight from TMEM, B = fp16 taps.
// The activation warps only do tcgen05.ld -> s = u + nhb*cos(a2*u) -> fp16 -> tcgen05.st, and later
// tcgen05.ld -> +hb -> bf16 -> z tile (the UMMA A operand of the conv).
// Tile = 256 output rows x all C_out columns of one utterance; the 256 + 2*hc activated rows of a
// 32-channel chunk are 4 time segments of S rows x 4 channel groups = 16 row groups = 128 TMEM lanes.
// Segment q is handled by the 4 warps with (warp & 3) == q, which split the columns.  Hand-overs between the
// activation warps and the MMA issuers cost ~1-2 k cycles per round trip, so the granularity is a HALF chunk:
// a segment's NUB = S/8 + 1 u-blocks of 16 s samples are issued as two batches into two D1 buffers, and the
// z-blocks (16 rows from r0 = min(16 zi, S-16), reading s columns [2 r0, 2 r0 + 48)) as two batches into two D2
// buffers; the second batch of a chunk is picked up while the next chunk's first half is in flight.
//
//   warps 0-15  activation          tcgen05.ld D1 -> snake -> fp16 -> tcgen05.st s;  tcgen05.ld D2 -> +hb -> bf16 z tile
//   warp 16     lane 0: TMA producer, 16 boxes {8 ch, 96 rows} per chunk -> x ring; lane 1: weight producer (as k_amp_tc)
//   warp 17     conv MMA issuer     (as k_amp_tc) on the z ring
//   warp 18     up-FIR MMA issuer   two MMAs (hi, lo taps) per u-block, one commit per half chunk
//   warp 19     down-FIR MMA issuer three TS MMAs per z-block, one commit per half chunk
//   warps 20-23 epilogue            lean conv-mode epilogue
// Sequence edges (replicate clamps, activations.py / filter.py) are re-evaluated exactly for the <= 12
// affected rows per utterance by a scalar path; rows outside [0, T) are the conv's zero padding.
#pragma once
#include "amp_tc.cuh"

namespace bvg {
namespace fir {
using namespace tc;

constexpr int NW_ACT = 16;               // 4 TMEM lane quarters (= time segments) x 4 warps sharing a quarter's columns
constexpr int WARP_XW = NW_ACT, WARP_CONV = NW_ACT + 1, WARP_UP = NW_ACT + 2, WARP_DN = NW_ACT + 3, WARP_EPI = NW_ACT + 4;
constexpr int NTHREADS_F = (WARP_EPI + 4) * 32;      // 768
constexpr int NXF = 4, NZF = 4;          // x / z ring depths
constexpr int XB = 96;                   // TMA box rows per (segment, channel group): S + 16 <= 96
constexpr int X_SLOT = 16 * XB * 16;     // 24576
constexpr int ZRF = 322;                 // z rows per channel group: >= 4*80, = 2 (mod 8) -> conflict-free 2-byte stores
constexpr int Z_SLOT = 4 * ZRF * 16;     // 20608
constexpr int W_STAGES_F = 2;
constexpr int MAX_NTILE_F = 64;
// TMEM columns: [0,128) conv accumulators | [128,320) two D1 half-chunk buffers of 96 | [320,416) s (fp16 pairs) |
// [416,512) two D2 half-chunk buffers of 48
constexpr int TM_ACC = 128, TM_D1 = 128, TM_S = 320, TM_D2 = 416;

constexpr int FOFF_BIAS = 0;
constexpr int FOFF_PREFIX = FOFF_BIAS + 2 * 256 * 4;
constexpr int FOFF_BAR = FOFF_PREFIX + (MAX_B + 8) * 4;
constexpr int F_NUM_BARS = 2 * NXF + 10 + 2 * NZF + 2 * W_STAGES_F + 4;
constexpr int FOFF_TMEM = FOFF_BAR + F_NUM_BARS * 8;
constexpr int FOFF_UPB = (FOFF_TMEM + 16 + 127) / 128 * 128;   // 2 x [2][16][8] bf16 up taps (hi, lo), then [6][16][8] fp16 down taps
constexpr int FOFF_DNB = FOFF_UPB + 1024;
constexpr int FOFF_X = FOFF_DNB + 1536;
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

__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, u64 bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_f16_e(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, u64 bdesc, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ uint32_t f2_to_h2_sat(float lo, float hi) {     // {hi, lo} -> packed fp16x2, finite-saturating
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int NUB>
__global__ void __launch_bounds__(NTHREADS_F, 1)
k_amp_fir(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int S = 8 * (NUB - 1);           // z rows per segment
  constexpr int NZB = (S + 15) / 16;         // z-blocks per segment (the last one may overlap its predecessor)
  constexpr int NB0 = (NUB + 1) / 2, NB1 = NUB - NB0;   // u-blocks per half chunk
  constexpr int NZ0 = 2, NZ1 = NZB - 2;                 // z-blocks per half chunk: windows [2 r0, 2 r0 + 48) inside the half's s
  static_assert(S + 16 <= XB && 4 * S <= ZRF && NB0 * 16 <= 96 && NZ1 * 16 <= 48 && 48 <= 16 * NB0 && NZ1 >= 1, "segment geometry");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = a.n_tile;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + FOFF_BAR;
  auto BAR_XFULL = [&](int i) { return bar0 + 8 * i; };
  auto BAR_XEMPTY = [&](int i) { return bar0 + 8 * (NXF + i); };
  auto BAR_DFULL = [&](int h) { return bar0 + 8 * (2 * NXF + h); };           // h = half
  auto BAR_DEMPTY = [&](int h) { return bar0 + 8 * (2 * NXF + 2 + h); };
  auto BAR_SFULL = [&](int h) { return bar0 + 8 * (2 * NXF + 4 + h); };
  auto BAR_D2FULL = [&](int h) { return bar0 + 8 * (2 * NXF + 6 + h); };
  auto BAR_D2EMPTY = [&](int h) { return bar0 + 8 * (2 * NXF + 8 + h); };
  constexpr int BZ = 2 * NXF + 10;
  auto BAR_ZFULL = [&](int i) { return bar0 + 8 * (BZ + i); };
  auto BAR_ZEMPTY = [&](int i) { return bar0 + 8 * (BZ + NZF + i); };
  auto BAR_WFULL = [&](int i) { return bar0 + 8 * (BZ + 2 * NZF + i); };
  auto BAR_WEMPTY = [&](int i) { return bar0 + 8 * (BZ + 2 * NZF + W_STAGES_F + i); };
  auto BAR_ACCFULL = [&](int i) { return bar0 + 8 * (BZ + 2 * NZF + 2 * W_STAGES_F + i); };
  auto BAR_ACCEMPTY = [&](int i) { return bar0 + 8 * (BZ + 2 * NZF + 2 * W_STAGES_F + 2 + i); };
  float* bias_s = reinterpret_cast<float*>(smem + FOFF_BIAS);
  int* prefix = reinterpret_cast<int*>(smem + FOFF_PREFIX);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FOFF_TMEM);

  const int hc = a.dil * (a.K - 1) / 2;
  const int NCH = (a.Cin + KC - 1) / KC;
  const int tile_bytes = n_tile * 64;
  const int tps = a.taps_per_stage;
  const int spc = (a.K + tps - 1) / tps;
  const int nacc = (4 * n_tile <= TM_ACC) ? 2 : 1;

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
    for (int i = 0; i < NXF; ++i) { mbar_init(BAR_XFULL(i), 1); mbar_init(BAR_XEMPTY(i), 1 + NW_ACT); }
    for (int h = 0; h < 2; ++h) {
      mbar_init(BAR_DFULL(h), 1); mbar_init(BAR_DEMPTY(h), NW_ACT); mbar_init(BAR_SFULL(h), NW_ACT);
      mbar_init(BAR_D2FULL(h), 1); mbar_init(BAR_D2EMPTY(h), NW_ACT);
    }
    for (int i = 0; i < NZF; ++i) { mbar_init(BAR_ZFULL(i), NW_ACT); mbar_init(BAR_ZEMPTY(i), 1); }
    for (int i = 0; i < W_STAGES_F; ++i) { mbar_init(BAR_WFULL(i), 1); mbar_init(BAR_WEMPTY(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR_ACCFULL(i), 1); mbar_init(BAR_ACCEMPTY(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
  }
  if (warp >= 2 && warp < 6) {
    // UP[k][cc]: u(block sample cc) = sum_k UP[k][cc] * x(box row 8*bi + k);  cc = 2i: taps up2[11-2m] at k = i+m,
    // cc = 2i+1: taps up2[10-2m] at k = i+1+m (resample.py:19-31 polyphase form, gain folded into up2).
    // fp32 taps = hi + lo: two bf16 MMAs accumulate into the same D1 columns.
    __nv_bfloat16* upb = reinterpret_cast<__nv_bfloat16*>(smem + FOFF_UPB);
    const int t = (warp - 2) * 32 + lane;
    for (int idx = t; idx < 16 * 16; idx += 128) {
      const int k = idx >> 4, n = idx & 15;
      const int i = n >> 1;
      const int m = (n & 1) ? k - i - 1 : k - i;
      float v = 0.f;
      if (m >= 0 && m < 6) v = (n & 1) ? a.up2[10 - 2 * m] : a.up2[11 - 2 * m];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      upb[((k >> 3) * 16 + n) * 8 + (k & 7)] = hi;
      upb[256 + ((k >> 3) * 16 + n) * 8 + (k & 7)] = lo;
    }
    // DN[k][n]: z(row r0 + n) = hb + sum_k DN[k][n] * s(column 2 r0 + k),  DN[2n + 3 + j][n] = dn[j]  (filter.py:87-96)
    __half* dnb = reinterpret_cast<__half*>(smem + FOFF_DNB);
    for (int idx = t; idx < 48 * 16; idx += 128) {
      const int k = idx >> 4, n = idx & 15;
      const int j = k - 2 * n - 3;
      dnb[((k >> 3) * 16 + n) * 8 + (k & 7)] = __float2half_rn((j >= 0 && j < 12) ? a.dn[j] : 0.f);
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
    // ===================== activation warps =====================
    // warp = (r, q): q = TMEM lane quarter = time segment, r = which quarter of a half chunk's columns
    const int q = warp & 3, r = warp >> 2, g = lane >> 3, c8 = lane & 7;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    struct Ctx {
      uint8_t* zrow0; const uint8_t* xrow0;
      int ts, T, zs, xs, n;
      float a2f, nhbf;
      bool edge;
    };
    // one stored z value: row rho of the segment; sequence-edge rows are re-evaluated exactly (scalar), rows outside
    // [0, T) are the conv's zero padding (utils.py:59)
    auto put_z = [&](const Ctx& cx, int rho, float zv) {
      if (cx.edge) {
        const int m = cx.ts + rho;
        if (m < 0 || m >= cx.T) zv = 0.f;
        else if (m < 6 || m >= cx.T - 6) zv = fir_edge_z(cx.xrow0, cx.ts - 7, m, cx.T, cx.a2f, cx.nhbf, a);
      }
      *reinterpret_cast<__nv_bfloat16*>(cx.zrow0 + rho * 16) = __float2bfloat16_rn(zv);
    };
    // D2 buffer of half h (z-blocks of that half, 16 columns each): my quarter of its columns -> +hb -> z tile
    auto extract = [&](const Ctx& cx, int h) {
      mbar_wait(BAR_D2FULL(h), cx.n & 1);
      tc_fence_after();
      uint32_t v[12];
      const int c0 = h == 0 ? 8 * r : 4 * NZ1 * r;                 // first of my columns in the buffer
      const int nc = h == 0 ? 8 : 4 * NZ1;                         // 8 | 4, 8 or 12 columns
      const uint32_t ta = tq + TM_D2 + (uint32_t)(h * 48 + c0);
      if (nc >= 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta));
      if (nc == 4)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(ta));
      if (nc == 12)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]) : "r"(ta + 8));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_D2EMPTY(h));
      if (h == 0 && !BVG_DBGBIT(a, 16)) mbar_wait(BAR_ZEMPTY(cx.zs), ((cx.n / NZF) & 1) ^ 1);   // conv MMAs of this z slot's previous chunk retired
      if (BVG_DBGBIT(a, 64)) return;
      const float hbf = -cx.nhbf;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (j >= nc) break;
        const int c = c0 + j;                                       // column in the half's buffer
        const int zb = (h == 0 ? 0 : NZ0) + (c >> 4);               // z-block of the segment
        const int r0 = (16 * zb < S - 16) ? 16 * zb : S - 16;
        put_z(cx, r0 + (c & 15), __uint_as_float(v[j]) + hbf);
      }
    };
    // D1 buffer of half h: my quarter of its columns -> snake -> fp16 pairs (returned in sw[]), D1 buffer released
    auto load_snake = [&](int h, int ph, u64 a2p, u64 nhbp, uint32_t* sw) {
      mbar_wait(BAR_DFULL(h), ph);
      tc_fence_after();
      const int cw = (h == 0 ? NB0 : NB1) * 4;                      // my columns of the half: 20 or 24
      const uint32_t ta = tq + TM_D1 + (uint32_t)(h * 96 + r * cw);
      uint32_t v[24];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(ta));
      if (cw == 24)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                     : "r"(ta + 16));
      else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]) : "r"(ta + 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_DEMPTY(h));
#pragma unroll
      for (int p = 0; p < 12; ++p) {
        if (2 * p >= cw) break;
        if (BVG_DBGBIT(a, 32)) { sw[p] = v[p]; continue; }
        // s' = u + nhb*cos(a2*u) on two consecutive up-sampled positions -> one fp16x2 TMEM column (even position low)
        const u64 u = pk(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
        float t0f, t1f, s0, s1;
        upk(mul2(a2p, u), t0f, t1f);
        upk(fma2(nhbp, pk(__cosf(t0f), __cosf(t1f)), u), s0, s1);
        sw[p] = f2_to_h2_sat(s0, s1);
      }
    };
    auto store_s = [&](int h, const uint32_t* sw) {
      const int cw2 = (h == 0 ? NB0 : NB1) * 2;                     // my s columns (fp16 pairs): 10 or 12
      const uint32_t ta = tq + TM_S + (uint32_t)((h == 0 ? 0 : NB0 * 8) + r * cw2);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(ta), "r"(sw[0]), "r"(sw[1]), "r"(sw[2]), "r"(sw[3]), "r"(sw[4]), "r"(sw[5]), "r"(sw[6]), "r"(sw[7]) : "memory");
      if (cw2 == 12)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                     ::"r"(ta + 8), "r"(sw[8]), "r"(sw[9]), "r"(sw[10]), "r"(sw[11]) : "memory");
      else
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta + 8), "r"(sw[8]), "r"(sw[9]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_SFULL(h));
    };
    auto finish = [&](const Ctx& cx) {                              // second half of a chunk, then hand the z / x slots over
      extract(cx, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // z stores -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR_ZFULL(cx.zs));
        mbar_arrive(BAR_XEMPTY(cx.xs));
      }
    };
    TileCursor cur{prefix};
    Ctx prev;
    prev.n = -1;
    for (int n = 0; n < total_chunks; ++n) {
      const int it = n / NCH, c = n - it * NCH;
      int b, t0, nt;
      cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
      Ctx cx;
      cx.n = n;
      cx.T = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
      cx.ts = t0 - hc + q * S;                     // time of this segment's z row 0
      const int ch = c * KC + g * 8 + c8;
      cx.a2f = __ldg(a.a2 + ch);
      cx.nhbf = __ldg(a.nhb + ch);
      const u64 a2p = pk(cx.a2f, cx.a2f), nhbp = pk(cx.nhbf, cx.nhbf);
      cx.edge = (cx.ts - 8 < 0) || (cx.ts + S + 8 > cx.T);          // warp-uniform
      cx.xs = n & (NXF - 1);
      cx.zs = n & (NZF - 1);
      cx.zrow0 = smem + FOFF_Z + cx.zs * Z_SLOT + g * (ZRF * 16) + (q * S) * 16 + c8 * 2;
      cx.xrow0 = smem + FOFF_X + cx.xs * X_SLOT + (q * 4 + g) * (XB * 16) + c8 * 2;
      uint32_t sw[12];
      load_snake(0, n & 1, a2p, nhbp, sw);
      // the previous chunk's second batch of z-blocks: its MMAs are done reading s before this chunk overwrites it
      if (prev.n >= 0) finish(prev);
      store_s(0, sw);
      load_snake(1, n & 1, a2p, nhbp, sw);
      store_s(1, sw);
      extract(cx, 0);
      prev = cx;
    }
    if (prev.n >= 0) finish(prev);
  } else if (warp < WARP_EPI) {
    reg_dec<40>();
    if (warp == WARP_XW) {
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
          // A box is XB consecutive 16-byte rows of one (utterance, channel group): contiguous in HBM.  Interior boxes
          // are single 1-D bulk copies (the tensor path walks a box row by row, ~1 row / clk); boxes that leave
          // [0, Tmax) or name a channel group beyond the tensor keep the tensor path for its zero fill.
#pragma unroll 1
          for (int j = 0; j < 16; ++j) {
            const int ts = t0 - hc + (j >> 2) * S - 7, grp = c * 4 + (j & 3);
            if (ts >= 0 && ts + XB <= a.Tmax && grp < a.xgroups && !BVG_DBGBIT(a, 128))
              bulk_load(dst + j * (XB * 16), a.xin + (((size_t)b * a.xgroups + grp) * a.Tmax + ts) * 8, XB * 16, BAR_XFULL(xs));
            else
              tma_load_4d(dst + j * (XB * 16), &tmx, 0, ts, grp, b, BAR_XFULL(xs));
          }
        }
      }
      // ===================== weight producer (bulk copies), second lane of the same warp =====================
      if (lane == 1) {
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
    } else if (warp == WARP_UP) {
      // ===================== up-FIR MMA issuer: D1[half] = X(blocks of the half) * (UP_hi + UP_lo) =====================
      // (warp-convergent: every lane walks the loop, the elected lane issues — see umma_bf16_e in amp_tc.cuh)
      {
        const uint32_t leader = elect_one();
        // A = x tile, MN-major SWIZZLE_NONE (LBO = stride between 8-row K groups, SBO = stride between 8-channel M groups)
        const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) |
                                  ((uint32_t)(128 >> 4) << 24);
        const u64 hiA = make_sdesc(0, 128, XB * 16);
        const u64 bhi = make_sdesc(s_base + FOFF_UPB, 16 * 16, 128), blo = make_sdesc(s_base + FOFF_UPB + 512, 16 * 16, 128);
        for (int n = 0; n < total_chunks; ++n) {
          const int xs = n & (NXF - 1);
          mbar_wait(BAR_XFULL(xs), (n / NXF) & 1);
          const uint32_t a0 = (s_base + FOFF_X + xs * X_SLOT) >> 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(BAR_DEMPTY(h), (n & 1) ^ 1);
            tc_fence_after();
            const int b0 = h == 0 ? 0 : NB0, nb = h == 0 ? NB0 : NB1;
#pragma unroll
            for (int i = 0; i < nb; ++i) {
              if (BVG_DBGBIT(a, 256) && i > 0) break;          // timing experiments only
              const uint32_t td = tmem + TM_D1 + (uint32_t)(h * 96 + i * 16);
              umma_bf16_e(leader, td, hiA | (a0 + (b0 + i) * 8), bhi, idesc_up, 0u);
              umma_bf16_e(leader, td, hiA | (a0 + (b0 + i) * 8), blo, idesc_up, 1u);
            }
            umma_commit_e(leader, BAR_DFULL(h));
          }
          umma_commit_e(leader, BAR_XEMPTY(xs));
        }
      }
    } else if (warp == WARP_DN) {
      // ===================== down-FIR MMA issuer: D2[half] = S(z-block windows, TMEM) * DN =====================
      {
        const uint32_t leader = elect_one();
        const uint32_t idesc_dn = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // fp16 x fp16
        const u64 bdn = make_sdesc(s_base + FOFF_DNB, 16 * 16, 128);
        for (int n = 0; n < total_chunks; ++n) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(BAR_SFULL(h), n & 1);
            mbar_wait(BAR_D2EMPTY(h), (n & 1) ^ 1);
            tc_fence_after();
            const int z0 = h == 0 ? 0 : NZ0, nz = h == 0 ? NZ0 : NZ1;
#pragma unroll
            for (int i = 0; i < nz; ++i) {
              if (BVG_DBGBIT(a, 512) && i > 0) break;          // timing experiments only
              const int r0 = (16 * (z0 + i) < S - 16) ? 16 * (z0 + i) : S - 16;
#pragma unroll
              for (int ks = 0; ks < 3; ++ks)
                umma_ts_f16_e(leader, tmem + TM_D2 + (uint32_t)(h * 48 + i * 16), tmem + TM_S + (uint32_t)(r0 + ks * 8),
                              bdn + (u64)(ks * 32), idesc_dn, ks > 0);
            }
            umma_commit_e(leader, BAR_D2FULL(h));
          }
        }
      }
    } else if (warp == WARP_CONV) {
      // ===================== conv MMA issuer (as k_amp_tc, A = z ring) =====================
      {
        const uint32_t leader = elect_one();
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
                if (BVG_DBGBIT(a, 4)) continue;                 // timing experiments only
                umma_bf16_e(leader, tm, hiA | a0, hiB | b0, idesc, accflag);
                umma_bf16_e(leader, tm, hiA | (a0 + ksA), hiB | (b0 + ksB), idesc, 1u);
                umma_bf16_e(leader, tm + n_tile, hiA | (a0 + 128), hiB | b0, idesc, accflag);
                umma_bf16_e(leader, tm + n_tile, hiA | (a0 + 128 + ksA), hiB | (b0 + ksB), idesc, 1u);
                accflag = 1u;
              }
              umma_commit_e(leader, BAR_WEMPTY(stage));
              if (++stage == W_STAGES_F) { stage = 0; phase ^= 1; }
            }
            umma_commit_e(leader, BAR_ZEMPTY(zs));
          }
          umma_commit_e(leader, BAR_ACCFULL(as));
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    reg_inc<120>();
    epilogue_fir<32>(a, bias_s, prefix, BAR_ACCFULL(0), BAR_ACCEMPTY(0), tmem, nacc, total_tiles, warp & 3, lane,
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
