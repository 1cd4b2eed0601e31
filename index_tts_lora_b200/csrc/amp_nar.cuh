// amp_nar.cuh — the fused AMPBlock1 layer of the NARROW stages (C = 96 / 48 / 24): both kaiser-sinc FIRs of
// Activation1d on the tensor cores, streamed block by block through small TMEM rings.
//
// Same contract as k_amp_tc<L, true>:  xt = conv_{k,d}(Activation1d(x)) [+ resid] [+ sum] [/div]
// (indextts/BigVGAN/models.py:65-74, alias_free_torch/act.py:9-29, resample.py:10-49, filter.py:60-96).
//
// Why: on the CUDA cores the two 12-tap FIRs cost ~24 of the ~28 FMA-pipe lane-ops and ~20 of the ~27 issued
// instructions per element; the narrow stages (58 % of a decode) are bound by exactly that.  As
// banded-Toeplitz GEMMs the FIRs are nearly free on the idle tensor pipe, and what is left for the CUDA cores is
// tcgen05.ld -> SnakeBeta -> tcgen05.st and tcgen05.ld -> +hb -> bf16 z row: ~10 instructions per element, MUFU bound.
// k_amp_fir (amp_fir.cuh, round 1) proved the formulation — layouts, descriptors, precision: per-layer SNR 50-52 dB
// like k_amp_tc — but ran the whole CTA as ONE dependency chain of half-chunk hand-overs and was latency bound.
//
// Here the same arithmetic is a STREAM of 16-column blocks:
//   u-block  UP MMA:  D1[lane, 16 u] = X[lane, 16 x-rows] * (UP_hi + UP_lo)      lane = (time segment, channel)
//            -> one warp per TMEM lane quarter: tcgen05.ld, s = u + nhb*cos(a2*u), fp16 pairs, tcgen05.st -> S ring
//   z-block  DN MMA:  D2[lane, 16 z] = S[lane, 3 u-blocks = 48 s] * DN            A straight from TMEM (TS form)
//            -> one warp per quarter: tcgen05.ld, + hb, bf16, 2-byte stores into the z tile (UMMA A layout of the conv)
// D1 / S / D2 live in 4 / 16 / 8 block slots (320 TMEM columns in all, next to 192 accumulator columns): the UP issuer
// keeps one block per activation group in flight, the DN issuer trails the snake warps by ~2 blocks; chunk and tile
// boundaries are not special.  The 16 activation warps form 4 groups of 4 (one warp per lane quarter); the blocks
// ("tasks") of the stream are dealt to the groups round-robin, so while one group waits for an MMA hand-over the other
// three compute: no hand-over latency is exposed once the rings are full.  Every wait in the kernel (ring slots
// included) is on a task that is EARLIER in the one global task order, hence no deadlock.
//
//   warps 0-15  activation   group = warp >> 2 takes tasks t = group (mod 4); quarter = warp & 3
//   warp 16     lane 0: TMA producer (16 boxes {8 ch, 96 rows} per chunk); lane 1: weights (+ residual rows, identity tiles)
//   warp 17     conv MMA issuer (as k_amp_tc, A = z ring; residual / running sum as D += R x I)
//   warp 18     up-FIR MMA issuer      warp 19  down-FIR MMA issuer       (all three warp-convergent)
//   warps 20-23 epilogue (epilogue_pipe of amp_tc.cuh)
// Sequence edges (replicate clamps of both FIRs) are re-evaluated exactly for the <= 12 affected rows per utterance
// by a scalar path (fir_edge_z); rows outside [0, T) are the conv's zero padding.
#pragma once
#include "amp_fir.cuh"

namespace bvg {
namespace nar {
using namespace tc;
using fir::XB;
using fir::X_SLOT;
using fir::Z_SLOT;
using fir::ZRF;

constexpr int NW_ACT = 16;
constexpr int WARP_XW = NW_ACT, WARP_CONV = NW_ACT + 1, WARP_UP = NW_ACT + 2, WARP_DN = NW_ACT + 3, WARP_EPI = NW_ACT + 4;
constexpr int NTHREADS_N = (WARP_EPI + 4) * 32;      // 768
constexpr int NXN = 3, NZN = 3;          // x / z ring depths (chunks)
constexpr int W_STAGES_N = 2;
// TMEM block slots.  D1 and D2 slots are OWNED by an activation group (two D1 slots per group; one or two D2 slots per
// group, whatever the accumulators leave), so that consecutive uses of a slot are waited for by the same warps in their
// own program order: an mbarrier only carries the parity of its phase, and a warp that could reach the wait for use k
// before use k-1 has completed would pass on the stale phase (seen on hardware with slots shared between groups: the
// warps of non-edge segments run a ring lap ahead of the ones in the scalar edge path).  The S ring is indexed by the
// global u-block number: its reuse distance (16 blocks) exceeds any lag the D1 slots allow.
constexpr int ND1 = 8, NS = 16, ND2_MAX = 8;
constexpr int NZB = 5;                   // z-blocks of 16 rows per segment (S = 72 or 80; the last one may overlap)
constexpr int MAX_NTILE_N = 96;
// TMEM columns: [0, FB) conv accumulators, FB = 128 (C <= 48: two D2 slots per group) or 192 (C = 96: one);
// then [FB, FB+128) D1 slots | [FB+128, FB+256) S ring (fp16 pairs) | [FB+256, 512) D2 slots
constexpr int TM_D1_OFF = 0, TM_S_OFF = ND1 * 16, TM_D2_OFF = TM_S_OFF + NS * 8;
static_assert(192 + TM_D2_OFF + 4 * 16 == 512 && 128 + TM_D2_OFF + 8 * 16 == 512, "TMEM plan");

constexpr int NOFF_BIAS = 0;
constexpr int NOFF_PREFIX = NOFF_BIAS + 2 * 256 * 4;
constexpr int NOFF_BAR = NOFF_PREFIX + (MAX_B + 8) * 4;
// barrier table
constexpr int B_XFULL = 0, B_XEMPTY = B_XFULL + NXN, B_D1FULL = B_XEMPTY + NXN, B_D1EMPTY = B_D1FULL + ND1,
              B_SFULL = B_D1EMPTY + ND1, B_SEMPTY = B_SFULL + NS, B_D2FULL = B_SEMPTY + NS, B_D2EMPTY = B_D2FULL + ND2_MAX,
              B_ZFULL = B_D2EMPTY + ND2_MAX, B_ZEMPTY = B_ZFULL + NZN, B_WFULL = B_ZEMPTY + NZN, B_WEMPTY = B_WFULL + W_STAGES_N,
              B_ACCFULL = B_WEMPTY + W_STAGES_N, B_ACCEMPTY = B_ACCFULL + 2, B_RFULL = B_ACCEMPTY + 2,
              B_REMPTY = B_RFULL + R_RING, N_NUM_BARS = B_REMPTY + R_RING;
constexpr int NOFF_TMEM = NOFF_BAR + N_NUM_BARS * 8;
constexpr int NOFF_UPB = (NOFF_TMEM + 16 + 127) / 128 * 128;   // 2 x [2][16][8] bf16 up taps (hi, lo)
constexpr int NOFF_DNB = NOFF_UPB + 1024;                      // [6][16][8] fp16 down taps
constexpr int NOFF_X = NOFF_DNB + 1536;
constexpr int NOFF_Z = NOFF_X + NXN * X_SLOT;
constexpr int NOFF_W = NOFF_Z + NZN * Z_SLOT;
constexpr int NOFF_R = NOFF_W + W_STAGES_N * W_STAGE_BYTES;    // residual rows (TMA, UMMA A layout) or epilogue staging
constexpr int N_SMEM = NOFF_R + R_STAGE_BYTES;
static_assert(N_SMEM <= 227 * 1024, "k_amp_nar shared-memory plan exceeds 227 KB");
static_assert(NOFF_X % 128 == 0 && X_SLOT % 128 == 0 && Z_SLOT % 128 == 0, "slot alignment");

// The task stream.  Period n = the NUB u-blocks of chunk n interleaved with five z-blocks; Z(zi) reads the s samples of
// u-blocks fb(zi) .. fb(zi)+2, fb = min(2 zi, NUB-3).  A hand-over through the tensor pipe (issue, MMA, commit, barrier)
// takes of the order of 1000 cycles, ~8 tasks at the rate the four groups retire them, so every Z task is placed 8-9
// tasks behind the last u-block it needs: z-blocks 2..4 of a chunk are extracted during the NEXT period (P = previous
// chunk), and one flush period follows the last chunk.  Bit t of ZMASK = position t is a Z task; the i-th Z position of
// a period is z-block (i + 2) % 5, of the previous chunk for i < 3.
//   NUB = 10:  U0 U1 P2 U2 U3 P3 U4 P4 U5 U6 U7 Z0 U8 U9 Z1          NUB = 11:  U0 P2 U1 U2 P3 U3 U4 P4 U5 U6 U7 Z0 U8 U9 Z1 U10
template <int NUB> struct Sched;
template <> struct Sched<10> { static constexpr uint32_t ZMASK = (1u << 2) | (1u << 5) | (1u << 7) | (1u << 11) | (1u << 14); };
template <> struct Sched<11> { static constexpr uint32_t ZMASK = (1u << 1) | (1u << 4) | (1u << 7) | (1u << 11) | (1u << 14); };

#define NAR_TRACE(a, slot, code) TC_TRACE(a, slot, code)

template <int NUB, bool RM>
__global__ void __launch_bounds__(NTHREADS_N, 1)
k_amp_nar(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmr,
          const __grid_constant__ CUtensorMap tmq, const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int S = 8 * (NUB - 1);           // z rows per segment
  constexpr int NT = NUB + NZB;              // tasks per chunk
  constexpr uint32_t ZMASK = Sched<NUB>::ZMASK;
  static_assert(S + 16 <= XB && 4 * S <= ZRF && (S - 16) % 8 == 0 && NZB * 16 >= S, "segment geometry");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = a.n_tile;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + NOFF_BAR;
  auto BAR = [&](int base, int i) { return bar0 + 8 * (base + i); };
  float* bias_s = reinterpret_cast<float*>(smem + NOFF_BIAS);
  int* prefix = reinterpret_cast<int*>(smem + NOFF_PREFIX);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + NOFF_TMEM);

  const int hc = a.dil * (a.K - 1) / 2;
  const int NCH = (a.Cin + KC - 1) / KC;
  const int tile_bytes = n_tile * 64;
  const int tps = a.taps_per_stage;
  const int spc = (a.K + tps - 1) / tps;
  // TMEM plan: accumulators first, then the FIR block slots
  const int fbase = (2 * n_tile <= 128) ? 128 : 192;
  const int nacc = (4 * n_tile <= fbase) ? 2 : 1;
  const int d2d = (fbase == 128) ? 2 : 1;                // D2 slots per activation group
  const uint32_t TM_D1 = (uint32_t)(fbase + TM_D1_OFF), TM_S = (uint32_t)(fbase + TM_S_OFF), TM_D2 = (uint32_t)(fbase + TM_D2_OFF);
  const int nstreams_r = RM ? (a.rmma_r ? 1 : 0) + (a.rmma_q ? 1 : 0) : 0;

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
    for (int i = 0; i < NXN; ++i) { mbar_init(BAR(B_XFULL, i), 1); mbar_init(BAR(B_XEMPTY, i), 1 + 4 * NZB); }
    for (int i = 0; i < ND1; ++i) { mbar_init(BAR(B_D1FULL, i), 1); mbar_init(BAR(B_D1EMPTY, i), 4); }
    for (int i = 0; i < NS; ++i) { mbar_init(BAR(B_SFULL, i), 4); mbar_init(BAR(B_SEMPTY, i), 1); }
    for (int i = 0; i < ND2_MAX; ++i) { mbar_init(BAR(B_D2FULL, i), 1); mbar_init(BAR(B_D2EMPTY, i), 4); }
    for (int i = 0; i < NZN; ++i) { mbar_init(BAR(B_ZFULL, i), 4 * NZB); mbar_init(BAR(B_ZEMPTY, i), 1); }
    for (int i = 0; i < W_STAGES_N; ++i) { mbar_init(BAR(B_WFULL, i), 1); mbar_init(BAR(B_WEMPTY, i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_ACCFULL, i), 1); mbar_init(BAR(B_ACCEMPTY, i), 4); }
    for (int i = 0; i < R_RING; ++i) { mbar_init(BAR(B_RFULL, i), 1); mbar_init(BAR(B_REMPTY, i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
  }
  if (warp >= 2 && warp < 6) {
    // UP[k][cc]: u(block sample cc) = sum_k UP[k][cc] * x(box row 8*bi + k);  cc = 2i: taps up2[11-2m] at k = i+m,
    // cc = 2i+1: taps up2[10-2m] at k = i+1+m (resample.py:19-31 polyphase form, gain folded into up2).
    // fp32 taps = hi + lo: two bf16 MMAs accumulate into the same D1 columns.
    __nv_bfloat16* upb = reinterpret_cast<__nv_bfloat16*>(smem + NOFF_UPB);
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
    __half* dnb = reinterpret_cast<__half*>(smem + NOFF_DNB);
    for (int idx = t; idx < 48 * 16; idx += 128) {
      const int k = idx >> 4, n = idx & 15;
      const int j = k - 2 * n - 3;
      dnb[((k >> 3) * 16 + n) * 8 + (k & 7)] = __float2half_rn((j >= 0 && j < 12) ? a.dn[j] : 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == WARP_CONV) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(s_base + NOFF_TMEM), "r"(512) : "memory");
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
    // register pool of the CTA = 768 x 80 (launch bounds): 16 x 32 x 88 + 4 x 32 x 40 + 4 x 32 x 64 fits; setmaxnreg.inc
    // blocks for ever if the budgets of the three roles add up to more than that
    reg_inc<88>();
    const int q = warp & 3, grp = warp >> 2, g = lane >> 3, c8 = lane & 7;
    const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
    struct Ctx {                               // one chunk as this warp sees it
      uint8_t* zrow0; const uint8_t* xrow0;
      int T, ts, xs, zs, zpar;
      float a2f, nhbf;
      bool edge;
    };
    TileCursor cur{prefix};
    auto make_ctx = [&](int n) {               // n = chunk index of this CTA (called in increasing order)
      Ctx cx;
      const int it = n / NCH, c = n - it * NCH;
      int b, t0, nt;
      cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
      cx.T = a.lengths ? __ldg(a.lengths + b) * a.rate : a.Tmax;
      cx.ts = t0 - hc + q * S;                    // time of this segment's z row 0
      const int ch = c * KC + lane;
      cx.a2f = __ldg(a.a2 + ch);
      cx.nhbf = __ldg(a.nhb + ch);
      cx.edge = (cx.ts - 8 < 0) || (cx.ts + S + 8 > cx.T);    // warp-uniform
      cx.xs = n % NXN;
      cx.zs = n % NZN;
      cx.zpar = ((n / NZN) & 1) ^ 1;
      cx.zrow0 = smem + NOFF_Z + cx.zs * Z_SLOT + g * (ZRF * 16) + (q * S) * 16 + c8 * 2;
      cx.xrow0 = smem + NOFF_X + cx.xs * X_SLOT + (q * 4 + g) * (XB * 16) + c8 * 2;
      return cx;
    };
    Ctx cxc{}, cxp{};                          // current / previous chunk
    int per = 0, tt = grp;                     // period (= chunk whose u-blocks it carries), position within it
    int ucnt = 0, zcnt = 0;                    // U / Z tasks this group has done (slot and phase of its own slots)
    if (total_chunks > 0) cxc = make_ctx(0);
    while (per <= total_chunks && total_chunks > 0) {
      const bool isz = (ZMASK >> tt) & 1u;
      const int nz_before = __popc(ZMASK & ((1u << tt) - 1u));
      if (!isz) {
        if (per < total_chunks) {
          // ---- U task: D1 slot -> snake -> S slot
          const int gub = per * NUB + (tt - nz_before);
          const int d1 = 2 * grp + (ucnt & 1), ss = gub & (NS - 1);
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 0, 100 + tt);
          mbar_wait(BAR(B_D1FULL, d1), (ucnt >> 1) & 1);
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 1, 200 + tt);
          ++ucnt;
          tc_fence_after();
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tq + TM_D1 + (uint32_t)(d1 * 16)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_D1EMPTY, d1));
          const u64 a2p = pk(cxc.a2f, cxc.a2f), nhbp = pk(cxc.nhbf, cxc.nhbf);
          uint32_t sw[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            // s' = u + nhb*cos(a2*u) on two consecutive up-sampled positions -> one fp16x2 TMEM column (even position low)
            const u64 u = pk(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
            float t0f, t1f, s0, s1;
            upk(mul2(a2p, u), t0f, t1f);
            upk(fma2(nhbp, pk(__cosf(t0f), __cosf(t1f)), u), s0, s1);
            sw[p] = fir::f2_to_h2_sat(s0, s1);
          }
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 2, 300 + tt);
          mbar_wait(BAR(B_SEMPTY, ss), ((gub / NS) & 1) ^ 1);    // the down MMAs that read this slot's previous block retired
          tc_fence_after();
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                       ::"r"(tq + TM_S + (uint32_t)(ss * 8)), "r"(sw[0]), "r"(sw[1]), "r"(sw[2]), "r"(sw[3]), "r"(sw[4]),
                         "r"(sw[5]), "r"(sw[6]), "r"(sw[7]) : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_SFULL, ss));
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 3, 400 + tt);
        }
      } else {
        // ---- Z task: D2 slot -> + hb -> bf16 rows of the z tile
        const int zi = (nz_before + 2) % NZB;
        const bool prevc = nz_before < 3;
        const int nchunk = prevc ? per - 1 : per;
        if (nchunk >= 0 && nchunk < total_chunks) {
          const Ctx& cx = prevc ? cxp : cxc;
          const int d2 = d2d * grp + (d2d == 2 ? (zcnt & 1) : 0);
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 0, 500 + tt);
          mbar_wait(BAR(B_D2FULL, d2), (d2d == 2 ? (zcnt >> 1) : zcnt) & 1);
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 1, 600 + tt);
          ++zcnt;
          tc_fence_after();
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tq + TM_D2 + (uint32_t)(d2 * 16)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_D2EMPTY, d2));
          mbar_wait(BAR(B_ZEMPTY, cx.zs), cx.zpar);      // conv MMAs of this z slot's previous chunk retired
          const int r0 = (16 * zi < S - 16) ? 16 * zi : S - 16;
          const float hbf = -cx.nhbf;
          uint8_t* zr = cx.zrow0 + r0 * 16;
          if (!cx.edge) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              *reinterpret_cast<__nv_bfloat16*>(zr + j * 16) = __float2bfloat16_rn(__uint_as_float(v[j]) + hbf);
          } else {
            // sequence-edge rows are re-evaluated exactly (scalar); rows outside [0, T) are the conv's zero padding
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = cx.ts + r0 + j;
              float zv = __uint_as_float(v[j]) + hbf;
              if (m < 0 || m >= cx.T) zv = 0.f;
              else if (m < 6 || m >= cx.T - 6) zv = fir::fir_edge_z(cx.xrow0, cx.ts - 7, m, cx.T, cx.a2f, cx.nhbf, a);
              *reinterpret_cast<__nv_bfloat16*>(zr + j * 16) = __float2bfloat16_rn(zv);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // z stores -> async proxy (UMMA)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(BAR(B_ZFULL, cx.zs));
            mbar_arrive(BAR(B_XEMPTY, cx.xs));
          }
          if (q == 0) NAR_TRACE(a, (per * NT + tt) * 8 + 3, 700 + tt);
        }
      }
      // next task of this group
      tt += 4;
      if (tt >= NT) {
        tt -= NT;
        ++per;
        cxp = cxc;
        if (per < total_chunks) cxc = make_ctx(per);
      }
    }
  } else if (warp < WARP_EPI) {
    reg_dec<40>();
    if (warp == WARP_XW) {
      // ===================== x producer (TMA): 16 boxes (segment, channel group) per chunk =====================
      if (lane == 0) {
        TileCursor cur{prefix};
        int b = 0, t0 = 0, nt, xs = 0, xph = 0, c = 0, it = 0;
        for (int n = 0; n < total_chunks; ++n) {
          if (c == 0) cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
          mbar_wait_relaxed(BAR(B_XEMPTY, xs), xph ^ 1, 200);
          mbar_expect_tx(BAR(B_XFULL, xs), X_SLOT);
          const uint32_t dst = s_base + NOFF_X + xs * X_SLOT;
          // A box is XB consecutive 16-byte rows of one (utterance, channel group): contiguous in HBM.  Interior boxes
          // are single 1-D bulk copies; boxes that leave [0, Tmax) or name a channel group beyond the tensor keep the
          // tensor path for its zero fill.
#pragma unroll 1
          for (int j = 0; j < 16; ++j) {
            const int tsj = t0 - hc + (j >> 2) * S - 7, grp = c * 4 + (j & 3);
            if (tsj >= 0 && tsj + XB <= a.Tmax && grp < a.xgroups)
              bulk_load(dst + j * (XB * 16), a.xin + (((size_t)b * a.xgroups + grp) * a.Tmax + tsj) * 8, XB * 16, BAR(B_XFULL, xs));
            else
              tma_load_4d(dst + j * (XB * 16), &tmx, 0, tsj, grp, b, BAR(B_XFULL, xs));
          }
          if (++xs == NXN) { xs = 0; xph ^= 1; }
          if (++c == NCH) { c = 0; ++it; }
        }
      }
      // ===================== weight producer (bulk copies), second lane of the same warp =====================
      // ... and the residual / running-sum rows of the tile with their identity tiles (D += R x I, as k_amp_tc)
      if (lane == 1) {
        int stage = 0, phase = 0, rs = 0, rph = 0;
        TileCursor cur{prefix};
        for (int it = 0; it < my_tiles; ++it) {
          int b, t0, nt;
          cur.locate((int)blockIdx.x + it * (int)gridDim.x, 1, b, t0, nt);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wt);
          for (int c = 0; c < NCH; ++c) {
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              const uint32_t bytes = (uint32_t)(taps * tile_bytes);
              mbar_wait_relaxed(BAR(B_WEMPTY, stage), phase ^ 1, 200);
              mbar_expect_tx(BAR(B_WFULL, stage), bytes);
              bulk_load(s_base + NOFF_W + stage * W_STAGE_BYTES, src, bytes, BAR(B_WFULL, stage));
              src += bytes;
              if (++stage == W_STAGES_N) { stage = 0; phase ^= 1; }
            }
            if (RM && c < a.nchr)
              for (int st = 0; st < nstreams_r; ++st) {
                const CUtensorMap* rm = (st == 0 && a.rmma_r) ? &tmr : &tmq;
                const uint8_t* isrc = reinterpret_cast<const uint8_t*>(a.idw) + (size_t)c * tile_bytes;
                mbar_wait_relaxed(BAR(B_WEMPTY, stage), phase ^ 1, 200);
                mbar_expect_tx(BAR(B_WFULL, stage), (uint32_t)tile_bytes);
                bulk_load(s_base + NOFF_W + stage * W_STAGE_BYTES, isrc, (uint32_t)tile_bytes, BAR(B_WFULL, stage));
                if (++stage == W_STAGES_N) { stage = 0; phase ^= 1; }
                mbar_wait_relaxed(BAR(B_REMPTY, rs), rph ^ 1, 200);
                mbar_expect_tx(BAR(B_RFULL, rs), R_SLOT_BYTES);
                const uint32_t dst = s_base + NOFF_R + rs * R_SLOT_BYTES;
#pragma unroll
                for (int kg = 0; kg < 4; ++kg)
#pragma unroll
                  for (int h = 0; h < 2; ++h)
                    tma_load_4d(dst + kg * (M_TILE * 16) + h * (128 * 16), rm, 0, t0 + h * 128, c * 4 + kg, b, BAR(B_RFULL, rs));
                if (++rs == R_RING) { rs = 0; rph ^= 1; }
              }
          }
        }
      }
    } else if (warp == WARP_UP) {
      // ===================== up-FIR MMA issuer: D1[slot] = X(block) * (UP_hi + UP_lo) =====================
      {
        const uint32_t leader = elect_one();
        // A = x tile, MN-major SWIZZLE_NONE (LBO = stride between 8-row K groups, SBO = stride between 8-channel M groups)
        const uint32_t idesc_up = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) |
                                  ((uint32_t)(128 >> 4) << 24);
        const u64 hiA = make_sdesc(0, 128, XB * 16);
        const u64 bhi = make_sdesc(s_base + NOFF_UPB, 16 * 16, 128), blo = make_sdesc(s_base + NOFF_UPB + 512, 16 * 16, 128);
        int xs = 0, xph = 0;
        uint32_t pbase = 0;                       // global task index of the period's position 0 (mod 4)
        uint32_t uc = 0;                          // 2 bits per group: u-blocks issued into the group's two slots (mod 4)
        for (int n = 0; n < total_chunks; ++n) {
          mbar_wait(BAR(B_XFULL, xs), xph);
          tc_fence_after();
          const uint32_t a0 = (s_base + NOFF_X + xs * X_SLOT) >> 4;
          int ub = 0;
#pragma unroll 1
          for (int t = 0; t < NT; ++t) {
            if ((ZMASK >> t) & 1u) continue;
            const int G = (int)((pbase + t) & 3u);           // the group that takes position t owns the slot
            const uint32_t cnt = (uc >> (2 * G)) & 3u;
            const int d1 = 2 * G + (int)(cnt & 1u);
            uc = (uc & ~(3u << (2 * G))) | (((cnt + 1u) & 3u) << (2 * G));
            mbar_wait(BAR(B_D1EMPTY, d1), ((cnt >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t td = tmem + TM_D1 + (uint32_t)(d1 * 16);
            umma_bf16_e(leader, td, hiA | (a0 + ub * 8), bhi, idesc_up, 0u);
            umma_bf16_e(leader, td, hiA | (a0 + ub * 8), blo, idesc_up, 1u);
            umma_commit_e(leader, BAR(B_D1FULL, d1));
            NAR_TRACE(a, (n * NT + t) * 8 + 4, 800 + t);
            ++ub;
          }
          umma_commit_e(leader, BAR(B_XEMPTY, xs));
          if (++xs == NXN) { xs = 0; xph ^= 1; }
          pbase = (pbase + NT) & 3u;
        }
      }
    } else if (warp == WARP_DN) {
      // ===================== down-FIR MMA issuer: D2[slot] = S(3 u-block slots, TMEM) * DN =====================
      {
        const uint32_t leader = elect_one();
        const uint32_t idesc_dn = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // fp16 x fp16
        const u64 bdn = make_sdesc(s_base + NOFF_DNB, 16 * 16, 128);
        uint32_t pbase = 0;                       // global task index of the period's position 0 (mod 4)
        uint32_t zc = 0;                          // 2 bits per group: z-blocks issued into the group's slot(s) (mod 4)
        for (int per = 0; per <= total_chunks; ++per) {
          int iz = 0;
#pragma unroll 1
          for (int t = 0; t < NT; ++t) {
            if (!((ZMASK >> t) & 1u)) continue;
            const int zi = (iz + 2) % NZB;
            const int nchunk = (iz < 3) ? per - 1 : per;
            ++iz;
            if (nchunk < 0 || nchunk >= total_chunks) continue;
            const int gub0 = nchunk * NUB;
            const int fb = (2 * zi < NUB - 3) ? 2 * zi : NUB - 3;
            const int G = (int)((pbase + t) & 3u);
            const uint32_t cnt = (zc >> (2 * G)) & 3u;
            const int d2 = d2d * G + (d2d == 2 ? (int)(cnt & 1u) : 0);
            const uint32_t par = (d2d == 2 ? (cnt >> 1) : cnt) & 1u;
            zc = (zc & ~(3u << (2 * G))) | (((cnt + 1u) & 3u) << (2 * G));
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const int gub = gub0 + fb + ks;
              mbar_wait(BAR(B_SFULL, gub & (NS - 1)), (gub / NS) & 1);
            }
            NAR_TRACE(a, (per * NT + t) * 8 + 5, 900 + t);
            mbar_wait(BAR(B_D2EMPTY, d2), par ^ 1u);
            tc_fence_after();
            NAR_TRACE(a, (per * NT + t) * 8 + 6, 1000 + t);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const int gub = gub0 + fb + ks;
              fir::umma_ts_f16_e(leader, tmem + TM_D2 + (uint32_t)(d2 * 16), tmem + TM_S + (uint32_t)((gub & (NS - 1)) * 8),
                                 bdn + (u64)(ks * 32), idesc_dn, ks > 0);
            }
            umma_commit_e(leader, BAR(B_D2FULL, d2));
            // s slots whose last reader this z-block is: u-blocks [fb, fb of the next z-block), the tail for the last one
            const int fbn = (zi == NZB - 1) ? NUB : ((2 * (zi + 1) < NUB - 3) ? 2 * (zi + 1) : NUB - 3);
            for (int ub = fb; ub < fbn; ++ub) umma_commit_e(leader, BAR(B_SEMPTY, (gub0 + ub) & (NS - 1)));
          }
          pbase = (pbase + NT) & 3u;
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
        const u64 hiR = make_sdesc(0, M_TILE * 16, 128);    // residual slot: 256 rows per 8-channel group
        int stage = 0, phase = 0, zs = 0, zph = 0, rr = 0, rrph = 0;
        for (int it = 0; it < my_tiles; ++it) {
          const int as = (nacc == 2) ? (it & 1) : 0;
          const int ause = (nacc == 2) ? (it >> 1) : it;
          mbar_wait(BAR(B_ACCEMPTY, as), (ause & 1) ^ 1);
          tc_fence_after();
          const uint32_t tm = tmem + (uint32_t)(as * 2 * n_tile);
          uint32_t accflag = 0;
          for (int c = 0; c < NCH; ++c) {
            mbar_wait_relaxed(BAR(B_ZFULL, zs), zph, 200);
            tc_fence_after();
            const uint32_t aU = (s_base + NOFF_Z + zs * Z_SLOT) >> 4;
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              mbar_wait(BAR(B_WFULL, stage), phase);
              tc_fence_after();
              const uint32_t wU = (s_base + NOFF_W + stage * W_STAGE_BYTES) >> 4;
              for (int tj = 0; tj < taps; ++tj) {
                const uint32_t a0 = aU + (uint32_t)((s * tps + tj) * a.dil);
                const uint32_t b0 = wU + (uint32_t)tj * tileU;
                umma_bf16_e(leader, tm, hiA | a0, hiB | b0, idesc, accflag);
                umma_bf16_e(leader, tm, hiA | (a0 + ksA), hiB | (b0 + ksB), idesc, 1u);
                umma_bf16_e(leader, tm + n_tile, hiA | (a0 + 128), hiB | b0, idesc, accflag);
                umma_bf16_e(leader, tm + n_tile, hiA | (a0 + 128 + ksA), hiB | (b0 + ksB), idesc, 1u);
                accflag = 1u;
              }
              umma_commit_e(leader, BAR(B_WEMPTY, stage));
              if (++stage == W_STAGES_N) { stage = 0; phase ^= 1; }
            }
            umma_commit_e(leader, BAR(B_ZEMPTY, zs));
            if (++zs == NZN) { zs = 0; zph ^= 1; }
            // + residual (+ running sum) chunk c: D += R x I, two K steps of 16 channels
            if (RM && c < a.nchr)
              for (int st = 0; st < nstreams_r; ++st) {
                mbar_wait(BAR(B_RFULL, rr), rrph);
                mbar_wait(BAR(B_WFULL, stage), phase);
                tc_fence_after();
                const uint32_t r0 = (s_base + NOFF_R + rr * R_SLOT_BYTES) >> 4;
                const uint32_t b0 = (s_base + NOFF_W + stage * W_STAGE_BYTES) >> 4;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  umma_bf16_e(leader, tm, hiR | (r0 + ks * (2 * M_TILE)), hiB | (b0 + ks * ksB), idesc, 1u);
                  umma_bf16_e(leader, tm + n_tile, hiR | (r0 + ks * (2 * M_TILE) + 128), hiB | (b0 + ks * ksB), idesc, 1u);
                }
                umma_commit_e(leader, BAR(B_WEMPTY, stage));
                if (++stage == W_STAGES_N) { stage = 0; phase ^= 1; }
                umma_commit_e(leader, BAR(B_REMPTY, rr));
                if (++rr == R_RING) { rr = 0; rrph ^= 1; }
              }
          }
          umma_commit_e(leader, BAR(B_ACCFULL, as));
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    reg_dec<64>();
    epilogue_pipe<false, RM>(a, bias_s, smem + NOFF_R, prefix, BAR(B_ACCFULL, 0), BAR(B_ACCEMPTY, 0), tmem, nacc, total_tiles,
                             warp & 3, lane, threadIdx.x - WARP_EPI * 32);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_CONV) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace nar
}  // namespace bvg
