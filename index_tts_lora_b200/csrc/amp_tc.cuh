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
// over C_in in chunks of 32 channels.  Persistent grid (one CTA per SM), 24 warps:
//   warps 0-15  activation: x -> 2x kaiser-sinc FIR up -> SnakeBeta -> FIR down -> bf16 z tile in the
//               UMMA A layout (fp32 math on packed f32x2 registers, two channels per thread, one
//               run of L consecutive rows per thread, neighbours' samples by warp shuffle)
//   warp 16     TMA producer: x rows [t0-32, t0+288) of the chunk -> x ring   (8 boxes of {8,160})
//   warp 17     weight producer: per-(chunk,tap) bf16 tiles [4][n_tile][8] via cp.async.bulk -> weight ring;
//               also the residual / running-sum rows (TMA) and identity tiles of D += R x I (RM)
//   warp 18     MMA issuer: for every tap j, M block, K step: tcgen05.mma.kind::f16
//               A = z tile rows (mb*128 + j*dil ..), B = weight tile, D = TMEM[stage][mb*n_tile ..]
//   warps 20-23 epilogue: TMEM -> +bias (+cond) [+resid] [+sum] [/3] -> bf16 -> HBM; one of three variants
//               per instantiation (EPI): conv / conv with several column tiles / ConvTranspose1d scatter
// Synchronisation is mbarrier-only inside the main loop; DESIGN.md §4.1 has the measurements behind each choice.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace bvg {
namespace tc {

constexpr int M_TILE = 256;
constexpr int KC = 32;        // input channels per chunk (4 groups of 8)
constexpr int XR = 320;       // x rows TMA-staged per chunk: t0-32 .. t0+287
constexpr int XRA = 368;      // x rows allocated (slack; rows >= XR only ever feed discarded z rows)
constexpr int X_LEAD = 32;
constexpr int BOXR = 160;     // TMA box rows (two boxes per channel group)
constexpr int ZR = 336;       // z rows allocated (16 runs x 21)
constexpr int NW_ACT = 16;    // warps 0-15 activation | 16 x-TMA | 17 weights | 18 MMA | 19 spare | 20-23 epilogue
constexpr int NX_MAX = 3;     // x ring depth
constexpr int NZ_MAX = 4;     // z ring depth: lets the activation run ahead while an epilogue drains TMEM
constexpr int W_STAGES_MAX = 8;
constexpr int W_STAGE_BYTES = 16384;
constexpr int NTHREADS = 768;
constexpr int MAX_B = 512;    // utterances per launch (tile-prefix table lives in smem; 2 KB — the wide layers use every byte)

constexpr int X_BUF_BYTES = 4 * XRA * 16;  // 22016
constexpr int X_TX_BYTES = 4 * XR * 16;    // 20480 bytes actually delivered by TMA
constexpr int Z_BUF_BYTES = 4 * ZR * 16;   // 21504
// shared-memory map: [bias 2x256 f32][tile prefix][barriers][tmem slot] | x ring | z ring | weight ring
// (ring depths are runtime: nx x-buffers and wst weight stages, chosen per layer by the host)
constexpr int OFF_BIAS = 0;
constexpr int OFF_PREFIX = OFF_BIAS + 2 * 256 * 4;
constexpr int OFF_BAR = OFF_PREFIX + (MAX_B + 8) * 4;
constexpr int NUM_BARS = 2 * NX_MAX + 2 * NZ_MAX + 2 * W_STAGES_MAX + 4 + 4;   // + residual ring full / empty x 2
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int OFF_X = (OFF_TMEM + 16 + 127) / 128 * 128;
__host__ __device__ constexpr int off_z(int nx) { return OFF_X + nx * X_BUF_BYTES; }
__host__ __device__ constexpr int off_w(int nx, int nz) { return off_z(nx) + nz * Z_BUF_BYTES; }
constexpr int R_SLOTS = 16;                       // 16-byte rows per epilogue thread staged for the residual / running sum
constexpr int R_STAGE_BYTES = R_SLOTS * 128 * 16; // 32 KB: [slot][epilogue thread][16 B]
// ... or, when the residual is added by the tensor core, two slots of [4 channel groups][256 rows][16 B] (UMMA A layout)
constexpr int R_RING = 2;
constexpr int R_SLOT_BYTES = 4 * M_TILE * 16;     // 16 KB
static_assert(R_RING * R_SLOT_BYTES == R_STAGE_BYTES, "the two uses share one region");
__host__ __device__ constexpr int off_r(int nx, int nz, int wst) { return off_w(nx, nz) + wst * W_STAGE_BYTES; }
__host__ __device__ constexpr int smem_bytes(int nx, int nz, int wst) { return off_r(nx, nz, wst) + R_STAGE_BYTES; }

struct TcArgs {
  const __nv_bfloat16* wt;     // [ntile][chunk][tap][4][n_tile][8] bf16
  const float* bias;           // [Cout]
  const float* bias_b;         // [B][bias_b_stride] speaker-conditioning add, or null
  int bias_b_stride;
  const __nv_bfloat16* resid;  // blocked, or null
  const __nv_bfloat16* acc_in; // blocked, or null
  __nv_bfloat16* out;          // blocked [B][Cout/8][Tstride][8]
  float div;
  int Cin, Cout, K, dil, n_tile, n_tiles, taps_per_stage;
  int nx, nz, wst;             // ring depths: x buffers (2..3), z buffers (2..4), weight stages (<= 8)
  const __nv_bfloat16* xin;    // k_amp_fir: the blocked input buffer itself (interior boxes are plain 1-D bulk copies)
  int xgroups;                 // channel groups of xin
  // residual / running-sum add on the tensor core (D += R x I): identity weight tiles [ntile][chunk][1][4][n_tile][8],
  // chunks of 32 channels per column tile, and which streams are on (their rows arrive through tmr / tmq)
  const __nv_bfloat16* idw;
  int nchr, rmma_r, rmma_q;
  int dbg;                     // timing experiments only (BVG_DBG env): 1 = 1 of 4 MMAs per tap, 2 = 16-byte weight copies
  int st_lo, st_hi;            // conv mode: only rows in [st_lo, st_hi) are stored (time-split shards keep
                               // their hands off the halo rows that the neighbouring GPUs write)
  int B;
  int Tstride;                 // rows per (b, channel group) in every activation buffer of this stage
  const int* lengths;
  int rate, Tmax;
  int lead;                    // ACT=false: rows of x before t0 that tap 0 reads (conv: d(k-1)/2)
  // ConvTranspose1d mode (up > 0), models.py:157-163: K = taps per phase (k/u), columns are
  // phase-major n = r*cphase + co, row q = input time, output time = q*up + r - pad.
  int up, pad, cphase;
  const float* a2;             // [Cin padded to 32]  2*exp(alpha)
  const float* nhb;            // [Cin padded to 32]  -0.5/(exp(beta)+1e-9)
  int cl;                      // thread-block cluster size (= n_tiles, 2 or 3) when the column tiles of a time tile share one
                               // activation through distributed shared memory; 0 / 1 = every CTA on its own
  long long* trace;            // BVG_EXPERIMENTS builds only: clock-stamped pipeline events of CTA 0 (tools/nar_trace.py)
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
// Polite wait for the roles that are normally far ahead of the activation warps (producers, MMA issuer of the narrow
// layers, epilogue): they wait most of a tile, and whatever they execute meanwhile comes out of the issue slots of the
// activation warps on the same scheduler.  try_wait with a suspend-time hint was measured to return after ~20 ns
// whatever the hint (ncu, round 1: ~140 poll iterations of 4 instructions per role and tile, 20 % of all issued
// instructions of a narrow layer), so the loop sleeps with a plain nanosleep between non-blocking tests: one test per
// `ns` nanoseconds.  Every one of these hand-overs has at least a ring slot (one chunk or tile) of slack, far more than
// the worst-case wake-up delay of one sleep.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns) {
  if (mbar_test_wait(bar, parity)) return;
#pragma unroll 1
  for (;;) {
    asm volatile("nanosleep.u32 %0;" ::"r"(ns) : "memory");
    if (mbar_test_wait(bar, parity)) break;
  }
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
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
// Warp-convergent forms: every lane executes the statement with identical (uniform) operands and the instruction
// itself is predicated on the elected lane.  Issued from a divergent `if (lane == 0)` region, ptxas wraps every
// UTCHMMA in an ELECT / BRA.U.ANY waterfall plus R2UR moves (~15 instructions and ~50 cycles per MMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_bf16_e(uint32_t leader, uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {   // my shared address -> the same one in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// remote arrive that also announces `bytes` of bulk-copy traffic for the current phase of that barrier
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cbar), "r"(bytes) : "memory");
}
// bulk copy (async proxy) from my shared memory into a peer CTA's; completes `bytes` on the peer's barrier `cbar`
__device__ __forceinline__ void dsmem_copy(uint32_t cdst, uint32_t src, uint32_t bytes, uint32_t cbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(cdst), "r"(src), "r"(bytes), "r"(cbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cbar) {             // release at cluster scope, remote barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\tbra WAIT_LOOP_C;\n\tDONE_C:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit that arrives on the barrier at this offset in EVERY CTA of the mask
__device__ __forceinline__ void umma_commit_mc_e(uint32_t leader, uint32_t bar, uint32_t mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\tsetp.ne.b32 e, %1, 0;\n\tcvt.u16.u32 m, %2;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n"
      ::"r"(bar), "r"(leader), "r"(mask) : "memory");
}

// pipeline trace (experiments builds): record (code, clock) for the first tiles of CTA 0; one writer per slot
#ifdef BVG_EXPERIMENTS
#define TC_TRACE(a, slot, code)                                                                        \
  do {                                                                                                 \
    if ((a).trace && blockIdx.x == 0 && (slot) < 4096 && (threadIdx.x & 31) == 0) {                    \
      (a).trace[2 * (slot)] = (long long)(code);                                                       \
      (a).trace[2 * (slot) + 1] = clock64();                                                           \
    }                                                                                                  \
  } while (0)
#else
#define TC_TRACE(a, slot, code) do { } while (0)
#endif

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
  return pk(__uint_as_float(__byte_perm(w, 0, 0x1044)), __uint_as_float(w & 0xffff0000u));
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

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
  return ((u64)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src) << 32) |
         (u64)__shfl_sync(0xffffffffu, (uint32_t)v, src);
}

// One thread: channel pair, L consecutive rows starting at local z row `rowS` (global time m0).
// It computes its own 2L up-sampled snake samples s'(2*m0 .. 2*m0+2L-1), receives the 5 before /
// 6 after from the neighbouring runs of the same warp by shuffle (lane -4 / +4), and writes the
// z rows that fall inside the warp's valid range [vlo, vhi) — the first / last 3 rows of a warp's
// span lack a neighbour and belong to the adjacent warp's range (spans overlap by 6 rows).
// Must be entered by all 32 lanes of the warp (EDGE is warp-uniform).
template <int L, bool EDGE>
__device__ __forceinline__ void act_run(const uint32_t* __restrict__ xk, uint32_t* __restrict__ zk, int rowS,
                                        uint32_t smask, int m0, int xlo, int T, const ActCtx& k, int lane) {
  u64 xw[L + 6];
  u64 sv[2 * L];
  u64 s_first = 0ull, s_last = 0ull;
  if (EDGE) {
    s_first = snake_s_at(xk, 0, xlo, T, k);
    s_last = snake_s_at(xk, 2 * T - 1, xlo, T, k);
  }
#pragma unroll
  for (int i = 0; i < L + 6; ++i) {
    int t = m0 - 3 + i;
    if (EDGE) t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    int r = t - xlo;
    if (EDGE) r = r < 0 ? 0 : (r > XRA - 1 ? XRA - 1 : r);   // interior runs: 1 <= r <= 8L*4+35 < XRA
    xw[i] = bf2_to_f2(xk[r * 4]);
  }
  // QB input positions per batch = 2*QB independent FMA chains in flight (even / odd phase of each): the chains of
  // one up-sampled sample are 6 dependent FFMA2s, and a warp that walks them two at a time stalls on the FMA latency
  constexpr int QB = 3;
#pragma unroll
  for (int q0 = 0; q0 < L; q0 += QB) {
    u64 ye[QB], yo[QB];
#pragma unroll
    for (int qq = 0; qq < QB; ++qq)
      if (q0 + qq < L) {
        ye[qq] = mul2(k.upE[0], xw[q0 + qq]);          // n even: taps f[11-2i] on x[q-3+i]
        yo[qq] = mul2(k.upO[0], xw[q0 + qq + 1]);      // n odd:  taps f[10-2i] on x[q-2+i]
      }
#pragma unroll
    for (int i = 1; i < 6; ++i)
#pragma unroll
      for (int qq = 0; qq < QB; ++qq)
        if (q0 + qq < L) {
          ye[qq] = fma2(k.upE[i], xw[q0 + qq + i], ye[qq]);
          yo[qq] = fma2(k.upO[i], xw[q0 + qq + 1 + i], yo[qq]);
        }
#pragma unroll
    for (int qq = 0; qq < QB; ++qq)
      if (q0 + qq < L) {
        u64 se = snake_s(ye[qq], k), so = snake_s(yo[qq], k);
        if (EDGE) {
          const int n = 2 * (m0 + q0 + qq);
          if (n < 0) se = s_first;
          else if (n > 2 * T - 1) se = s_last;
          if (n + 1 < 0) so = s_first;
          else if (n + 1 > 2 * T - 1) so = s_last;
        }
        sv[2 * (q0 + qq)] = se;
        sv[2 * (q0 + qq) + 1] = so;
      }
  }
  u64 sb[5], sa[6];
#pragma unroll
  for (int j = 0; j < 5; ++j) sb[j] = shfl64(sv[2 * L - 5 + j], (lane + 28) & 31);
#pragma unroll
  for (int j = 0; j < 6; ++j) sa[j] = shfl64(sv[j], (lane + 4) & 31);
  u64 zz[L];                             // tap-outer: L independent accumulation chains
#pragma unroll
  for (int r = 0; r < L; ++r) zz[r] = k.hb;
#pragma unroll
  for (int j = 0; j < 12; ++j)
#pragma unroll
    for (int r = 0; r < L; ++r) {
      const int q = 2 * r + j - 5;       // index into this run's own samples
      const u64 sj = q < 0 ? sb[5 + q] : (q < 2 * L ? sv[q] : sa[q - 2 * L]);
      zz[r] = fma2(k.dn[j], sj, zz[r]);
    }
#pragma unroll
  for (int r = 0; r < L; ++r) {
    const u64 z = zz[r];
    float z0, z1;
    upk(z, z0, z1);
    if (EDGE) {
      const int m = m0 + r;
      if (m < 0 || m >= T) { z0 = 0.f; z1 = 0.f; }      // conv zero padding (utils.py:59)
    }
    if ((smask >> r) & 1u) {             // row inside the warp's valid range
      __nv_bfloat162 o = __floats2bfloat162_rn(z0, z1);
      zk[(rowS + r) * 4] = *reinterpret_cast<uint32_t*>(&o);
    }
  }
}

template <class Args>
__device__ __forceinline__ void load_taps(ActCtx& k, const Args& a) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    k.upE[i] = pk(a.up2[11 - 2 * i], a.up2[11 - 2 * i]);
    k.upO[i] = pk(a.up2[10 - 2 * i], a.up2[10 - 2 * i]);
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) k.dn[i] = pk(a.dn[i], a.dn[i]);
}

// Sequence-edge tiles (2 per utterance) take this out-of-line copy so that none of its clamp /
// select arithmetic is hoisted into the interior path.
template <int L, class Args>
__device__ __noinline__ void act_run_edge(const uint32_t* xk, uint32_t* zk, int rowS, uint32_t smask, int m0,
                                          int xlo, int T, u64 a2, u64 nhb, const Args& a, int lane) {
  ActCtx k;
  load_taps(k, a);
  k.a2 = a2;
  k.nhb = nhb;
  float h0, h1;
  upk(nhb, h0, h1);
  k.hb = pk(-h0, -h1);
  act_run<L, true>(xk, zk, rowS, smask, m0, xlo, T, k, lane);
}

// ------------------------------------------------------------------------------ epilogue role
// Four warps (q = TMEM lane quarter, etid = 0..127): TMEM -> (+bias, +cond, +resid, +sum, /div) -> bf16 -> HBM; they walk
// the same static tile sequence as the other roles.  (The generic first-generation epilogue that also did the
// ConvTranspose1d scatter was retired in favour of epilogue_pipe / epilogue_up below; epilogue_fir serves amp_fir.cuh.)
// The static tile sequence of a CTA: w = w0, w0 + wstep, ...  (w = time tile * n_tiles + column tile).  Plain launch: CTA i
// starts at tile i and strides by the grid.  Cluster launch (a.cl = n_tiles CTAs per cluster): the CTAs of a cluster take the
// n_tiles column tiles of the SAME time tile, cluster j starts at time tile j and strides by the number of clusters.
__device__ __forceinline__ int tile_w0(const TcArgs& a) {
  return a.cl > 1 ? (int)(blockIdx.x / a.cl) * a.n_tiles + (int)(blockIdx.x % a.cl) : (int)blockIdx.x;
}
__device__ __forceinline__ int tile_wstep(const TcArgs& a) {
  return a.cl > 1 ? (int)(gridDim.x / a.cl) * a.n_tiles : (int)gridDim.x;
}

struct TileCursor {            // monotone walk over the per-utterance tile prefix table
  const int* prefix;
  int b = 0;
  __device__ __forceinline__ void locate(int w, int n_tiles, int& bb, int& t0, int& nt) {
    int r = w;
    nt = 0;
    if (n_tiles > 1) {               // skip the integer division for single-column-tile layers
      r = w / n_tiles;
      nt = w - r * n_tiles;
    }
    while (r >= prefix[b + 1]) ++b;
    bb = b;
    t0 = (r - prefix[b]) * M_TILE;
  }
};

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Lean conv-mode epilogue (no ConvTranspose scatter, one column tile): TMEM -> +bias(+cond) (+resid) (+sum) (x1/div)
// -> bf16 -> 16-byte stores.  32 accumulator columns per step with the residual / running-sum rows of the step
// already in flight when the TMEM load is issued; packed f32x2 arithmetic.
template <int NCOL, bool HAS_R, bool HAS_Q>
__device__ __forceinline__ void epi_step(const TcArgs& a, const float* bs, uint32_t taddr, int cb0, int ngs, int code,
                                         const __nv_bfloat16* resid, const __nv_bfloat16* accin, __nv_bfloat16* outp,
                                         int rowoff, int gstride, u64 rdiv2) {
  constexpr int NG = NCOL / 8;
  uint4 rr[HAS_R ? NG : 1], qq[HAS_Q ? NG : 1];
  const int o0 = (cb0 >> 3) * gstride + rowoff;
  if (code == 1) {
    if constexpr (HAS_R) {
#pragma unroll
      for (int kk = 0; kk < NG; ++kk)      // dead groups re-read the last live one (no predicated array slots)
        rr[kk] = *reinterpret_cast<const uint4*>(resid + o0 + min(kk, ngs - 1) * gstride);
    }
    if constexpr (HAS_Q) {
#pragma unroll
      for (int kk = 0; kk < NG; ++kk)
        qq[kk] = *reinterpret_cast<const uint4*>(accin + o0 + min(kk, ngs - 1) * gstride);
    }
  }
  uint32_t v[NCOL];
  if constexpr (NCOL == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr + (uint32_t)cb0));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + (uint32_t)cb0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (code == 0) return;
#pragma unroll
  for (int kk = 0; kk < NG; ++kk) {
    if (kk >= ngs) break;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (code == 1) {
      const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(bs + cb0 + kk * 8);
      const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(bs + cb0 + kk * 8 + 4);
      u64 f[4];
      f[0] = add2(pk(__uint_as_float(v[kk * 8 + 0]), __uint_as_float(v[kk * 8 + 1])), b01.x);
      f[1] = add2(pk(__uint_as_float(v[kk * 8 + 2]), __uint_as_float(v[kk * 8 + 3])), b01.y);
      f[2] = add2(pk(__uint_as_float(v[kk * 8 + 4]), __uint_as_float(v[kk * 8 + 5])), b23.x);
      f[3] = add2(pk(__uint_as_float(v[kk * 8 + 6]), __uint_as_float(v[kk * 8 + 7])), b23.y);
      if constexpr (HAS_R) {
        f[0] = add2(f[0], bf2_to_f2(rr[kk].x)); f[1] = add2(f[1], bf2_to_f2(rr[kk].y));
        f[2] = add2(f[2], bf2_to_f2(rr[kk].z)); f[3] = add2(f[3], bf2_to_f2(rr[kk].w));
      }
      if constexpr (HAS_Q) {
        f[0] = add2(f[0], bf2_to_f2(qq[kk].x)); f[1] = add2(f[1], bf2_to_f2(qq[kk].y));
        f[2] = add2(f[2], bf2_to_f2(qq[kk].z)); f[3] = add2(f[3], bf2_to_f2(qq[kk].w));
      }
      if (a.div != 1.0f) {
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = mul2(f[e], rdiv2);
      }
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo, hi;
        upk(f[e], lo, hi);
        __nv_bfloat162 pb = __floats2bfloat162_rn(lo, hi);
        w[e] = *reinterpret_cast<uint32_t*>(&pb);
      }
      o = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(outp + o0 + kk * gstride) = o;
  }
}

template <bool HAS_R, bool HAS_Q, int STEPW, bool MULTI>
__device__ __forceinline__ void epilogue_fir_t(const TcArgs& a, float* bias_s, const int* prefix, uint32_t bar_accfull0,
                                             uint32_t bar_accempty0, uint32_t tmem, int nacc, int total_tiles, int q,
                                             int lane, int etid) {
  const int n_tile = a.n_tile, n_tiles = MULTI ? a.n_tiles : 1;   // MULTI = several column tiles per time tile
  const int cg_total = a.Cout >> 3;
  const int gstride = a.Tstride * 8;
  const float rdiv = 1.0f / a.div;
  const u64 rdiv2 = pk(rdiv, rdiv);
  TileCursor cur{prefix};
  int it = 0, last_b = -1, last_nt = -1;
  for (int w = tile_w0(a); w < total_tiles; w += tile_wstep(a), ++it) {
    int b, t0, nt;
    cur.locate(w, n_tiles, b, t0, nt);
    const int T = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
    const int as = (nacc == 2) ? (it & 1) : 0;
    const int ause = (nacc == 2) ? (it >> 1) : it;
    if (!MULTI) nt = 0;
    const int cgn0 = nt * (n_tile >> 3);                               // first column group of this tile
    const int ng = min(n_tile >> 3, cg_total - cgn0);                  // live column groups
    if (last_b < 0 || nt != last_nt || (a.bias_b && b != last_b)) {    // bias row changes with (utterance, column tile)
      asm volatile("bar.sync 1, 128;" ::: "memory");                   // every warp is done with the previous row
      for (int i = etid; i < n_tile; i += 128) {
        const int co = nt * n_tile + i;
        float v = 0.f;
        if (co < a.Cout) {
          v = __ldg(a.bias + co);
          if (a.bias_b) v += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
        }
        bias_s[i] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      last_b = b;
      last_nt = nt;
    }
    const size_t ubase = ((size_t)b * cg_total + cgn0) * a.Tstride * 8;   // this (utterance, column tile)'s first group
    const __nv_bfloat16* resid = a.resid ? a.resid + ubase : nullptr;
    const __nv_bfloat16* accin = a.acc_in ? a.acc_in + ubase : nullptr;
    __nv_bfloat16* outp = a.out + ubase;
    if (resid || accin) {
      // warm L2 with this tile's residual / running-sum rows while its MMAs are still running
#pragma unroll 1
      for (int mb = 0; mb < 2; ++mb) {
        const int t = t0 + mb * 128 + q * 32 + lane;
        if (t >= T) continue;
        int o = t * 8;
#pragma unroll 1
        for (int g = 0; g < ng; ++g, o += gstride) {
          if (resid) asm volatile("prefetch.global.L2 [%0];" ::"l"(resid + o));
          if (accin) asm volatile("prefetch.global.L2 [%0];" ::"l"(accin + o));
        }
      }
    }
    if (q == 0) mbar_wait_relaxed(bar_accfull0 + 8 * as, ause & 1, a.Cin <= 96 ? 600u : 120u);
    asm volatile("bar.sync 2, 128;" ::: "memory");
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * n_tile);
#pragma unroll 1
    for (int mb = 0; mb < (BVG_DBGBIT(a, 8) ? 0 : 2); ++mb) {
      const int t = t0 + mb * 128 + q * 32 + lane;
      const int code = (t >= a.Tmax || t < a.st_lo || t >= a.st_hi) ? 0 : (t < T ? 1 : 2);
      constexpr int STEP = (HAS_Q || STEPW == 16) ? 16 : 32;   // fewer rows in flight where registers are short
#pragma unroll 1
      for (int cb0 = 0; cb0 < n_tile; cb0 += STEP) {
        const int ngs = ng - (cb0 >> 3);
        if (ngs <= 0) break;                                            // warp-uniform
        if (n_tile - cb0 >= STEP)
          epi_step<STEP, HAS_R, HAS_Q>(a, bias_s, taddr + (uint32_t)(mb * n_tile), cb0, ngs, code, resid, accin, outp, t * 8, gstride, rdiv2);
        else
          epi_step<16, HAS_R, HAS_Q>(a, bias_s, taddr + (uint32_t)(mb * n_tile), cb0, ngs, code, resid, accin, outp, t * 8, gstride, rdiv2);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_accempty0 + 8 * as);
    if (q == 0) TC_TRACE(a, 3300 + it, 6);
    // un-activated consumers (ConvTranspose1d) read one row past the end: keep rows [T, T+8) zero
    // when the utterance ends exactly on this tile's boundary (otherwise they were zeroed above)
    if (T == t0 + M_TILE && T < a.Tmax && T < a.st_hi && q == 0) {
      for (int i = lane; i < ng * 8; i += 32) {
        const int g = i >> 3, r = T + (i & 7);
        if (r < a.Tmax)
          *reinterpret_cast<uint4*>(a.out + (((size_t)b * cg_total + cgn0 + g) * a.Tstride + r) * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

template <int STEPW, bool MULTI = false>
__device__ __forceinline__ void epilogue_fir(const TcArgs& a, float* bias_s, const int* prefix, uint32_t bar_accfull0,
                                             uint32_t bar_accempty0, uint32_t tmem, int nacc, int total_tiles, int q,
                                             int lane, int etid) {
  if (a.resid) {
    if (a.acc_in) epilogue_fir_t<true, true, STEPW, MULTI>(a, bias_s, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
    else epilogue_fir_t<true, false, STEPW, MULTI>(a, bias_s, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
  } else {
    if (a.acc_in) epilogue_fir_t<false, true, STEPW, MULTI>(a, bias_s, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
    else epilogue_fir_t<false, false, STEPW, MULTI>(a, bias_s, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
  }
}


// ------------------------------------------------------------------------------ pipelined conv-mode epilogue
// The residual / running-sum rows of a tile were written several launches ago and are not in L2 any more.  Loading
// them into registers right before the TMEM read leaves ~4 KB in flight per SM (128 threads x 2 x 16 B — the epilogue
// warps have 64 registers), and per-launch events showed every layer with a residual 50-110 us slower than its twin
// without one: the epilogue, not the activation or the MMAs, set the pace.  Here the rows go through shared memory
// instead: each epilogue thread owns R_SLOTS 16-byte slots ([slot][thread] layout, conflict-free), fills them with
// cp.async — the first LA steps before it even waits for the accumulator, the others as steps retire — and reads a
// step's rows back with one LDS each just before the math.  32 KB per SM in flight, no registers held; the rows of the
// NEXT tile are L2-prefetched meanwhile.
template <int NG, bool HAS_R, bool HAS_Q>
__device__ __forceinline__ void epi_rows_finish(const uint4* rr, const uint4* qq, const TcArgs& a, const float* bs,
                                                uint32_t taddr, int cb0, int ngs, int code, __nv_bfloat16* outp, int o0,
                                                int gstride, u64 rdiv2) {
  uint32_t v[NG * 8];
  if constexpr (NG == 4) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr + (uint32_t)cb0));
  } else {
    static_assert(NG == 2, "epilogue step = 16 or 32 accumulator columns");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr + (uint32_t)cb0));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (code == 0) return;
#pragma unroll
  for (int kk = 0; kk < NG; ++kk) {
    if (kk >= ngs) break;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (code == 1) {
      const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(bs + cb0 + kk * 8);
      const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(bs + cb0 + kk * 8 + 4);
      u64 f[4];
      f[0] = add2(pk(__uint_as_float(v[kk * 8 + 0]), __uint_as_float(v[kk * 8 + 1])), b01.x);
      f[1] = add2(pk(__uint_as_float(v[kk * 8 + 2]), __uint_as_float(v[kk * 8 + 3])), b01.y);
      f[2] = add2(pk(__uint_as_float(v[kk * 8 + 4]), __uint_as_float(v[kk * 8 + 5])), b23.x);
      f[3] = add2(pk(__uint_as_float(v[kk * 8 + 6]), __uint_as_float(v[kk * 8 + 7])), b23.y);
      if constexpr (HAS_R) {
        f[0] = add2(f[0], bf2_to_f2(rr[kk].x)); f[1] = add2(f[1], bf2_to_f2(rr[kk].y));
        f[2] = add2(f[2], bf2_to_f2(rr[kk].z)); f[3] = add2(f[3], bf2_to_f2(rr[kk].w));
      }
      if constexpr (HAS_Q) {
        f[0] = add2(f[0], bf2_to_f2(qq[kk].x)); f[1] = add2(f[1], bf2_to_f2(qq[kk].y));
        f[2] = add2(f[2], bf2_to_f2(qq[kk].z)); f[3] = add2(f[3], bf2_to_f2(qq[kk].w));
      }
      if (a.div != 1.0f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = mul2(f[i], rdiv2);
      }
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk(f[i], lo, hi);
        __nv_bfloat162 pb = __floats2bfloat162_rn(lo, hi);
        w[i] = *reinterpret_cast<uint32_t*>(&pb);
      }
      o = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (!BVG_DBGBIT(a, 64)) *reinterpret_cast<uint4*>(outp + o0 + kk * gstride) = o;
  }
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NG channel groups (8 accumulator columns each) per step; LA steps of row copies in flight
// (LA * NG * (HAS_R + HAS_Q) = R_SLOTS slots per thread).
template <int NG, bool HAS_R, bool HAS_Q, bool MULTI>
__device__ __forceinline__ void epilogue_pipe_t(const TcArgs& a, float* bias_s, uint8_t* rstage, const int* prefix,
                                                uint32_t bar_accfull0, uint32_t bar_accempty0, uint32_t tmem, int nacc,
                                                int total_tiles, int q, int lane, int etid) {
  constexpr int NSTREAM = (HAS_R ? 1 : 0) + (HAS_Q ? 1 : 0);
  constexpr int LA = NSTREAM ? R_SLOTS / (NG * NSTREAM) : 1;
  const int n_tile = a.n_tile, n_tiles = MULTI ? a.n_tiles : 1;
  const int cg_total = a.Cout >> 3;
  const int gstride = a.Tstride * 8;
  const float rdiv = 1.0f / a.div;
  const u64 rdiv2 = pk(rdiv, rdiv);
  // slot (step ring position r, stream st, group kk) of this thread
  uint8_t* my_stage = rstage + etid * 16;
  const uint32_t my_stage_u = smem_u32(my_stage);
  auto slot_off = [&](int r, int st, int kk) { return (uint32_t)(((r * NSTREAM + st) * NG + kk) * 128 * 16); };
  TileCursor cur{prefix};
  int it = 0, last_b = -1, last_nt = -1;
  int T_next = 0;                      // length of the next tile's utterance, requested one tile early (L2 round trip)
  auto prefetch_tile = [&](int w) {     // only the utterance length is fetched ahead (an L2 prefetch of the rows
    if (w >= total_tiles) return;       // themselves was measured to cost more than it saves once they are staged)
    TileCursor c2 = cur;
    int b, t0, nt;
    c2.locate(w, n_tiles, b, t0, nt);
    T_next = a.lengths ? __ldg(a.lengths + b) * a.rate : a.Tmax;
  };
  prefetch_tile(tile_w0(a));
  for (int w = tile_w0(a); w < total_tiles; w += tile_wstep(a), ++it) {
    int b, t0, nt;
    cur.locate(w, n_tiles, b, t0, nt);
    const int T = T_next;
    const int as = (nacc == 2) ? (it & 1) : 0;
    const int ause = (nacc == 2) ? (it >> 1) : it;
    if (!MULTI) nt = 0;
    const int cgn0 = nt * (n_tile >> 3);                               // first column group of this tile
    const int ng = min(n_tile >> 3, cg_total - cgn0);                  // live column groups
    if (last_b < 0 || nt != last_nt || (a.bias_b && b != last_b)) {    // bias row changes with (utterance, column tile)
      asm volatile("bar.sync 1, 128;" ::: "memory");                   // every warp is done with the previous row
      for (int i = etid; i < n_tile; i += 128) {
        const int co = nt * n_tile + i;
        float v = 0.f;
        if (co < a.Cout) {
          v = __ldg(a.bias + co);
          if (a.bias_b) v += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
        }
        bias_s[i] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      last_b = b;
      last_nt = nt;
    }
    const size_t ubase = ((size_t)b * cg_total + cgn0) * a.Tstride * 8;   // this (utterance, column tile)'s first group
    const __nv_bfloat16* resid = HAS_R ? a.resid + ubase : nullptr;
    const __nv_bfloat16* accin = HAS_Q ? a.acc_in + ubase : nullptr;
    __nv_bfloat16* outp = a.out + ubase;
    // steps: s -> (mb, cs); spm steps per M block
    const int spm = (ng + NG - 1) / NG;
    const int nst = BVG_DBGBIT(a, 8) ? 0 : 2 * spm;
    const int tr0 = t0 + q * 32 + lane, tr1 = tr0 + 128;
    const int code0 = (tr0 >= a.Tmax || tr0 < a.st_lo || tr0 >= a.st_hi) ? 0 : (tr0 < T ? 1 : 2);
    const int code1 = (tr1 >= a.Tmax || tr1 < a.st_lo || tr1 >= a.st_hi) ? 0 : (tr1 < T ? 1 : 2);
    // request the rows of step s into ring position s % LA (one commit group per step, empty ones included)
    auto issue = [&](int s) {
      if constexpr (NSTREAM > 0) {
        if (s < nst) {
          const int mb = s >= spm ? 1 : 0, cs = s - mb * spm;
          if ((mb ? code1 : code0) == 1 && !BVG_DBGBIT(a, 32)) {
            const int o0 = cs * NG * gstride + (mb ? tr1 : tr0) * 8, ngs = ng - cs * NG, r = s % LA;
#pragma unroll
            for (int kk = 0; kk < NG; ++kk) {          // dead groups re-read the last live one
              const int o = o0 + min(kk, ngs - 1) * gstride;
              if constexpr (HAS_R) cp_async16(my_stage_u + slot_off(r, 0, kk), resid + o);
              if constexpr (HAS_Q) cp_async16(my_stage_u + slot_off(r, HAS_R ? 1 : 0, kk), accin + o);
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
#pragma unroll 1
    for (int s = 0; s < LA; ++s) issue(s);
    prefetch_tile(w + tile_wstep(a));
    if (q == 0) mbar_wait_relaxed(bar_accfull0 + 8 * as, ause & 1, a.Cin <= 96 ? 600u : 120u);
    asm volatile("bar.sync 2, 128;" ::: "memory");
    tc_fence_after();
    if (q == 0) TC_TRACE(a, 3200 + it, 5);
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * n_tile);
#pragma unroll 1
    for (int s = 0; s < nst; ++s) {
      const int mb = s >= spm ? 1 : 0, cs = s - mb * spm;
      uint4 rr[NG], qq[NG];
      if constexpr (NSTREAM > 0) {
        cp_async_wait<LA - 1>();                       // step s's group (and every older one) has landed
        const int r = s % LA;
#pragma unroll
        for (int kk = 0; kk < NG; ++kk) {
          if constexpr (HAS_R) rr[kk] = *reinterpret_cast<const uint4*>(my_stage + slot_off(r, 0, kk));
          if constexpr (HAS_Q) qq[kk] = *reinterpret_cast<const uint4*>(my_stage + slot_off(r, HAS_R ? 1 : 0, kk));
        }
      }
      epi_rows_finish<NG, HAS_R, HAS_Q>(rr, qq, a, bias_s, taddr + (uint32_t)(mb * n_tile), cs * NG * 8, ng - cs * NG,
                                        mb ? code1 : code0, outp, cs * NG * gstride + (mb ? tr1 : tr0) * 8, gstride,
                                        rdiv2);
      issue(s + LA);
    }
    if constexpr (NSTREAM > 0) cp_async_wait<0>();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_accempty0 + 8 * as);
    if (q == 0) TC_TRACE(a, 3300 + it, 6);
    // un-activated consumers (ConvTranspose1d) read one row past the end: keep rows [T, T+8) zero
    // when the utterance ends exactly on this tile's boundary (otherwise they were zeroed above)
    if (T == t0 + M_TILE && T < a.Tmax && T < a.st_hi && q == 0) {
      for (int i = lane; i < ng * 8; i += 32) {
        const int g = i >> 3, r = T + (i & 7);
        if (r < a.Tmax)
          *reinterpret_cast<uint4*>(a.out + (((size_t)b * cg_total + cgn0 + g) * a.Tstride + r) * 8) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

template <bool MULTI, bool PLAIN_ONLY>
__device__ __forceinline__ void epilogue_pipe(const TcArgs& a, float* bias_s, uint8_t* rstage, const int* prefix,
                                              uint32_t bar_accfull0, uint32_t bar_accempty0, uint32_t tmem, int nacc,
                                              int total_tiles, int q, int lane, int etid) {
  if constexpr (PLAIN_ONLY) {
    epilogue_pipe_t<4, false, false, MULTI>(a, bias_s, rstage, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
    return;
  }
  if (a.resid) {
    if (a.acc_in) epilogue_pipe_t<2, true, true, MULTI>(a, bias_s, rstage, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
    else epilogue_pipe_t<4, true, false, MULTI>(a, bias_s, rstage, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
  } else {
    if (a.acc_in) epilogue_pipe_t<2, false, true, MULTI>(a, bias_s, rstage, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
    else epilogue_pipe_t<4, false, false, MULTI>(a, bias_s, rstage, prefix, bar_accfull0, bar_accempty0, tmem, nacc, total_tiles, q, lane, etid);
  }
}

// ------------------------------------------------------------------------------ ConvTranspose1d epilogue
// Phase scatter of the (k/u)-tap implicit GEMM (models.py:157-163,232-236): accumulator column n = r*cphase + co holds
// output time q*up + r - pad of channel co, q = the row's input time.  ncu on the generic first-generation epilogue showed these
// launches idle on every pipe (issue 22 %, DRAM 16 %) behind ONE serial chain: ~2100 instructions per epilogue warp and
// tile (an integer division per column group, 16-column steps).  Here: 32 columns per TMEM load, phase / channel-group
// counters instead of divisions, packed adds, the next tile's utterance length requested a tile early.
__device__ __forceinline__ void epilogue_up(const TcArgs& a, float* bias_s, const int* prefix, uint32_t bar_accfull0,
                                            uint32_t bar_accempty0, uint32_t tmem, int nacc, int total_tiles, int extra,
                                            int q, int lane, int etid) {
  const int n_tile = a.n_tile, n_tiles = a.n_tiles;
  const int gpp = a.cphase >> 3;                       // channel groups per phase
  const int ncg = (a.up * a.cphase) >> 3;              // column groups in all
  const int gstride = a.Tstride * 8;
  TileCursor cur{prefix};
  int it = 0;
  int Tin_next = a.Tmax;
  auto ahead = [&](int w) {
    if (w >= total_tiles || !a.lengths) return;
    TileCursor c2 = cur;
    int b, t0, nt;
    c2.locate(w, n_tiles, b, t0, nt);
    Tin_next = __ldg(a.lengths + b) * a.rate;
  };
  ahead(tile_w0(a));
  for (int w = tile_w0(a); w < total_tiles; w += tile_wstep(a), ++it) {
    int b, t0, nt;
    cur.locate(w, n_tiles, b, t0, nt);
    const int Tin = Tin_next;
    ahead(w + tile_wstep(a));
    const int T = Tin + extra, Tout = Tin * a.up;
    const int as = (nacc == 2) ? (it & 1) : 0;
    const int ause = (nacc == 2) ? (it >> 1) : it;
    float* bs = bias_s + as * 256;
    if (a.bias_b || n_tiles > 1 || it < nacc) {                       // bias row changes with (b, nt) only
      asm volatile("bar.sync 1, 128;" ::: "memory");                  // previous user of bias_s[as] is done
      for (int i = etid; i < n_tile; i += 128) {
        int co = nt * n_tile + i;
        float v = 0.f;
        if (co < a.up * a.cphase) {
          co %= a.cphase;
          v = __ldg(a.bias + co);
          if (a.bias_b) v += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
        }
        bs[i] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    __nv_bfloat16* outp = a.out + (size_t)b * gpp * a.Tstride * 8;   // this utterance's block
    const int cgn0 = nt * (n_tile >> 3);
    const int ng = min(n_tile >> 3, ncg - cgn0);                      // live column groups of this tile
    const int r0 = cgn0 / gpp, cg0 = cgn0 - r0 * gpp;                 // phase / channel group of the first one
    if (q == 0) mbar_wait_relaxed(bar_accfull0 + 8 * as, ause & 1, a.Cin <= 96 ? 600u : 120u);
    asm volatile("bar.sync 2, 128;" ::: "memory");
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * n_tile);
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
      const int tq = t0 + mb * 128 + q * 32 + lane;                   // input time of this thread's row
      const int tb = tq * a.up - a.pad;                               // output time of phase 0
      int r = r0, cg = cg0;
#pragma unroll 1
      for (int g0 = 0; g0 < ng; g0 += 4) {                            // 4 column groups = 32 accumulator columns
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)(mb * n_tile + g0 * 8)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (g0 + kk < ng) {
            const int to = tb + r;
            if (tq < T && to >= 0 && to < Tout) {
              const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(bs + (g0 + kk) * 8);
              const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(bs + (g0 + kk) * 8 + 4);
              const u64 f0 = add2(pk(__uint_as_float(v[kk * 8 + 0]), __uint_as_float(v[kk * 8 + 1])), b01.x);
              const u64 f1 = add2(pk(__uint_as_float(v[kk * 8 + 2]), __uint_as_float(v[kk * 8 + 3])), b01.y);
              const u64 f2 = add2(pk(__uint_as_float(v[kk * 8 + 4]), __uint_as_float(v[kk * 8 + 5])), b23.x);
              const u64 f3 = add2(pk(__uint_as_float(v[kk * 8 + 6]), __uint_as_float(v[kk * 8 + 7])), b23.y);
              float l0, h0, l1, h1, l2, h2, l3, h3;
              upk(f0, l0, h0); upk(f1, l1, h1); upk(f2, l2, h2); upk(f3, l3, h3);
              __nv_bfloat162 p0 = __floats2bfloat162_rn(l0, h0), p1 = __floats2bfloat162_rn(l1, h1);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(l2, h2), p3 = __floats2bfloat162_rn(l3, h3);
              *reinterpret_cast<uint4*>(outp + cg * gstride + to * 8) =
                  make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                             *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
            }
            if (++cg == gpp) { cg = 0; ++r; }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_accempty0 + 8 * as);
  }
}

// ------------------------------------------------------------------------------ the kernel
// Persistent: grid = min(#tiles, #SMs); every role walks the same static tile sequence
// w = blockIdx.x, +gridDim.x, ... (tile = 256 rows x n_tile columns of one utterance; column
// tile fastest so that CTAs running side by side share x in L2).  The x / z / weight rings and
// the TMEM accumulator stages keep running across tile boundaries, so the epilogue of tile i,
// the MMAs of tile i+1 and the activation of tile i+1/i+2 overlap.
template <int REGS>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }


// EPI selects the one epilogue compiled into an instantiation (0 = conv, one column tile; 1 = conv, several column
// tiles; 2 = ConvTranspose1d phase scatter): with all of them behind run-time branches in one function, adding a path
// changed the register allocation of the others (measured: +30 us on the C = 24 residual layers).
template <int L, bool ACT, bool RM, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
k_amp_tc(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmr,
         const __grid_constant__ CUtensorMap tmq, const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = a.n_tile, n_tiles = a.n_tiles;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + OFF_BAR;
  const int NX = a.nx, NZ = a.nz, W_STAGES = a.wst;
  const int OFF_Z = off_z(NX), OFF_W = off_w(NX, NZ);
  auto BAR_XFULL = [&](int i) { return bar0 + 8 * (0 + i); };
  auto BAR_XEMPTY = [&](int i) { return bar0 + 8 * (NX_MAX + i); };
  auto BAR_ZFULL = [&](int i) { return bar0 + 8 * (2 * NX_MAX + i); };
  auto BAR_ZEMPTY = [&](int i) { return bar0 + 8 * (2 * NX_MAX + NZ_MAX + i); };
  auto BAR_WFULL = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + i); };
  auto BAR_WEMPTY = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + W_STAGES_MAX + i); };
  auto BAR_ACCFULL = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + 2 * W_STAGES_MAX + i); };
  auto BAR_ACCEMPTY = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + 2 * W_STAGES_MAX + 2 + i); };
  auto BAR_RFULL = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + 2 * W_STAGES_MAX + 4 + i); };
  auto BAR_REMPTY = [&](int i) { return bar0 + 8 * (2 * NX_MAX + 2 * NZ_MAX + 2 * W_STAGES_MAX + 6 + i); };
  const int OFF_R = off_r(NX, NZ, W_STAGES);
  const int nstreams_r = RM ? (a.rmma_r ? 1 : 0) + (a.rmma_q ? 1 : 0) : 0;   // residual-like streams added by identity MMAs
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  int* prefix = reinterpret_cast<int*>(smem + OFF_PREFIX);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  const int extra = a.up ? a.K - 1 : 0;      // transposed mode also runs the K-1 rows past the end
  const int hc = a.dil * (a.K - 1) / 2;
  const int NCH = (a.Cin + KC - 1) / KC;
  const int tile_bytes = n_tile * 64;
  const int tps = a.taps_per_stage;
  const int spc = (a.K + tps - 1) / tps;
  const int nacc = (4 * n_tile <= 512) ? 2 : 1;

  // ---- prologue: tile prefix table, barriers, TMEM
  if (warp == 0) {
    int run = 0;
    for (int b0 = 0; b0 < a.B; b0 += 32) {
      const int b = b0 + lane;
      int inc = 0;
      if (b < a.B) {
        const int Tin = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
        inc = (Tin + extra + M_TILE - 1) / M_TILE;
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
  if (warp == NW_ACT && lane == 0) {
    // two MMA issuer warps (one per M block): every "the MMAs that read this are done" barrier takes two commits
    // activation warps whose channel group lies past C_in in EVERY chunk (C_in < 32: C = 24 -> the 4 warps of group 3)
    // take no part in the ring hand-shakes at all
    const int n_act = NW_ACT - 4 * max(0, 4 - (a.Cin >> 3));
    for (int i = 0; i < NX_MAX; ++i) { mbar_init(BAR_XFULL(i), 1); mbar_init(BAR_XEMPTY(i), ACT ? n_act : 2); }
    for (int i = 0; i < NZ_MAX; ++i) { mbar_init(BAR_ZFULL(i), ACT ? n_act : NW_ACT); mbar_init(BAR_ZEMPTY(i), 2 * ((ACT && a.cl > 1) ? a.cl : 1)); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR_ACCFULL(i), 2); mbar_init(BAR_ACCEMPTY(i), 4); }
    for (int i = 0; i < R_RING; ++i) { mbar_init(BAR_RFULL(i), 1); mbar_init(BAR_REMPTY(i), 2); }
    for (int i = 0; i < W_STAGES_MAX; ++i) { mbar_init(BAR_WFULL(i), 1); mbar_init(BAR_WEMPTY(i), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
  }
  if (warp == NW_ACT + 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(s_base + OFF_TMEM), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int total_tiles = prefix[a.B] * n_tiles;
  // cluster mode: the n_tiles column tiles of a time tile run in one cluster; chunk c of the tile is activated by the CTA
  // of rank c % ncl, which writes the z tile into every CTA's ring (distributed shared memory) and arrives on every
  // CTA's "z full" barrier; "z empty" takes a multicast commit from every CTA's MMA issuers
  const int ncl = (ACT && a.cl > 1) ? a.cl : 1;
  const uint32_t crank = ncl > 1 ? cluster_ctarank() : 0u;
  const int w0 = tile_w0(a), wstep = tile_wstep(a);
  if (ncl > 1) cluster_sync_all();     // every CTA's barriers are initialised before anyone arrives remotely
  pdl_wait();          // the prologue above only read launch constants (lengths, parameters); activations from here on
  if (warp == 0) TC_TRACE(a, 1, 9);

  // The x producer's loop: channel groups [g_lo, g_hi) of every chunk; `lead` announces the bytes.  Activated layers run
  // it from ONE thread (groups 0-3).  Un-activated layers (conv_pre, ConvTranspose1d) feed the MMAs straight from the x
  // ring and their 16 activation warps have nothing to do, so three of them take a channel group each beside the
  // producer warp: one thread gets a copy out every ~0.3 us whatever the ring depth, and copies from different WARPS
  // proceed in parallel (tools/bulk_probe.cu) — the narrow ConvTranspose1d launches were bound by exactly that rate.
  auto x_loop = [&](const int g_lo, const int g_hi, const bool lead) {
        TileCursor cur{prefix};
        int xb = 0, xph = 0;
        [[maybe_unused]] int trx = 0;
        for (int w = w0; w < total_tiles; w += wstep) {
          int b, t0, nt;
          cur.locate(w, n_tiles, b, t0, nt);
          for (int c = 0; c < NCH; ++c) {
            if (ncl > 1 && (uint32_t)(c % ncl) != crank) continue;    // that chunk is activated by a peer CTA
            mbar_wait_relaxed(BAR_XEMPTY(xb), xph ^ 1, a.Cin <= 96 ? 600u : 150u);
            if (lead && trx < 480) TC_TRACE(a, 2500 + trx, 11);
            ++trx;
            // activated layers never read the channel groups past C_in (their warps skip); plain convs feed the
            // x tile to the MMA as it is and need the TMA zero fill of those groups
            const int lg = ACT ? min(4, (a.Cin >> 3) - c * 4) : 4;
            if (lead) mbar_expect_tx(BAR_XFULL(xb), (uint32_t)lg * (X_TX_BYTES / 4));
            const uint32_t dst = s_base + OFF_X + xb * X_BUF_BYTES;
            // The XR rows of one (utterance, channel group) are contiguous in the blocked layout: a tile whose window lies
            // inside the buffer takes one 1-D bulk copy per group (5 KB) instead of two tensor boxes, which the TMA unit
            // walks 16-byte row by row (~1 row / clk).  Windows that leave [0, Tmax) keep the tensor path for its zero fill.
            if (a.xin && t0 - X_LEAD >= 0 && t0 - X_LEAD + XR <= a.Tmax && (ACT || c * 4 + 4 <= a.xgroups)) {
              const __nv_bfloat16* src = a.xin + (((size_t)b * a.xgroups + c * 4) * a.Tmax + (t0 - X_LEAD)) * 8;
#pragma unroll
              for (int kg = 0; kg < 4; ++kg)
                if (kg < lg && kg >= g_lo && kg < g_hi) bulk_load(dst + kg * (XRA * 16), src + (size_t)kg * a.Tmax * 8, XR * 16, BAR_XFULL(xb));
            } else {
#pragma unroll
              for (int kg = 0; kg < 4; ++kg)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  if (kg < lg && kg >= g_lo && kg < g_hi) tma_load_4d(dst + kg * (XRA * 16) + h * (BOXR * 16), &tmx, 0, t0 - X_LEAD + h * BOXR, c * 4 + kg, b,
                              BAR_XFULL(xb));
            }
            if (++xb == NX) { xb = 0; xph ^= 1; }
          }
        }
  };
  if (warp < NW_ACT) {
    // ===================== activation warps =====================
    reg_inc<96>();
    if (ACT) {
      ActCtx k;
      load_taps(k, a);
      // 4 channel groups x 4 warps; a warp's 8 runs of L rows span 8L rows and yield V = 8L-6 z rows
      constexpr int V = 8 * L - 6;
      // warp -> (channel group kg, time quarter wq).  Warps w, w+4, w+8, w+12 share one SM sub-partition (scheduler
      // = warp % 4): they take the four channel groups of ONE time quarter, so that the groups past C_in (C = 24: kg 3;
      // C = 48, second chunk: kg 2-3), whose warps idle, are spread over all four schedulers instead of leaving one
      // scheduler empty and the other three issue-bound.
      // Measured: with no idle groups (C a multiple of 32) the transposed assignment is the faster one by ~5 %.
      const bool spread = BVG_DBGBIT(a, 256) ? false : (BVG_DBGBIT(a, 512) ? true : (a.Cin & 31) != 0);
      const int kg = spread ? (warp >> 2) : (warp & 3), wq = spread ? (warp & 3) : (warp >> 2);
      const int g = lane >> 2, p = lane & 3;
      const int ZW = M_TILE + 2 * hc;
      const int vlo = wq * V, vhi = min(vlo + V, ZR);
      const int rowS = vlo - 3 + g * L;
      uint32_t smask = 0;                 // bit r: local row rowS + r lies in [vlo, vhi)
#pragma unroll
      for (int r = 0; r < L; ++r)
        if (rowS + r >= vlo && rowS + r < vhi) smask |= 1u << r;
      // Channel groups past C_in (C = 24: group 3; C = 48: groups 2-3 of the second chunk) carry zero weights: their
      // warps skip the arithmetic.  Their z rows are zeroed once so that no stale NaN pattern reaches the MMA.
      const int live_groups = a.Cin >> 3;
      uint32_t peer_base[2] = {0u, 0u};           // shared-window base of the other CTAs of my cluster
      if (ncl > 1) {
        peer_base[0] = mapa_u32(s_base, (crank + 1) % ncl);
        if (ncl > 2) peer_base[1] = mapa_u32(s_base, (crank + 2) % ncl);
      }
      if ((NCH - 1) * 4 + kg >= live_groups) {
        for (int s = 0; s < NZ; ++s) {
          uint32_t* zk = reinterpret_cast<uint32_t*>(smem + OFF_Z + s * Z_BUF_BYTES) + kg * (ZR * 4) + p;
#pragma unroll
          for (int r = 0; r < L; ++r)
            if ((smask >> r) & 1u) zk[(rowS + r) * 4] = 0u;
        }
      }
      // Shared memory leaves the SM little L1, so every global load below is an L2 round trip (0.3-1 us
      // under load).  The utterance length of the NEXT tile and the snake parameters of the NEXT chunk are therefore
      // requested one iteration early (ncu: 40 % of the activation warps' stall samples sat on the length load).
      TileCursor cur{prefix};
      int xb = 0, xph = 0, zb = 0, zph = 0;     // ring slot / phase of the next chunk
      int b = 0, t0 = 0, nt = 0, T = 0;
      const bool never_live = kg >= live_groups;      // idle in every chunk: not counted by the barriers (see their init)
      if (live_groups < 4) {
        // such warps exist (C_in < 32): their zeroed rows above must be visible to the UMMA before the first live arrival
        if (never_live) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 3, 512;" ::: "memory");      // the 16 activation warps
      }
      if (w0 < total_tiles && !never_live) {
        cur.locate(w0, n_tiles, b, t0, nt);
        T = a.lengths ? __ldg(a.lengths + b) * a.rate : a.Tmax;
      }
      float2 a2n = __ldg(reinterpret_cast<const float2*>(a.a2 + (int)crank * KC + kg * 8 + 2 * p));     // my first chunk
      float2 nhbn = __ldg(reinterpret_cast<const float2*>(a.nhb + (int)crank * KC + kg * 8 + 2 * p));
      for (int w = never_live ? total_tiles : w0; w < total_tiles; w += wstep) {
        int b2 = 0, t02 = 0, nt2 = 0, T2 = 0;
        if (w + wstep < total_tiles) {
          cur.locate(w + wstep, n_tiles, b2, t02, nt2);
          T2 = a.lengths ? __ldg(a.lengths + b2) * a.rate : a.Tmax;
        }
        const int m0 = t0 - hc + rowS;
        const int xlo = t0 - X_LEAD;
        // runs whose x window [m0-3, m0+L+2] leaves [0, T) need the replicate clamps (warp-uniform
        // because the run takes part in shuffles)
        const bool edge = __any_sync(0xffffffffu, (m0 - 3 < 0) || (m0 + L + 2 > T - 1));
        for (int c = 0; c < NCH; ++c) {
          if (ncl > 1 && (uint32_t)(c % ncl) != crank) {      // a peer activates this chunk and fills my z slot
            if (++zb == NZ) { zb = 0; zph ^= 1; }
            continue;
          }
          k.a2 = pk(a2n.x, a2n.y);
          k.nhb = pk(nhbn.x, nhbn.y);
          k.hb = pk(-nhbn.x, -nhbn.y);
          if (NCH > 1) {                                        // parameters of my NEXT chunk (ncl apart in cluster mode)
            const int cn = (c + ncl >= NCH) ? (int)crank : c + ncl;
            const int chn = cn * KC + kg * 8 + 2 * p;
            a2n = __ldg(reinterpret_cast<const float2*>(a.a2 + chn));
            nhbn = __ldg(reinterpret_cast<const float2*>(a.nhb + chn));
          }
          // warps without work in this chunk (channel groups past C_in, row spans past the z window) still take part in
          // the ring hand-shakes, but asleep: spinning, they ran NX chunks ahead and then burnt 10-20 % of the SM's
          // issue slots on try_wait loops (ncu: 750-1900 polls per tile on the C = 24 layers)
          const bool live = vlo < ZW && c * 4 + kg < live_groups;
          if (live) {
            mbar_wait(BAR_XFULL(xb), xph);
            mbar_wait(BAR_ZEMPTY(zb), zph ^ 1);
          } else {
            // Nothing to read or write: this warp only owes the two barriers its arrival, and it pays EARLY — as soon as
            // the previous phase of each barrier is complete (so that the arrival counts for this chunk's phase), long
            // before the live warps finish the chunk.  (It used to wait for the x tile and the z slot like a live warp,
            // asleep between polls: every hand-over of a C = 24 / 48 layer then waited for the sleepiest idle warp — the
            // MMA issuer got a chunk 5-9 kcycles after the live warps had finished it, pipeline trace, round 2.)
            mbar_wait_relaxed(BAR_ZFULL(zb), zph ^ 1, 1000);     // (nanosleep <= 500 returns at once on this part)
            mbar_wait_relaxed(BAR_XEMPTY(xb), xph ^ 1, 1000);
          }
          if (live) {
            const uint32_t* xk = reinterpret_cast<const uint32_t*>(smem + OFF_X + xb * X_BUF_BYTES) + kg * (XRA * 4) + p;
            uint32_t* zk = reinterpret_cast<uint32_t*>(smem + OFF_Z + zb * Z_BUF_BYTES) + kg * (ZR * 4) + p;
            if (t0 - hc + vlo >= T) {
              // the warp's whole row span lies past the end of the utterance: conv zero padding
#pragma unroll
              for (int r = 0; r < L; ++r)
                if ((smask >> r) & 1u) zk[(rowS + r) * 4] = 0u;
            } else if (edge) {
              act_run_edge<L>(xk, zk, rowS, smask, m0, xlo, T, k.a2, k.nhb, a, lane);
            } else {
              act_run<L, false>(xk, zk, rowS, smask, m0, xlo, T, k, lane);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // z stores -> async proxy (UMMA, bulk copies)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(BAR_ZFULL(zb));
            if (ncl > 1) {
              // my rows [vlo, vhi) of channel group kg go to the same place in every peer's ring as one bulk copy through
              // the async proxy, which completes its bytes on the peer's "z full" barrier: data and signal travel together
              // (remote generic-proxy stores followed by a remote arrive were NOT ordered before the peer's UMMA reads)
              const uint32_t bytes = live ? (uint32_t)(vhi - vlo) * 16u : 0u;
              const uint32_t off = (uint32_t)(OFF_Z + zb * Z_BUF_BYTES + (kg * ZR + vlo) * 16);
              const uint32_t boff = BAR_ZFULL(zb) - s_base;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                if (j < ncl - 1) {
                  if (bytes) {
                    mbar_arrive_expect_tx_cluster(peer_base[j] + boff, bytes);
                    dsmem_copy(peer_base[j] + off, s_base + off, bytes, peer_base[j] + boff);
                  } else {
                    mbar_arrive_cluster(peer_base[j] + boff);
                  }
                }
            }
            mbar_arrive(BAR_XEMPTY(xb));
          }
          if (++xb == NX) { xb = 0; xph ^= 1; }
          if (++zb == NZ) { zb = 0; zph ^= 1; }
        }
        b = b2; t0 = t02; nt = nt2; T = T2;
      }
    } else if (warp >= 1 && warp < 4 && lane == 0) {
      x_loop(warp, warp + 1, false);          // un-activated layer: channel group `warp` of every x tile
    }
  } else if (warp < NW_ACT + 4) {
    reg_dec<32>();
    if (warp == NW_ACT) {
      // ===================== x producer (TMA) =====================
      if (lane == 0) x_loop(0, ACT ? 4 : 1, true);
    } else if (warp == NW_ACT + 1) {
      // ===================== weight producer (bulk copies) =====================
      // ... and the residual / running-sum rows of the tile: the layer's `+ x` (models.py:72) and the sum over the three
      // blocks (:239-245) are accumulated by the tensor core as D += R x I.  R = rows [t0, t0+256) of 32 channels,
      // staged by TMA in the UMMA A layout; I = identity weight tile streamed through the weight ring like a tap.
      if (lane == 0) {
        int stage = 0, phase = 0, rs = 0, rph = 0;
        // narrow layers: a tile's worth of slack on every ring, and whatever this thread executes while waiting comes out of
        // the activation warps' issue slots — sleep for real (nanosleep <= 500 returns at once on this part)
        const uint32_t nap = a.Cin <= 96 ? 800u : 300u;
        TileCursor cur{prefix};
        for (int w = w0; w < total_tiles; w += wstep) {
          int b, t0, nt;
          cur.locate(w, n_tiles, b, t0, nt);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wt) + (size_t)nt * NCH * a.K * tile_bytes;
          for (int c = 0; c < NCH; ++c) {
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              const uint32_t bytes = BVG_DBGBIT(a, 2) ? 16u : (uint32_t)(taps * tile_bytes);
              mbar_wait_relaxed(BAR_WEMPTY(stage), phase ^ 1, nap);
              mbar_expect_tx(BAR_WFULL(stage), bytes);
              bulk_load(s_base + OFF_W + stage * W_STAGE_BYTES, src, bytes, BAR_WFULL(stage));
              src += bytes;
              if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
            }
            // residual chunk c of every stream rides behind conv chunk c (nchr <= NCH): the loads are spread over the
            // tile, so the two-slot ring always has a chunk time to cover the TMA latency
            if (RM && c < a.nchr)
              for (int st = 0; st < nstreams_r; ++st) {
                const CUtensorMap* rm = (st == 0 && a.rmma_r) ? &tmr : &tmq;
                // identity tile of (column tile nt, input chunk nt * nchr + c)
                const uint8_t* isrc = reinterpret_cast<const uint8_t*>(a.idw) + ((size_t)nt * NCH + nt * a.nchr + c) * tile_bytes;
                mbar_wait_relaxed(BAR_WEMPTY(stage), phase ^ 1, nap);
                mbar_expect_tx(BAR_WFULL(stage), (uint32_t)tile_bytes);
                bulk_load(s_base + OFF_W + stage * W_STAGE_BYTES, isrc, (uint32_t)tile_bytes, BAR_WFULL(stage));
                if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
                mbar_wait_relaxed(BAR_REMPTY(rs), rph ^ 1, nap);
                mbar_expect_tx(BAR_RFULL(rs), R_SLOT_BYTES);
                const uint32_t dst = s_base + OFF_R + rs * R_SLOT_BYTES;
                const int g0 = nt * (n_tile >> 3) + c * 4;
#pragma unroll
                for (int kg = 0; kg < 4; ++kg)
#pragma unroll
                  for (int h = 0; h < 2; ++h)
                    tma_load_4d(dst + kg * (M_TILE * 16) + h * (128 * 16), rm, 0, t0 + h * 128, g0 + kg, b, BAR_RFULL(rs));
                if (++rs == R_RING) { rs = 0; rph ^= 1; }
              }
          }
        }
      }
    } else {
      // ===================== MMA issuers =====================
      // Two warps, one per M block (accumulator) of the tile: an issuer is a single thread's worth of dependent
      // uniform-datapath code — ~9 instructions per tcgen05.mma at ~14 cycles each (ncu, stage 5, k = 11: the narrow
      // many-tap layers were bound by exactly this loop, their activation warps spinning on full z slots) — and the two
      // accumulators are independent, so each warp issues its own block's MMAs with its TMEM address and instruction
      // descriptor loop-invariant.  One issuer per accumulator keeps the summation order, hence the result, fixed.
      // 32 registers: everything loop-carried is kept to a minimum (one operand ring index instead of separate x / z
      // cursors, descriptors rebuilt from 32-bit address units) because a spilled value costs an L2 round trip here.
      {
        const uint32_t mbk = (uint32_t)(warp - (NW_ACT + 2));   // my M block: rows [128 mbk, 128 mbk + 128)
        const uint32_t leader = elect_one();      // the whole warp walks the loop; one lane issues
        const uint32_t idesc = make_idesc_bf16(128, n_tile);
        const uint32_t lboA = (ACT ? ZR : XRA) * 16;
        const uint32_t lboB = (uint32_t)n_tile * 16;
        // descriptor = constant high part | (smem address >> 4); taps / M blocks / K steps only
        // move the 14-bit address field (rows are 16 B apart: +1 unit = +1 time sample)
        const u64 hiA = make_sdesc(0, lboA, 128), hiB = make_sdesc(0, lboB, 128);
        const uint32_t ksA = 2 * lboA / 16, ksB = 2 * lboB / 16, tileU = (uint32_t)tile_bytes / 16;
        const int ND = ACT ? NZ : NX;                       // depth of the A-operand ring (z tiles, or raw x tiles)
        // Under a cluster launch the shared::cta window address of CTA rank r carries r in its upper bits: only the low
        // 18 bits (offset inside this SM's shared memory) belong in a UMMA descriptor's 14-bit address field — unmasked,
        // the rank bit landed in the leading-byte-offset field of every CTA but rank 0.
        const uint32_t s_loc = s_base & 0x3FFFFu;
        const uint32_t ring0 = (ACT ? (s_loc + OFF_Z) : (s_loc + OFF_X + (X_LEAD - a.lead) * 16)) >> 4;
        const uint32_t ringU = (ACT ? Z_BUF_BYTES : X_BUF_BYTES) >> 4;
        const uint32_t barF = ACT ? BAR_ZFULL(0) : BAR_XFULL(0), barE = ACT ? BAR_ZEMPTY(0) : BAR_XEMPTY(0);
        const bool lazy = ACT && a.Cin <= 96 && !BVG_DBGBIT(a, 4);
        const u64 hiR = make_sdesc(0, M_TILE * 16, 128);    // residual slot: 256 rows per 8-channel group
        int stage = 0, phase = 0, rb = 0, rph = 0, it = 0, rr = 0, rrph = 0;
        for (int w = tile_w0(a); w < total_tiles; w += tile_wstep(a), ++it) {
          const int as = (nacc == 2) ? (it & 1) : 0;
          mbar_wait(BAR_ACCEMPTY(as), ((((nacc == 2) ? (it >> 1) : it) & 1) ^ 1));   // epilogue has drained this stage
          tc_fence_after();
          const uint32_t tm = tmem + (uint32_t)(as * 2 * n_tile) + mbk * (uint32_t)n_tile;
          uint32_t accflag = 0;
          if (mbk == 0) TC_TRACE(a, 3000 + it, 3);
          for (int c = 0; c < NCH; ++c) {
            // narrow layers are bound by the activation warps' issue slots: the issuer sleeps between polls there
            if (ncl > 1) mbar_wait_acq_cluster(barF + 8 * rb, rph);     // arrivals (and bulk-copy bytes) come from peer CTAs too
            else if (lazy) mbar_wait_relaxed(barF + 8 * rb, rph, 800);
            else mbar_wait(barF + 8 * rb, rph);
            tc_fence_after();
            if (mbk == 0) TC_TRACE(a, 1100 + it * NCH + c, 2);
            const uint32_t aU = ring0 + (uint32_t)rb * ringU;
            for (int s = 0; s < spc; ++s) {
              const int taps = min(tps, a.K - s * tps);
              mbar_wait(BAR_WFULL(stage), phase);
              tc_fence_after();
              const uint32_t wU = (s_loc + OFF_W + stage * W_STAGE_BYTES) >> 4;
              for (int tj = 0; tj < taps; ++tj) {
                const uint32_t a0 = aU + (uint32_t)((s * tps + tj) * a.dil) + mbk * 128u;
                const uint32_t b0 = wU + (uint32_t)tj * tileU;
                umma_bf16_e(leader, tm, hiA | a0, hiB | b0, idesc, accflag);
                if (!BVG_DBGBIT(a, 1)) umma_bf16_e(leader, tm, hiA | (a0 + ksA), hiB | (b0 + ksB), idesc, 1u);
                accflag = 1u;
              }
              umma_commit_e(leader, BAR_WEMPTY(stage));          // weight stage free once these MMAs retire
              if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
            }
            if (ncl > 1 && !BVG_DBGBIT(a, 4096)) umma_commit_mc_e(leader, barE + 8 * rb, (1u << ncl) - 1u);   // frees the slot in every CTA's ring
            else umma_commit_e(leader, barE + 8 * rb);
            if (++rb == ND) { rb = 0; rph ^= 1; }
            // + residual (+ running sum) chunk c: D += R x I, two K steps of 16 channels
            if (RM && c < a.nchr)
              for (int st = 0; st < nstreams_r; ++st) {
                mbar_wait(BAR_RFULL(rr), rrph);
                mbar_wait(BAR_WFULL(stage), phase);
                tc_fence_after();
                const uint32_t r0 = (s_loc + OFF_R + rr * R_SLOT_BYTES) >> 4;
                const uint32_t b0 = (s_loc + OFF_W + stage * W_STAGE_BYTES) >> 4;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  umma_bf16_e(leader, tm, hiR | (r0 + ks * (2 * M_TILE) + mbk * 128u), hiB | (b0 + ks * ksB), idesc, 1u);
                umma_commit_e(leader, BAR_WEMPTY(stage));
                if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
                umma_commit_e(leader, BAR_REMPTY(rr));
                if (++rr == R_RING) { rr = 0; rrph ^= 1; }
              }
          }
          umma_commit_e(leader, BAR_ACCFULL(as));
          if (mbk == 0) TC_TRACE(a, 3100 + it, 4);
        }
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> (+bias, +resid, +sum, /div) -> bf16 -> HBM ====
    reg_dec<64>();
    // (instantiations whose residual goes through the tensor core, RM, only ever see a plain conv here)
    if constexpr (EPI == 0)
      epilogue_pipe<false, RM>(a, bias_s, smem + off_r(NX, NZ, W_STAGES), prefix, BAR_ACCFULL(0), BAR_ACCEMPTY(0), tmem, nacc, total_tiles, warp & 3, lane,
                           threadIdx.x - (NW_ACT + 4) * 32);
    else if constexpr (EPI == 1)
      epilogue_pipe<true, RM>(a, bias_s, smem + off_r(NX, NZ, W_STAGES), prefix, BAR_ACCFULL(0), BAR_ACCEMPTY(0), tmem, nacc, total_tiles, warp & 3, lane,
                          threadIdx.x - (NW_ACT + 4) * 32);
    else
      epilogue_up(a, bias_s, prefix, BAR_ACCFULL(0), BAR_ACCEMPTY(0), tmem, nacc, total_tiles, extra, warp & 3, lane,
                  threadIdx.x - (NW_ACT + 4) * 32);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) TC_TRACE(a, 2, 10);
  if (ncl > 1) cluster_sync_all();     // nobody leaves while a peer may still store into its rings or arrive on its barriers
  if (warp == NW_ACT + 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace tc
}  // namespace bvg
