// ecapa.cu — the ECAPA-TDNN speaker encoder of BigVGAN.forward (indextts/BigVGAN/models.py:204 ->
// ECAPA_TDNN.py:543-581) as hand-written fp32 CUDA: mel_ref [B, Tm, 100] -> embedding [B, 512].
//
// The encoder is ~3 GFLOP in ~45 small layers at T = 300: through PyTorch / cuDNN it is 60+ launch-bound kernels
// (2.9 ms eager, 0.7 ms replayed from a CUDA graph) — 3 % of a 16 x 10 s decode but a fifth of the 6.7 s latency case —
// and with cuDNN's default it computes in TF32.  Here: five kernels, fp32 FFMA (the embedding feeds the fp32 exactness
// path), every layer's ReLU / BatchNorm(eval) / bias folded into the producing kernel, the Res2Net "cat" and the MFA
// "cat" written in place, the attentive-statistics context concat folded into a per-utterance bias, and the whole
// sequence replayed from a CUDA graph per (B, Tm).
//   TDNNBlock   conv (reflect "same" padding, nnet/CNN.py:458-487) -> ReLU -> BatchNorm           ECAPA_TDNN.py:79-128
//   Res2Net     8 chunks of 64 channels, y_i = TDNN_i(x_i + y_{i-1})                             :131-191
//   SEBlock     gate = sigmoid(W2 relu(W1 mean_t(x)))                                            :194-242
//   ASP         attentive statistics pooling with global context                                 :245-338
// lens = None only (the inference call, infer.py:886-888, passes none); the PyTorch module keeps the lens path.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace bvg {
namespace ec {

// --------------------------------------------------------------------------------------- conv1d as a tiled GEMM
struct ConvArgsE {
  const float* in;      // channel-major [B][*][T] (pointer already at the first input channel), or time-major mel
  const float* in2;     // optional second input added element-wise (Res2Net x_i + y_{i-1}), same geometry, or null
  long long in_bstride, in2_bstride;   // elements between utterances
  int in_tm;            // 1: `in` is time-major [B][T][Cin] (the mel)
  int in_tanh;          // tanh on the input (ASP: conv(tanh(tdnn(ctx))))
  const float* w;       // [Cin][K][Cout]
  const float* bias;    // [Cout]
  const float* eb;      // optional per-(b, co) extra bias [B][eb_stride], or null
  int eb_stride;
  const float* bn_sc;   // BatchNorm(eval) scale / shift applied AFTER the ReLU (TDNNBlock order), or null
  const float* bn_sh;
  int relu;
  float* out;           // channel-major, pointer at the first output channel
  long long out_bstride;
  int Cin, Cout, K, dil, T;
};



__device__ __forceinline__ int reflect(int t, int T) { return t < 0 ? -t : (t >= T ? 2 * (T - 1) - t : t); }

__device__ __forceinline__ void ec_epilogue(const ConvArgsE& a, int b, int co, int t, float v) {
  float bi = a.bias ? __ldg(a.bias + co) : 0.f;
  if (a.eb) bi += __ldg(a.eb + (size_t)b * a.eb_stride + co);
  v += bi;
  if (a.relu) v = fmaxf(v, 0.f);
  if (a.bn_sc) v = v * __ldg(a.bn_sc + co) + __ldg(a.bn_sh + co);
  a.out[(size_t)b * a.out_bstride + (size_t)co * a.T + t] = v;
}

// CTA tile = TCOv output channels x TTv time steps, NTHR threads, each RCO channels x RT time steps; the input channels
// of a CTA ([ci_lo, ci_hi): split-K slice blockIdx.z % S) are walked in chunks of CK with the next chunk's weights and
// inputs prefetched into registers while the current one is multiplied out of shared memory.  Two shapes are built:
//   <32, 32, 128, 4, 2>  register-tiled (8 FMA per 3 shared loads) for the one FLOP-heavy layer (MFA, 1536 -> 1536);
//   <16, 16, 256, 1, 1>  one output per thread for everything else: at T = 300 those layers are a few MFLOP each and
//                        bound by the LENGTH of a thread's dependent chain and by how many warps an SM has to switch
//                        between (the register-tiled shape left one warp per scheduler on 20-80 SMs: 13-38 us a layer).
// S > 1: partial sums go to `part` [S][B][Cout][T]; k_ec_reduce adds them in a fixed order (deterministic) and applies the epilogue.
template <int K, int TCOv, int TTv, int NTHR, int RCO, int RT>
__global__ void __launch_bounds__(NTHR) k_ec_conv(const ConvArgsE a, int S, int cper, float* __restrict__ part) {
  constexpr int CK = (K == 1) ? 32 : 16;
  constexpr int XWMAX = TTv + ((K == 1) ? 0 : (K == 3 ? 8 : 4));         // halo <= 4 (k = 3, dilation <= 4) / 2 (k = 5)
  constexpr int NW = (CK * K * TCOv / 4 + NTHR - 1) / NTHR;              // float4 weight loads per thread and chunk
  constexpr int NX = (CK * XWMAX + NTHR - 1) / NTHR;                     // input loads per thread and chunk
  constexpr int TPR = TTv / RT;                                          // threads along time
  static_assert(TCOv % 4 == 0 && (TCOv / RCO) * TPR == NTHR, "thread tiling");
  __shared__ __align__(16) float ws[CK * K * TCOv];
  __shared__ float xs[CK * XWMAX];
  const int halo = a.dil * (K - 1) / 2;
  const int XW = TTv + 2 * halo;
  const int b = blockIdx.z / S, sl = blockIdx.z - b * S;
  const int co0 = blockIdx.y * TCOv, t0 = blockIdx.x * TTv;
  const int ci_lo = sl * cper, ci_hi = min(a.Cin, ci_lo + cper);
  const int tid = threadIdx.x, tc = tid % TPR, cc = tid / TPR;
  const float* inb = a.in + (size_t)b * a.in_bstride;
  const float* in2b = a.in2 ? a.in2 + (size_t)b * a.in2_bstride : nullptr;
  float4 wr[NW];
  float xr[NX];
  auto fetch_w = [&](int ci0) {
    const int nci = min(CK, ci_hi - ci0);
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int e = (i * NTHR + tid) * 4;                // element in [CK][K][TCOv]
      const int co = e % TCOv, r = e / TCOv;             // r = ci * K + j
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nci * K) {
        const float* src = a.w + ((size_t)ci0 * K + r) * a.Cout + co0 + co;
        if (co0 + co + 3 < a.Cout && ((a.Cout & 3) == 0)) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (co0 + co + 0 < a.Cout) v.x = __ldg(src + 0);
          if (co0 + co + 1 < a.Cout) v.y = __ldg(src + 1);
          if (co0 + co + 2 < a.Cout) v.z = __ldg(src + 2);
          if (co0 + co + 3 < a.Cout) v.w = __ldg(src + 3);
        }
      }
      wr[i] = v;
    }
  };
  auto fetch_x = [&](int ci0) {
    const int nci = min(CK, ci_hi - ci0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int e = i * NTHR + tid;
      const int ci = e / XW, tt = e - ci * XW;
      float v = 0.f;
      if (ci < nci) {
        const int t = reflect(t0 - halo + tt, a.T);
        if (t >= 0 && t < a.T) {
          if (a.in_tm) v = __ldg(inb + (size_t)t * a.Cin + ci0 + ci);
          else {
            v = __ldg(inb + (size_t)(ci0 + ci) * a.T + t);
            if (in2b) v += __ldg(in2b + (size_t)(ci0 + ci) * a.T + t);
          }
          if (a.in_tanh) v = tanhf(v);
        }
      }
      xr[i] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int e = (i * NTHR + tid) * 4;
      if (e < CK * K * TCOv) *reinterpret_cast<float4*>(ws + e) = wr[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int e = i * NTHR + tid;
      if (e < CK * XW) xs[e] = xr[i];
    }
  };
  float acc[RCO][RT] = {};
  // programmatic dependent launch: this CTA may have started while the producing layer is still running — the weights
  // do not depend on it, so the first weight tile is already on its way when the wait returns
  pdl_launch_dependents();
  fetch_w(ci_lo);
  pdl_wait();
  fetch_x(ci_lo);
  for (int ci0 = ci_lo; ci0 < ci_hi; ci0 += CK) {
    __syncthreads();                       // everyone is done with the previous chunk's tiles
    stash();
    __syncthreads();
    if (ci0 + CK < ci_hi) { fetch_w(ci0 + CK); fetch_x(ci0 + CK); }
    const int nci = min(CK, ci_hi - ci0);
    for (int ci = 0; ci < nci; ++ci)
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float w[RCO], x[RT];
        if constexpr (RCO == 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + (ci * K + j) * TCOv + 4 * cc);
          w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
        } else {
#pragma unroll
          for (int r = 0; r < RCO; ++r) w[r] = ws[(ci * K + j) * TCOv + RCO * cc + r];
        }
#pragma unroll
        for (int e = 0; e < RT; ++e) x[e] = xs[ci * XW + RT * tc + e + j * a.dil];
#pragma unroll
        for (int r = 0; r < RCO; ++r)
#pragma unroll
          for (int e = 0; e < RT; ++e) acc[r][e] = fmaf(w[r], x[e], acc[r][e]);
      }
  }
#pragma unroll
  for (int r = 0; r < RCO; ++r) {
    const int co = co0 + RCO * cc + r;
    if (co >= a.Cout) continue;
#pragma unroll
    for (int e = 0; e < RT; ++e) {
      const int t = t0 + RT * tc + e;
      if (t >= a.T) continue;
      if (S == 1) ec_epilogue(a, b, co, t, acc[r][e]);
      else part[(((size_t)sl * (gridDim.z / S) + b) * a.Cout + co) * a.T + t] = acc[r][e];
    }
  }
}

__global__ void k_ec_reduce(const ConvArgsE a, int S, int B, const float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = (size_t)B * a.Cout * a.T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part[(size_t)s * n + i];
    const int t = (int)(i % a.T);
    const size_t r = i / a.T;
    ec_epilogue(a, (int)(r / a.Cout), (int)(r % a.Cout), t, v);
  }
}

// --------------------------------------------------------------------------------------- squeeze-excitation gate
// out[b][c][t] = gate[b][c] * x[b][c][t] + r[b][c][t]   (SERes2NetBlock: se_block(x) + residual)
__global__ void k_ec_scale_res(const float* __restrict__ x, long long xbs, const float* __restrict__ gate,
                               const float* __restrict__ r, long long rbs, float* __restrict__ out, long long obs, int C,
                               int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * T; i += gridDim.x * blockDim.x) {
    const int c = i / T;
    out[(size_t)b * obs + i] = gate[(size_t)b * C + c] * x[(size_t)b * xbs + i] + r[(size_t)b * rbs + i];
  }
}

// --------------------------------------------------------------------------------------- (attentive) statistics
// One warp per (b, c) row.  att == null: uniform weights 1/T;  else w = softmax_t(att[b][c][:]).
// mean = sum w x,  std = sqrt(max(sum w (x - mean)^2, 1e-12))      (ECAPA_TDNN.py:283-296)
__global__ void k_ec_stats(const float* __restrict__ x, const float* __restrict__ att, int C, int T,
                           float* __restrict__ mean_out, float* __restrict__ std_out, int out_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (row >= C) return;
  const float* xr = x + ((size_t)b * C + row) * T;
  const float* ar = att ? att + ((size_t)b * C + row) * T : nullptr;
  float mx = -INFINITY;
  if (ar) {
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, ar[t]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  float sw = 0.f, swx = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float w = ar ? expf(ar[t] - mx) : 1.f;
    sw += w;
    swx = fmaf(w, xr[t], swx);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    swx += __shfl_xor_sync(0xffffffffu, swx, o);
  }
  const float mean = swx / sw;
  float sv = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float w = ar ? expf(ar[t] - mx) : 1.f;
    const float d = xr[t] - mean;
    sv = fmaf(w, d * d, sv);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
  if (lane == 0) {
    mean_out[(size_t)b * out_stride + row] = mean;
    std_out[(size_t)b * out_stride + row] = sqrtf(fmaxf(sv / sw, 1e-12f));
  }
}

// y[b][o] = act(bias[o] + sum_i W[o * wstride + i] * (x[b][i] * sc[i] + sh[i]));  one warp per output
// act: 0 none, 1 relu, 2 sigmoid
__global__ void k_ec_gemv(const float* __restrict__ W, int wstride, const float* __restrict__ bias,
                          const float* __restrict__ x, int xstride, const float* __restrict__ sc,
                          const float* __restrict__ sh, int I, int O, float* __restrict__ y, int ystride, int act) {
  pdl_launch_dependents();
  pdl_wait();
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  if (o >= O) return;
  float v = 0.f;
  for (int i = lane; i < I; i += 32) {
    float xv = x[(size_t)b * xstride + i];
    if (sc) xv = xv * sc[i] + sh[i];
    v = fmaf(__ldg(W + (size_t)o * wstride + i), xv, v);
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) {
    v += bias ? bias[o] : 0.f;
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 1.f / (1.f + expf(-v));
    y[(size_t)b * ystride + o] = v;
  }
}

// mean over time of every (b, c) row: one warp per row (SEBlock's squeeze)
__global__ void k_ec_rowmean(const float* __restrict__ x, long long bstride, int C, int T, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  if (row >= C) return;
  const float* xr = x + (size_t)b * bstride + (size_t)row * T;
  float v = 0.f;
  for (int t = lane; t < T; t += 32) v += xr[t];
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[(size_t)b * C + row] = v / (float)T;
}

}  // namespace ec
}  // namespace bvg

using namespace bvg;
using namespace bvg::ec;

// =============================================================================== host side
struct bvg_ecapa {
  bvg_ecapa_desc d;
  int device = 0;
  // workspace for (capB, capT)
  int capB = 0, capT = 0;
  float *x0 = nullptr, *y = nullptr, *z = nullptr, *t2 = nullptr, *f = nullptr, *m = nullptr, *a1 = nullptr, *att = nullptr;
  float *gate = nullptr, *stat = nullptr /*[B][3072] mean|std*/, *eb = nullptr /*[B][att]*/, *pooled = nullptr /*[B][3072]*/;
  float *semean = nullptr /*[B][C]*/, *sehid = nullptr /*[B][se]*/, *part = nullptr /* split-K partial sums */;
  size_t part_elems = 0;
  void* mel_stage = nullptr;       // fixed-address copy of the caller's mel (graph replay)
  size_t mel_bytes = 0;
  float* emb_stage = nullptr;
  // one captured graph per (B, Tm, mel dtype)
  struct G { int B, T, dt; cudaGraphExec_t exec; };
  std::vector<G> graphs;
  cudaStream_t cap = nullptr;
  int launches = 0;
};

static int ec_conv(bvg_ecapa* e, const bvg_ecapa_tdnn& L, const float* in, long long in_bs, const float* in2, long long in2_bs,
                   int in_tm, int in_tanh, const float* eb, int eb_stride, float* out, long long out_bs, int B, int T,
                   cudaStream_t st) {
  ConvArgsE a{};
  a.in = in; a.in2 = in2; a.in_bstride = in_bs; a.in2_bstride = in2_bs; a.in_tm = in_tm; a.in_tanh = in_tanh;
  a.w = L.w; a.bias = L.bias; a.eb = eb; a.eb_stride = eb_stride; a.bn_sc = L.bn_scale; a.bn_sh = L.bn_shift;
  a.relu = L.relu; a.out = out; a.out_bstride = out_bs;
  a.Cin = L.cin; a.Cout = L.cout; a.K = L.k; a.dil = L.dil; a.T = T;
  const int halo = L.dil * (L.k - 1) / 2;
  if (halo >= T) return fail(BVG_ERR_ARG, "speaker encoder: prompt of %d mel frames is shorter than a conv halo (%d)", T, halo);
  if ((L.k == 3 && halo > 4) || (L.k == 5 && halo > 2) || (L.k != 1 && L.k != 3 && L.k != 5))
    return fail(BVG_ERR_UNSUPPORTED, "speaker encoder: conv k=%d dilation=%d outside the built kernel shapes", L.k, L.dil);
  // tile shape: register-tiled for the FLOP-heavy layer, one output per thread otherwise; split-K so that a layer
  // keeps every SM busy (>= ~600 CTAs of the small shape) without starving a CTA of work (>= 64 input channels)
  const bool big = (double)L.cin * L.k * L.cout >= 1.0e6;
  const int tco = big ? 32 : 16, tt = big ? 32 : 16;
  const int ck = (L.k == 1) ? 32 : 16;
  const int tiles = ceil_div(T, tt) * ceil_div(L.cout, tco) * B;
  int S = 1;
  while (S < 8 && tiles * S < (big ? 296 : 592) && L.cin / (2 * S) >= 64) S *= 2;
  const int cper = ceil_div(ceil_div(L.cin, S), ck) * ck;
  S = ceil_div(L.cin, cper);
  if (S > 1 && (size_t)S * B * L.cout * T > e->part_elems) return fail(BVG_ERR_STATE, "speaker encoder: split-K scratch too small");
  dim3 grid(ceil_div(T, tt), ceil_div(L.cout, tco), B * S);
#define EC_LAUNCH(KK)                                                                                              \
  do {                                                                                                             \
    if (big) BVG_CUDA(launch_k(k_ec_conv<KK, 32, 32, 128, 4, 2>, grid, dim3(128), 0, st, true, a, S, cper, e->part)); \
    else BVG_CUDA(launch_k(k_ec_conv<KK, 16, 16, 256, 1, 1>, grid, dim3(256), 0, st, true, a, S, cper, e->part));    \
  } while (0)
  if (L.k == 1) EC_LAUNCH(1);
  else if (L.k == 3) EC_LAUNCH(3);
  else EC_LAUNCH(5);
#undef EC_LAUNCH
  if (S > 1) {
    const size_t n = (size_t)B * L.cout * T;
    BVG_CUDA(launch_k(k_ec_reduce, dim3((unsigned)std::min<size_t>((n + 255) / 256, 592)), dim3(256), 0, st, true, a, S, B, (const float*)e->part));
  }
  return 0;
}

static int ec_run(bvg_ecapa* e, const float* mel, int B, int T, float* emb, cudaStream_t st) {
  const bvg_ecapa_desc& d = e->d;
  const int C = d.channels, MF = d.mfa_channels, S = d.scale, W = C / S, AT = d.att_channels;
  const long long cs = (long long)C * T, ms = (long long)MF * T;
  int rc;
  // blocks[0]: TDNN on the time-major mel
  if ((rc = ec_conv(e, d.block0, mel, (long long)T * d.in_channels, nullptr, 0, 1, 0, nullptr, 0, e->x0, cs, B, T, st))) return rc;
  const float* xin = e->x0;
  long long xin_bs = cs;
  for (int i = 0; i < 3; ++i) {
    const bvg_ecapa_block& bl = d.blocks[i];
    if ((rc = ec_conv(e, bl.tdnn1, xin, xin_bs, nullptr, 0, 0, 0, nullptr, 0, e->y, cs, B, T, st))) return rc;
    // Res2Net: chunk 0 passes through, chunk 1 = TDNN(y_1), chunk i = TDNN(y_i + z_{i-1})
    BVG_CUDA(cudaMemcpy2DAsync(e->z, (size_t)cs * 4, e->y, (size_t)cs * 4, (size_t)W * T * 4, B, cudaMemcpyDeviceToDevice, st));
    for (int k = 1; k < S; ++k)
      if ((rc = ec_conv(e, bl.res2[k - 1], e->y + (size_t)k * W * T, cs, k > 1 ? e->z + (size_t)(k - 1) * W * T : nullptr, cs, 0, 0,
                        nullptr, 0, e->z + (size_t)k * W * T, cs, B, T, st)))
        return rc;
    if ((rc = ec_conv(e, bl.tdnn2, e->z, cs, nullptr, 0, 0, 0, nullptr, 0, e->t2, cs, B, T, st))) return rc;
    // squeeze-excitation gate: mean over time, two small GEMVs (one warp per output)
    const float* nof = nullptr;
    BVG_CUDA(launch_k(k_ec_rowmean, dim3(ceil_div(C * 32, 256), B), dim3(256), 0, st, true, (const float*)e->t2, cs, C, T, e->semean));
    BVG_CUDA(launch_k(k_ec_gemv, dim3(ceil_div(d.se_channels * 32, 256), B), dim3(256), 0, st, true, bl.se_w1, C, bl.se_b1,
                      (const float*)e->semean, C, nof, nof, C, (int)d.se_channels, e->sehid, (int)d.se_channels, 1));
    BVG_CUDA(launch_k(k_ec_gemv, dim3(ceil_div(C * 32, 256), B), dim3(256), 0, st, true, bl.se_w2, (int)d.se_channels, bl.se_b2,
                      (const float*)e->sehid, (int)d.se_channels, nof, nof, (int)d.se_channels, C, e->gate, C, 2));
    float* fo = e->f + (size_t)i * C * T;           // the block's output lands in its slice of the MFA concat
    BVG_CUDA(launch_k(k_ec_scale_res, dim3(ceil_div(C * T, 256 * 4), B), dim3(256), 0, st, true, (const float*)e->t2, cs,
                      (const float*)e->gate, xin, xin_bs, fo, ms, C, T));
    xin = fo;
    xin_bs = ms;
  }
  if ((rc = ec_conv(e, d.mfa, e->f, ms, nullptr, 0, 0, 0, nullptr, 0, e->m, ms, B, T, st))) return rc;
  // attentive statistics pooling: global context (mean, std) folded into a per-utterance bias of the 1x1 TDNN
  k_ec_stats<<<dim3(ceil_div(MF * 32, 256), B), 256, 0, st>>>(e->m, nullptr, MF, T, e->stat, e->stat + MF, 2 * MF);
  BVG_CUDA(cudaGetLastError());
  k_ec_gemv<<<dim3(ceil_div(AT * 32, 256), B), 256, 0, st>>>(d.asp_ctx_w, 2 * MF, nullptr, e->stat, 2 * MF, nullptr, nullptr,
                                                            2 * MF, AT, e->eb, AT, 0);
  BVG_CUDA(cudaGetLastError());
  if ((rc = ec_conv(e, d.asp_tdnn, e->m, ms, nullptr, 0, 0, 0, e->eb, AT, e->a1, (long long)AT * T, B, T, st))) return rc;
  if ((rc = ec_conv(e, d.asp_conv, e->a1, (long long)AT * T, nullptr, 0, 0, 1, nullptr, 0, e->att, ms, B, T, st))) return rc;
  k_ec_stats<<<dim3(ceil_div(MF * 32, 256), B), 256, 0, st>>>(e->m, e->att, MF, T, e->pooled, e->pooled + MF, 2 * MF);
  BVG_CUDA(cudaGetLastError());
  k_ec_gemv<<<dim3(ceil_div(d.emb_dim * 32, 256), B), 256, 0, st>>>(d.fc_w, 2 * MF, d.fc_b, e->pooled, 2 * MF, d.asp_bn_scale,
                                                                   d.asp_bn_shift, 2 * MF, d.emb_dim, emb, d.emb_dim, 0);
  BVG_CUDA(cudaGetLastError());
  return 0;
}

__global__ void k_ec_mel_to_f32(const void* __restrict__ src, int dtype, float* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ld_dyn(src, i, dtype);
}

extern "C" {

int bvg_ecapa_create(const bvg_ecapa_desc* desc, int device, bvg_ecapa** out) {
  BVG_REQUIRE(desc && out, "bvg_ecapa_create: null argument");
  BVG_REQUIRE(desc->channels % desc->scale == 0 && desc->mfa_channels == 3 * desc->channels, "bvg_ecapa_create: unsupported geometry");
  BVG_REQUIRE(desc->channels <= 1024 && desc->se_channels <= 512, "bvg_ecapa_create: channel counts out of range");
  BVG_CUDA(cudaSetDevice(device));
  bvg_ecapa* e = new bvg_ecapa();
  e->d = *desc;
  e->device = device;
  *out = e;
  return 0;
}

static void ec_free_ws(bvg_ecapa* e) {
  for (float** p : {&e->x0, &e->y, &e->z, &e->t2, &e->f, &e->m, &e->a1, &e->att, &e->gate, &e->stat, &e->eb, &e->pooled, &e->emb_stage, &e->semean, &e->sehid, &e->part}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (e->mel_stage) cudaFree(e->mel_stage);
  e->mel_stage = nullptr;
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

int bvg_ecapa_destroy(bvg_ecapa* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  ec_free_ws(e);
  if (e->cap) cudaStreamDestroy(e->cap);
  delete e;
  return 0;
}

int bvg_ecapa_forward(bvg_ecapa* e, const void* mel, int mel_dtype, int B, int Tm, float* emb_out, void* stream) {
  BVG_REQUIRE(e && mel && emb_out, "bvg_ecapa_forward: null argument");
  BVG_REQUIRE(B >= 1 && Tm >= 1 && B <= 65535, "bvg_ecapa_forward: B=%d Tm=%d", B, Tm);
  BVG_REQUIRE(mel_dtype >= BVG_F32 && mel_dtype <= BVG_F16, "bvg_ecapa_forward: bad mel dtype");
  BVG_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const bvg_ecapa_desc& d = e->d;
  if (B > e->capB || Tm > e->capT) {
    BVG_CUDA(cudaDeviceSynchronize());
    ec_free_ws(e);
    const int cb = std::max(B, e->capB), ct = std::max(Tm, e->capT);
    const size_t c = (size_t)cb * d.channels * ct, m = (size_t)cb * d.mfa_channels * ct;
    for (float** p : {&e->x0, &e->y, &e->z, &e->t2}) BVG_CUDA(cudaMalloc((void**)p, c * 4));
    for (float** p : {&e->f, &e->m, &e->att}) BVG_CUDA(cudaMalloc((void**)p, m * 4));
    BVG_CUDA(cudaMalloc((void**)&e->a1, (size_t)cb * d.att_channels * ct * 4));
    BVG_CUDA(cudaMalloc((void**)&e->gate, (size_t)cb * d.channels * 4));
    BVG_CUDA(cudaMalloc((void**)&e->stat, (size_t)cb * 2 * d.mfa_channels * 4));
    BVG_CUDA(cudaMalloc((void**)&e->pooled, (size_t)cb * 2 * d.mfa_channels * 4));
    BVG_CUDA(cudaMalloc((void**)&e->eb, (size_t)cb * d.att_channels * 4));
    BVG_CUDA(cudaMalloc((void**)&e->emb_stage, (size_t)cb * d.emb_dim * 4));
    BVG_CUDA(cudaMalloc((void**)&e->semean, (size_t)cb * d.channels * 4));
    BVG_CUDA(cudaMalloc((void**)&e->sehid, (size_t)cb * d.se_channels * 4));
    e->part_elems = (size_t)8 * m;                       // up to 8 split-K slices of the largest layer output
    BVG_CUDA(cudaMalloc((void**)&e->part, e->part_elems * 4));
    e->mel_bytes = (size_t)cb * ct * d.in_channels * 4;
    BVG_CUDA(cudaMalloc(&e->mel_stage, e->mel_bytes));
    e->capB = cb; e->capT = ct;
  }
  // the caller's mel (any float dtype) -> fp32 at a fixed address, the graph's input
  const size_t n = (size_t)B * Tm * d.in_channels;
  k_ec_mel_to_f32<<<(int)std::min<size_t>((n + 255) / 256, 1024), 256, 0, st>>>(mel, mel_dtype, (float*)e->mel_stage, n);
  BVG_CUDA(cudaGetLastError());
  cudaGraphExec_t exec = nullptr;
  for (auto& g : e->graphs)
    if (g.B == B && g.T == Tm) exec = g.exec;
  if (!exec) {
    if (!e->cap) BVG_CUDA(cudaStreamCreateWithFlags(&e->cap, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(e->cap, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (ce == cudaSuccess) {
      rc = ec_run(e, (const float*)e->mel_stage, B, Tm, e->emb_stage, e->cap);
      ce = cudaStreamEndCapture(e->cap, &graph);
      if (rc == 0 && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    if (rc) return rc;
    if (ce != cudaSuccess || !exec) {          // capture unavailable: plain launches (same kernels)
      cudaGetLastError();
      if ((rc = ec_run(e, (const float*)e->mel_stage, B, Tm, e->emb_stage, st))) return rc;
      BVG_CUDA(cudaMemcpyAsync(emb_out, e->emb_stage, (size_t)B * d.emb_dim * 4, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    if (e->graphs.size() >= 16) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
    e->graphs.push_back({B, Tm, mel_dtype, exec});
  }
  BVG_CUDA(cudaGraphLaunch(exec, st));
  BVG_CUDA(cudaMemcpyAsync(emb_out, e->emb_stage, (size_t)B * d.emb_dim * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
