// decode_tc.cu — host side of the bf16 tcgen05/TMEM path: weight tiling, TMA tensor maps,
// launch selection and the decode orchestration (models.py:212-252) on the blocked bf16 layout.
#include <cuda.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <tuple>
#include <vector>

#include "act_blk.cuh"
#include "amp_fir.cuh"
#include "amp_nar.cuh"
#include "amp_tc.cuh"
#include "tc_api.h"

namespace bvg {
using namespace tc;

constexpr int kSplitMinCDefault = 0;   // see split_layer()

// ------------------------------------------------------------------------------ small kernels
// folded fp32 tap-major weights [Cin][K][Cout] -> bf16 UMMA tiles [ntile][chunk][tap][4][n_tile][8]
__global__ void k_pack_wt(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wt, int Cin, int Cout,
                          int K, int n_tile, int NCH, int NT) {
  const size_t total = (size_t)NT * NCH * K * 32 * n_tile;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = idx % 8;
    size_t r = idx / 8;
    const int n = r % n_tile; r /= n_tile;
    const int kg = r % 4; r /= 4;
    const int j = r % K; r /= K;
    const int c = r % NCH;
    const int nt = r / NCH;
    const int co = nt * n_tile + n, ci = c * 32 + kg * 8 + e;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = wp[((size_t)ci * K + j) * Cout + co];
    wt[idx] = __float2bfloat16_rn(v);
  }
}

// ConvTranspose1d weights [Cin][KK][Cout] (tap-major fp32) -> bf16 tiles for the 2-tap implicit GEMM:
// column n = r*Cout + co (phase-major), tap j reads x[q-(M-1-j)] with kernel index kk = r + (M-1-j)*u
__global__ void k_pack_wt_tr(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wt, int Cin, int Cout,
                             int KK, int U, int n_tile, int NCH, int NT) {
  const int M = KK / U;
  const size_t total = (size_t)NT * NCH * M * 32 * n_tile;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = idx % 8;
    size_t r0 = idx / 8;
    const int n = r0 % n_tile; r0 /= n_tile;
    const int kg = r0 % 4; r0 /= 4;
    const int j = r0 % M; r0 /= M;
    const int c = r0 % NCH;
    const int nt = r0 / NCH;
    const int col = nt * n_tile + n, ci = c * 32 + kg * 8 + e;
    float v = 0.f;
    if (col < U * Cout && ci < Cin) {
      const int r = col / Cout, co = col % Cout;
      const int kk = r + (M - 1 - j) * U;
      v = wp[((size_t)ci * KK + kk) * Cout + co];
    }
    wt[idx] = __float2bfloat16_rn(v);
  }
}

// identity "weights" in the layout of a 1-tap conv, [ntile][chunk][1][4][n_tile][8]: the residual add as D += R x I
__global__ void k_pack_identity(__nv_bfloat16* __restrict__ wt, int C, int n_tile, int NCH, int NT) {
  const size_t total = (size_t)NT * NCH * 32 * n_tile;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = idx % 8;
    size_t r = idx / 8;
    const int n = r % n_tile; r /= n_tile;
    const int kg = r % 4; r /= 4;
    const int c = r % NCH;
    const int nt = r / NCH;
    const int co = nt * n_tile + n, ci = c * 32 + kg * 8 + e;
    wt[idx] = __float2bfloat16_rn((co < C && ci < C && co == ci) ? 1.f : 0.f);
  }
}

// a2 = 2*exp(alpha), nhb = -0.5/(exp(beta)+1e-9), zero-padded to a multiple of 32 channels
__global__ void k_tc_params(const float* a, const float* invb, float* a2, float* nhb, int C, int Cpad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cpad) return;
  a2[i] = i < C ? 2.0f * a[i] : 0.f;
  nhb[i] = i < C ? -0.5f * invb[i] : 0.f;
}

// latent [B][Tmax][C] (f32/bf16/f16, time-major as the GPT emits it, gpt/model.py:459-474) ->
// blocked bf16 [B][C/8][Tmax][8], zero beyond each utterance's length.
__global__ void k_latent_blk(const void* __restrict__ lat, int dtype, __nv_bfloat16* __restrict__ out, int C,
                             int Tmax, const int* __restrict__ lengths) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, cg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Tmax) return;
  const int T = lengths ? lengths[b] : Tmax;
  __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    v[e] = __float2bfloat16_rn(t < T ? ld_dyn(lat, ((size_t)b * Tmax + t) * C + cg * 8 + e, dtype) : 0.f);
  *reinterpret_cast<uint4*>(out + (((size_t)b * (C >> 3) + cg) * Tmax + t) * 8) = *reinterpret_cast<uint4*>(v);
}

// per-op helpers: channel-major fp32 [B][C][T] <-> blocked bf16
__global__ void k_cm_to_blk(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int T) {
  const int b = blockIdx.z, cg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  __nv_bfloat16 v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(x[((size_t)b * C + cg * 8 + e) * T + t]);
  *reinterpret_cast<uint4*>(out + (((size_t)b * (C >> 3) + cg) * T + t) * 8) = *reinterpret_cast<uint4*>(v);
}
__global__ void k_blk_to_cm(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int C, int T) {
  const int b = blockIdx.z, cg = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint4 raw = *reinterpret_cast<const uint4*>(x + (((size_t)b * (C >> 3) + cg) * T + t) * 8);
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
  for (int e = 0; e < 8; ++e) out[((size_t)b * C + cg * 8 + e) * T + t] = __bfloat162float(v[e]);
}

// conv_post (C -> 1 channel, k taps, zero padding) + tanh (models.py:249-250) over z = activation_post(x), z blocked
// bf16 [B][C/8][Tstride][8] as k_act_blk leaves it.  One thread per output sample; the k x C/8 16-byte rows it reads
// are shared with its neighbours through L1.  Output in the caller's dtype (int16 = infer.py:892,911), zero past T_b.
__global__ void __launch_bounds__(256) k_conv_post_blk(const __nv_bfloat16* __restrict__ z, const float* __restrict__ wp,
                                                       const float* __restrict__ bias, void* __restrict__ out, int out_dtype,
                                                       int groups, int K, int Tstride, const int* __restrict__ lengths,
                                                       int rate) {
  __shared__ float ws[8 * 16 * 16];                 // [C][K] tap weights, C <= 128, K <= 16
  pdl_launch_dependents();
  const int C = groups * 8;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) ws[i] = wp[i];   // wp is [Cin][K][Cout = 1]
  __syncthreads();
  pdl_wait();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Tstride) return;
  const int T = lengths ? lengths[b] * rate : Tstride;
  float acc = 0.f;
  if (t < T) {
    acc = bias[0];
    const int hk = (K - 1) / 2;
    for (int j = 0; j < K; ++j) {
      const int tt = t + j - hk;
      if (tt < 0 || tt >= T) continue;
      for (int g = 0; g < groups; ++g) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * groups + g) * Tstride + tt) * 8));
        const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc = fmaf(ws[(g * 8 + 2 * e) * K + j], __uint_as_float(wds[e] << 16), acc);
          acc = fmaf(ws[(g * 8 + 2 * e + 1) * K + j], __uint_as_float(wds[e] & 0xffff0000u), acc);
        }
      }
    }
    asm("tanh.approx.f32 %0, %1;" : "=f"(acc) : "f"(acc));
  }
  st_dyn(out, (size_t)b * Tstride + t, out_dtype, acc);
}

// The same for K = 7 (config.yaml's conv_post), four consecutive outputs per thread: the 10 x C/8 rows they need are
// loaded once (16-byte loads, 7.5 per output instead of 21) and every tap weight read from shared memory serves four
// FMAs — the one-output form issues one LDS per FMA and is bound by exactly that (154 us for 16 x 10 s).
template <int K>
__global__ void __launch_bounds__(256) k_conv_post_blk4(const __nv_bfloat16* __restrict__ z, const float* __restrict__ wp,
                                                        const float* __restrict__ bias, void* __restrict__ out, int out_dtype,
                                                        int groups, int Tstride, const int* __restrict__ lengths, int rate) {
  __shared__ float ws[8 * 16 * K];                  // [C][K] tap weights, C <= 128
  pdl_launch_dependents();
  const int C = groups * 8;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) ws[i] = wp[i];   // wp is [Cin][K][Cout = 1]
  __syncthreads();
  pdl_wait();
  constexpr int NT = 4, hk = (K - 1) / 2, NR = NT + K - 1;
  const int b = blockIdx.y;
  const int t0 = (blockIdx.x * blockDim.x + threadIdx.x) * NT;
  if (t0 >= Tstride) return;
  const int T = lengths ? lengths[b] * rate : Tstride;
  float acc[NT];
  const float b0 = bias[0];
#pragma unroll
  for (int o = 0; o < NT; ++o) acc[o] = b0;
  if (t0 < T) {
    for (int g = 0; g < groups; ++g) {
      uint4 rows[NR];
      const __nv_bfloat16* zg = z + ((size_t)b * groups + g) * Tstride * 8;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int tt = t0 + r - hk;
        rows[r] = (tt >= 0 && tt < T) ? __ldg(reinterpret_cast<const uint4*>(zg + (size_t)tt * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float w = ws[(g * 8 + e) * K + j];
#pragma unroll
          for (int o = 0; o < NT; ++o) {
            const uint4& rw = rows[o + j];
            const uint32_t wd = (e >> 1) == 0 ? rw.x : (e >> 1) == 1 ? rw.y : (e >> 1) == 2 ? rw.z : rw.w;
            acc[o] = fmaf(w, __uint_as_float((e & 1) ? (wd & 0xffff0000u) : (wd << 16)), acc[o]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NT; ++o) {
    const int t = t0 + o;
    if (t >= Tstride) break;
    float v = 0.f;
    if (t < T) asm("tanh.approx.f32 %0, %1;" : "=f"(v) : "f"(acc[o]));
    st_dyn(out, (size_t)b * Tstride + t, out_dtype, v);
  }
}

// ------------------------------------------------------------------------------ plan state
struct TcLayer {
  __nv_bfloat16* wt = nullptr;
  float* a2 = nullptr;
  float* nhb = nullptr;
  int n_tile = 0, n_tiles = 0, tps = 1, tmem_cols = 32, nch = 0;
  __nv_bfloat16* idw = nullptr;   // identity tiles (square layers only): residual add on the tensor core
};

struct GraphEntry {                 // one captured tc_middle (see tc_decode)
  int B = 0, Tmax = 0;
  bool ragged = false, failed = false;
  std::vector<int32_t> lens;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
};

struct TcPlan {
  TcLayer pre;
  TcLayer ups[kMaxStages];
  TcLayer rb1[kMaxBlocks][BVG_MAX_DIL], rb2[kMaxBlocks][BVG_MAX_DIL];
  void* lat_blk = nullptr;
  size_t lat_bytes = 0;
  // the 3 AMP blocks of a stage are independent given xin (models.py:239-244): they run on three
  // streams (caller's + 2 owned) so their persistent grids back-fill each other's tails
  void* cbuf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* post_a2 = nullptr;                      // activation_post snake parameters in k_act_blk's form
  float* post_nhb = nullptr;
  void* zbuf[3] = {nullptr, nullptr, nullptr};   // Activation1d output of the split layers (one per block stream)
  size_t zbuf_bytes = 0;
  size_t cbuf_bytes = 0;
  cudaStream_t aux[2] = {nullptr, nullptr};
  void* shard = nullptr;            // ShardState (time-split P2P decode)
  cudaEvent_t ev_fork = nullptr, ev_last[BVG_MAX_KERNELS] = {};
  std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> maps;
  std::vector<void*> owned;
  std::vector<GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  unsigned long long graph_gen = 0;
};

static void pick_tile(int Cout, int* n_tile, int* n_tiles) {
  if (Cout <= 256) {
    *n_tile = std::max(16, (Cout + 15) / 16 * 16);
    *n_tiles = 1;
    return;
  }
  for (int nt = (Cout + 255) / 256;; ++nt) {
    if (Cout % nt == 0 && (Cout / nt) % 16 == 0) {
      *n_tile = Cout / nt;
      *n_tiles = nt;
      return;
    }
  }
}

static int pow2_cols(int need) {
  int c = 32;
  while (c < need) c <<= 1;
  return c;
}

static int build_layer(std::vector<void*>& owned, TcLayer& L, const ConvW& cw, const ActW* aw, cudaStream_t st) {
  pick_tile(cw.Cout, &L.n_tile, &L.n_tiles);
  L.nch = (cw.Cin + KC - 1) / KC;
  L.tps = std::max(1, std::min(cw.K, W_STAGE_BYTES / (L.n_tile * 64)));
  L.tmem_cols = pow2_cols(2 * L.n_tile);
  const size_t elems = (size_t)L.n_tiles * L.nch * cw.K * 32 * L.n_tile;
  if (!L.wt) {
    BVG_CUDA(cudaMalloc((void**)&L.wt, elems * sizeof(__nv_bfloat16)));
    owned.push_back(L.wt);
  }
  k_pack_wt<<<(int)std::min<size_t>((elems + 255) / 256, 8192), 256, 0, st>>>(cw.wp, L.wt, cw.Cin, cw.Cout, cw.K,
                                                                             L.n_tile, L.nch, L.n_tiles);
  BVG_CUDA(cudaGetLastError());
  if (cw.Cin == cw.Cout && (L.n_tiles == 1 || L.n_tile % 32 == 0)) {
    const size_t ie = (size_t)L.n_tiles * L.nch * 32 * L.n_tile;
    if (!L.idw) {
      BVG_CUDA(cudaMalloc((void**)&L.idw, ie * sizeof(__nv_bfloat16)));
      owned.push_back(L.idw);
    }
    k_pack_identity<<<(int)std::min<size_t>((ie + 255) / 256, 4096), 256, 0, st>>>(L.idw, cw.Cout, L.n_tile, L.nch, L.n_tiles);
    BVG_CUDA(cudaGetLastError());
  }
  if (aw) {
    const int Cpad = L.nch * KC;
    if (!L.a2) {
      BVG_CUDA(cudaMalloc((void**)&L.a2, Cpad * sizeof(float)));
      BVG_CUDA(cudaMalloc((void**)&L.nhb, Cpad * sizeof(float)));
      owned.push_back(L.a2);
      owned.push_back(L.nhb);
    }
    k_tc_params<<<ceil_div(Cpad, 128), 128, 0, st>>>(aw->a, aw->invb, L.a2, L.nhb, cw.Cin, Cpad);
    BVG_CUDA(cudaGetLastError());
  }
  return 0;
}

// ConvTranspose1d(Cin, Cout, KK, stride U) as an implicit GEMM with M = KK/U taps and U*Cout columns
static int build_layer_tr(std::vector<void*>& owned, TcLayer& L, const ConvW& cw, int U, cudaStream_t st) {
  const int M = cw.K / U, ncols = U * cw.Cout;
  pick_tile(ncols, &L.n_tile, &L.n_tiles);
  L.nch = (cw.Cin + KC - 1) / KC;
  L.tps = std::max(1, std::min(M, W_STAGE_BYTES / (L.n_tile * 64)));
  L.tmem_cols = pow2_cols(2 * L.n_tile);
  const size_t elems = (size_t)L.n_tiles * L.nch * M * 32 * L.n_tile;
  if (!L.wt) {
    BVG_CUDA(cudaMalloc((void**)&L.wt, elems * sizeof(__nv_bfloat16)));
    owned.push_back(L.wt);
  }
  k_pack_wt_tr<<<(int)std::min<size_t>((elems + 255) / 256, 8192), 256, 0, st>>>(cw.wp, L.wt, cw.Cin, cw.Cout, cw.K, U,
                                                                                L.n_tile, L.nch, L.n_tiles);
  BVG_CUDA(cudaGetLastError());
  return 0;
}

int tc_plan_pack(bvg_plan* p, cudaStream_t st) {
  if (!p->tc) p->tc = new TcPlan();
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  int rc;
  if ((rc = build_layer(t->owned, t->pre, p->conv_pre, nullptr, st))) return rc;
  for (int i = 0; i < p->n_stages; ++i)
    if ((rc = build_layer_tr(t->owned, t->ups[i], p->ups[i], p->cfg.upsample_rates[i], st))) return rc;
  {
    const int Cp = p->C[p->n_stages], Cpad = (Cp + KC - 1) / KC * KC;
    if (!t->post_a2) {
      BVG_CUDA(cudaMalloc((void**)&t->post_a2, Cpad * sizeof(float)));
      BVG_CUDA(cudaMalloc((void**)&t->post_nhb, Cpad * sizeof(float)));
      t->owned.push_back(t->post_a2);
      t->owned.push_back(t->post_nhb);
    }
    k_tc_params<<<ceil_div(Cpad, 128), 128, 0, st>>>(p->act_post.a, p->act_post.invb, t->post_a2, t->post_nhb, Cp, Cpad);
    BVG_CUDA(cudaGetLastError());
  }
  for (int i = 0; i < p->n_stages; ++i)
    for (int j = 0; j < p->cfg.num_kernels; ++j) {
      const int n = i * p->cfg.num_kernels + j;
      for (int m = 0; m < BVG_MAX_DIL; ++m) {
        const int K = p->rb1[n][m].K, d = p->cfg.resblock_dilation_sizes[j][m];
        if (d * (K - 1) / 2 > 25 || (K - 1) / 2 > 25)
          return fail(BVG_ERR_UNSUPPORTED, "tcgen05 path supports conv halos up to 25 samples (k=%d, d=%d)", K, d);
        if ((rc = build_layer(t->owned, t->rb1[n][m], p->rb1[n][m], &p->rba[n][2 * m], st))) return rc;
        if ((rc = build_layer(t->owned, t->rb2[n][m], p->rb2[n][m], &p->rba[n][2 * m + 1], st))) return rc;
      }
    }
  return 0;
}

void tc_shard_free(TcPlan* t);

void tc_plan_free(bvg_plan* p) {
  if (!p->tc) return;
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  tc_shard_free(t);
  for (void* q : t->owned) cudaFree(q);
  if (t->lat_blk) cudaFree(t->lat_blk);
  for (void* q : t->cbuf) if (q) cudaFree(q);
  for (void* q : t->zbuf) if (q) cudaFree(q);
  for (cudaStream_t q : t->aux) if (q) cudaStreamDestroy(q);
  for (GraphEntry& g : t->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (t->cap_stream) cudaStreamDestroy(t->cap_stream);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  for (cudaEvent_t e : t->ev_last) if (e) cudaEventDestroy(e);
  delete t;
  p->tc = nullptr;
}

int64_t tc_plan_workspace_bytes(const bvg_plan* p) {
  if (!p->tc) return 0;
  const TcPlan* t = static_cast<const TcPlan*>(p->tc);
  return (int64_t)(t->lat_bytes + 6 * t->cbuf_bytes + 3 * t->zbuf_bytes);
}

// ------------------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(f);
  }();
  return fn;
}

// blocked bf16 activation buffer [B][C/8][Tstride][8] as a 4-D tensor {8, Tstride, C/8, B};
// box {8, BOXR, 1, 1}; out-of-range rows / channel groups read as zero.
static int make_map(const void* base, int C, int Tstride, int B, CUtensorMap* out, int rows = 0, int box_rows = BOXR) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(BVG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // `rows` < Tstride when `base` points into the middle of a buffer (time-split windows): rows past
  // the end of the allocation are then out of bounds for the TMA unit (zero-filled, never fetched)
  cuuint64_t dims[4] = {8, (cuuint64_t)(rows > 0 ? rows : Tstride), (cuuint64_t)(C / 8), (cuuint64_t)B};
  cuuint64_t strides[3] = {16, (cuuint64_t)Tstride * 16, (cuuint64_t)(C / 8) * Tstride * 16};
  cuuint32_t box[4] = {8, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BVG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) C=%d T=%d B=%d", (int)r, C, Tstride, B);
  return 0;
}

// Cached per plan; the map is COPIED out (128 B), so no caller ever holds a reference into the cache, and the cache is
// only ever emptied between decodes (tc_prepare) — a decode adds ~80 keys per distinct (Tmax, B) shape.
constexpr size_t kMapCacheMax = 4096;
static int get_map(TcPlan* t, const void* base, int C, int Tstride, int B, CUtensorMap* out, int rows = 0,
                   int box_rows = BOXR) {
  auto key = std::make_tuple(base, C, Tstride, B, rows, box_rows);
  auto it = t->maps.find(key);
  if (it == t->maps.end()) {
    CUtensorMap m;
    int rc = make_map(base, C, Tstride, B, &m, rows, box_rows);
    if (rc) return rc;
    it = t->maps.emplace(key, m).first;
  }
  *out = it->second;
  return 0;
}

#ifdef BVG_EXPERIMENTS
static long long* g_trace = nullptr;
extern "C" int bvg_exp_dump_trace(void) {
  if (!g_trace) return -1;
  std::vector<long long> h(8192);
  cudaDeviceSynchronize();
  cudaMemcpy(h.data(), g_trace, 4096 * 16, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 4096; ++i)
    if (h[2 * i]) fprintf(stderr, "nar_trace %d %lld %lld\n", i, h[2 * i], h[2 * i + 1]);
  return 0;
}
#endif

// ------------------------------------------------------------------------------ launch
struct TcLaunch {
  const void* x;                // blocked bf16 input
  const __nv_bfloat16* resid = nullptr;
  const __nv_bfloat16* acc_in = nullptr;
  __nv_bfloat16* out = nullptr;
  float div = 1.f;
  const float* bias_b = nullptr;
  int bias_b_stride = 0;
  int dil = 1;
  int B = 1, Tstride = 0, rate = 1;
  const int* d_len = nullptr;
  int cls = 0;
  const int32_t* h_len = nullptr;   // host copy of the lengths (grid sizing)
  int sm_count = 0;
  int up = 0, pad = 0, cphase = 0;   // ConvTranspose1d mode
  int out_tstride = 0;               // defaults to Tstride
  int st_lo = 0, st_hi = 0x7fffffff; // store range (rows) of a conv-mode launch
  __nv_bfloat16* zbuf = nullptr;     // scratch for the split form (act_blk.cuh), same geometry as x
  int acc_rows = 0;                  // rows that exist behind acc_in / out (time-split windows); 0 = Tstride
};

template <int L, bool ACT, bool RM, int EPI>
static int launch_inst_rm(const CUtensorMap& map, const CUtensorMap& mapr, const CUtensorMap& mapq, const TcArgs& a, dim3 grid,
                          cudaStream_t st) {
  auto kern = k_amp_tc<L, ACT, RM, EPI>;
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int smem = smem_bytes(a.nx, a.nz, a.wst);
  if (smem > 227 * 1024) return fail(BVG_ERR_STATE, "k_amp_tc smem plan %d B exceeds 227 KB", smem);
  // Whatever the rings leave of the 228 KB goes to L1, which is what serves the register spills of the 32- / 64-register
  // roles (local memory) and the few global scalars: ask for the smallest carve-out that holds this launch.
  const int want = std::min(100, (smem + 1024) * 100 / (228 * 1024) + 1);
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributePreferredSharedMemoryCarveout, want));
  BVG_CUDA(launch_kc(kern, grid, dim3(NTHREADS), (size_t)smem, st, true, a.cl, map, mapr, mapq, a));
  return 0;
}

// RM = the residual-add-by-identity-MMA code is compiled in (it costs the single-thread roles registers, and the narrow
// many-tap layers that do not use it are bound by exactly those threads)
template <int L, bool ACT>
static int launch_inst(const CUtensorMap& map, const CUtensorMap& mapr, const CUtensorMap& mapq, const TcArgs& a, dim3 grid,
                       cudaStream_t st) {
  if (a.up) {
    if constexpr (!ACT) return launch_inst_rm<L, false, false, 2>(map, mapr, mapq, a, grid, st);
    else return fail(BVG_ERR_STATE, "ConvTranspose1d launches are never activated");
  }
  const bool rm = a.rmma_r || a.rmma_q;
  if (a.n_tiles == 1)
    return rm ? launch_inst_rm<L, ACT, true, 0>(map, mapr, mapq, a, grid, st) : launch_inst_rm<L, ACT, false, 0>(map, mapr, mapq, a, grid, st);
  return rm ? launch_inst_rm<L, ACT, true, 1>(map, mapr, mapq, a, grid, st) : launch_inst_rm<L, ACT, false, 1>(map, mapr, mapq, a, grid, st);
}

template <int NUB>
static int launch_fir_inst(const CUtensorMap& map, const TcArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = fir::k_amp_fir<NUB>;
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fir::F_SMEM));
  kern<<<grid, fir::NTHREADS_F, fir::F_SMEM, st>>>(map, a);
  BVG_CUDA(cudaGetLastError());
  return 0;
}

template <int NUB, bool RM>
static int launch_nar_inst(const CUtensorMap& map, const CUtensorMap& mapr, const CUtensorMap& mapq, const TcArgs& a, dim3 grid,
                           cudaStream_t st) {
  auto kern = nar::k_amp_nar<NUB, RM>;
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nar::N_SMEM));
  kern<<<grid, nar::NTHREADS_N, nar::N_SMEM, st>>>(map, mapr, mapq, a);
  BVG_CUDA(cudaGetLastError());
  return 0;
}

// Square activated layers with C <= this many channels take k_amp_nar (amp_nar.cuh: both FIRs of Activation1d streamed
// through the tensor cores); 0 = every layer on k_amp_tc.  Default 0: parity-green (50 dB per layer, like k_amp_tc) but
// measured 19.9 ms against 13.2 ms for the narrow stages of a 16 x 10 s decode (the single-thread MMA
// issuers and the per-block hand-overs are latency bound, and halo / idle-lane waste eats the instruction savings).
constexpr int kNarMaxCDefault = 0;
static std::atomic<int> g_nar_max_c{-1};
int tc_set_nar_max_c(int v) {
  const int cur = g_nar_max_c.load(), old = cur < 0 ? kNarMaxCDefault : cur;
  g_nar_max_c = v < 0 ? 0 : v;
  return old;
}
static int nar_max_c() {
  if (g_nar_max_c < 0) {
    const char* e = getenv("BVG_NAR_MAX_C");
    g_nar_max_c = e ? atoi(e) : kNarMaxCDefault;
  }
  return g_nar_max_c;
}

// narrow activated layers may take k_amp_fir (both FIRs on the tensor cores): opt-in, see bvg_set_tc_fir_max_channels
static std::atomic<int> g_fir_max_c{-1};
int tc_set_fir_max_c(int v) {
  const int cur = g_fir_max_c.load(), old = cur < 0 ? 0 : cur;
  g_fir_max_c = v < 0 ? 0 : v;
  return old;
}
static int fir_max_c() {
  if (g_fir_max_c < 0) {
    const char* e = getenv("BVG_FIR_MAX_C");
    g_fir_max_c = e ? atoi(e) : 0;
  }
  return g_fir_max_c;
}

// Layers with C_in >= this many channels run Activation1d once in k_act_blk and the conv as k_amp_tc<ACT = false>
// (act_blk.cuh) instead of repeating the activation per column tile inside the fused kernel.  0 = never.
static std::atomic<int> g_split_min_c{-1};
int tc_set_split_min_c(int v) {
  const int cur = g_split_min_c.load(), old = cur < 0 ? kSplitMinCDefault : cur;
  g_split_min_c = v < 0 ? 0 : v;
  return old;
}
static int split_min_c() {
  if (g_split_min_c < 0) {
    const char* e = getenv("BVG_SPLIT_MIN_C");
    g_split_min_c = e ? atoi(e) : kSplitMinCDefault;
  }
  return g_split_min_c;
}
static bool split_layer(int Cin) { return split_min_c() > 0 && Cin >= split_min_c(); }

// Thread-block clusters for the multi-column-tile activated layers (see launch_tc); default on.
static std::atomic<int> g_cluster{-1};
int tc_set_cluster(int on) {
  const int cur = g_cluster.load(), old = cur < 0 ? 1 : cur;
  g_cluster = on ? 1 : 0;
  return old;
}
static bool cluster_on() {
  if (g_cluster < 0) {
    const char* e = getenv("BVG_CLUSTER");
    g_cluster = (!e || atoi(e) != 0) ? 1 : 0;
  }
  return g_cluster != 0;
}

// Residual / running-sum add by identity MMAs (amp_tc.cuh) on the layers where it pays; off = always in the epilogue.
static std::atomic<int> g_rmma{-1};
int tc_set_residual_mma(int on) {
  const int cur = g_rmma.load(), old = cur < 0 ? 1 : cur;
  g_rmma = on ? 1 : 0;
  return old;
}
static bool residual_mma_on() {
  if (g_rmma < 0) {
    const char* e = getenv("BVG_RMMA");
    g_rmma = (!e || atoi(e) != 0) ? 1 : 0;
  }
  return g_rmma != 0;
}

static int launch_act_blk(bvg_plan* p, const ConvW& cw, const ActW* aw, const TcLayer& L, const TcLaunch& q,
                          cudaStream_t st) {
  AbArgs b{};
  b.x = static_cast<const __nv_bfloat16*>(q.x); b.z = q.zbuf; b.a2 = L.a2; b.nhb = L.nhb; b.lengths = q.d_len;
  b.B = q.B; b.groups = cw.Cin / 8; b.Tstride = q.Tstride; b.rate = q.rate; b.Tmax = q.Tstride;
  b.nslices = (q.Tstride + AB_V - 1) / AB_V;
  for (int i = 0; i < 12; ++i) { b.up2[i] = 2.0f * aw->up[i]; b.dn[i] = aw->dn[i]; }
  const long long units = (long long)b.B * b.groups * b.nslices;
  if (units >= (1LL << 31)) return fail(BVG_ERR_UNSUPPORTED, "k_act_blk: %lld work units exceed 2^31", units);
  const int sms = q.sm_count > 0 ? q.sm_count : 148;
  const int smem = AB_WARPS * AB_SMEM_PER_WARP + 8192;   // tail slack: the edge path's clamped reads of unused rows
  BVG_CUDA(func_attr_once((const void*)k_act_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)std::min<long long>((units + AB_WARPS - 1) / AB_WARPS, 2LL * sms));
  const double samples = (p ? p->cur_sum_frames : 0.0) * q.rate;
  prof_begin(p, st, q.cls, 0.0, samples * 2.0 * 2.0 * cw.Cin);
  BVG_CUDA(launch_k(k_act_blk, grid, dim3(AB_WARPS * 32), (size_t)smem, st, true, b));
  prof_end(p, st);
  BVG_CUDA(cudaGetLastError());
  if (p) ++p->last_launches;
  return 0;
}

#ifdef BVG_EXPERIMENTS
struct RingOv { int nx = 0, nz = 0, wst = 0; };
// all three fields must parse and stay inside the mbarrier table; anything else = no override
static RingOv ring_override(const char* name) {
  RingOv o;
  int x = 0, z = 0, w = 0;
  const char* e = getenv(name);
  if (e && sscanf(e, "%d,%d,%d", &x, &z, &w) == 3 && x >= 2 && x <= NX_MAX && z >= 2 && z <= NZ_MAX && w >= 2 &&
      w <= W_STAGES_MAX) { o.nx = x; o.nz = z; o.wst = w; }
  return o;
}
#endif

static int launch_tc(bvg_plan* p, const CUtensorMap& map, const TcLayer& L, const ConvW& cw, const ActW* aw,
                     const TcLaunch& q, cudaStream_t st) {
  if (aw && !q.up && q.zbuf && split_layer(cw.Cin)) {
    // split form: z = Activation1d(x) once, then the plain dilated conv over z
    int rc;
    if ((rc = launch_act_blk(p, cw, aw, L, q, st))) return rc;
    CUtensorMap zm;
    TcPlan* t = p ? static_cast<TcPlan*>(p->tc) : nullptr;
    if (t) rc = get_map(t, q.zbuf, cw.Cin, q.Tstride, q.B, &zm);
    else rc = make_map(q.zbuf, cw.Cin, q.Tstride, q.B, &zm);
    if (rc) return rc;
    TcLaunch qc = q;
    qc.x = q.zbuf; qc.zbuf = nullptr;
    return launch_tc(p, zm, L, cw, nullptr, qc, st);
  }
  TcArgs a{};
  a.wt = L.wt; a.bias = cw.bias; a.bias_b = q.bias_b; a.bias_b_stride = q.bias_b_stride;
  a.resid = q.resid; a.acc_in = q.acc_in; a.out = q.out; a.div = q.div;
  a.Cin = cw.Cin; a.Cout = cw.Cout; a.K = cw.K; a.dil = q.dil;
  a.n_tile = L.n_tile; a.n_tiles = L.n_tiles; a.taps_per_stage = L.tps; a.B = q.B;
  // ring depths.  Timing experiments (BVG_DBG) showed the weight stream is not the limiter (no change with
  // 16-byte copies) while the single TMEM accumulator stage of the wide layers (4*n_tile > 512) stalls the
  // MMAs of the next tile during the epilogue; 4 z buffers let the activation warps run ahead meanwhile.
  // ring depths (22 / 21.5 / 16 KB per slot; 32 KB go to the epilogue's residual staging).  Narrow layers stream x
  // tile after tile (one chunk per tile) and need little weight staging; wide layers walk 6-24 chunks per tile.
  if (!aw) { a.nx = 3; a.nz = 2; a.wst = 4; }                       // plain conv / ConvTranspose1d: the MMA reads the x ring
  else if (cw.Cin <= 96) { a.nx = 3; a.nz = 3; a.wst = 3; }
  // wide layers: the weight stream (16 KB per tap, 25-32 B/clk/SM at full MMA rate — most of what L2 can deliver to 148
  // SMs) is the latency-critical one: 5 stages in flight measured 4.5 % faster than 4, 3 stages 4 % slower, 2 stages 27 %
  // slower; a third x slot or a fourth z slot bought nothing (per-launch A/B, round 2)
  else { a.nx = 2; a.nz = 3; a.wst = 5; }
#ifdef BVG_EXPERIMENTS
  {
    // ring-depth experiments, "nx,nz,wst": BVG_RINGS (every launch), BVG_RINGS_NARROW (activated, C_in <= 96),
    // BVG_RINGS_MID (activated, C_in > 96, one column tile), BVG_RINGS_WIDE (activated, several column tiles)
    static const RingOv ov = ring_override("BVG_RINGS"), ovn = ring_override("BVG_RINGS_NARROW"),
                        ovm = ring_override("BVG_RINGS_MID"), ovw = ring_override("BVG_RINGS_WIDE");
    auto apply = [&](const RingOv& o) { if (o.nx > 0) { a.nx = o.nx; a.nz = o.nz; a.wst = o.wst; } };
    apply(ov);
    if (aw && cw.Cin <= 96) apply(ovn);
    if (aw && cw.Cin > 96 && L.n_tiles == 1) apply(ovm);
    if (aw && L.n_tiles > 1) apply(ovw);
  }
#endif
  a.Tstride = q.out_tstride ? q.out_tstride : q.Tstride;
  a.lengths = q.d_len; a.rate = q.rate; a.Tmax = q.Tstride;
  a.up = q.up; a.pad = q.pad; a.cphase = q.cphase;
  a.st_lo = q.st_lo; a.st_hi = q.st_hi;
#ifdef BVG_EXPERIMENTS
  {
    static const int dbg = [] { const char* e = getenv("BVG_DBG"); return e ? atoi(e) : 0; }();
    a.dbg = dbg;
    // BVG_TRACE_LAYER=<cls>,<C>,<K>,<dil>: the matching k_amp_nar launches stamp their pipeline events into a device buffer
    // that is dumped to stderr by bvg_exp_dump_trace()
    static const std::tuple<int, int, int> tl = [] {
      int c = 0, k = 0, d = 0;
      const char* e = getenv("BVG_TRACE_LAYER");
      if (!e || sscanf(e, "%d,%d,%d", &c, &k, &d) != 3) return std::make_tuple(0, 0, 0);
      return std::make_tuple(c, k, d);
    }();
    if (std::get<0>(tl) == cw.Cin && std::get<1>(tl) == cw.K && std::get<2>(tl) == q.dil && aw) {
      if (!g_trace) { cudaMalloc(&g_trace, 4096 * 16); }
      cudaMemsetAsync(g_trace, 0, 4096 * 16, st);
      a.trace = g_trace;
    }
  }
#endif
  const int hc = q.dil * (cw.K - 1) / 2;
  a.lead = q.up ? cw.K - 1 : hc;
  a.xin = static_cast<const __nv_bfloat16*>(q.x); a.xgroups = cw.Cin / 8;   // interior x tiles: 1-D bulk copies
  if (aw) {
    a.a2 = L.a2; a.nhb = L.nhb;
    for (int i = 0; i < 12; ++i) { a.up2[i] = 2.0f * aw->up[i]; a.dn[i] = aw->dn[i]; }
  }
  if (q.B > MAX_B) return fail(BVG_ERR_UNSUPPORTED, "tcgen05 path: at most %d utterances per launch", MAX_B);
  // persistent grid: one CTA per SM walking the (utterance, time tile, column tile) sequence
  const int extra = q.up ? cw.K - 1 : 0;
  long long tiles = 0;
  for (int b = 0; b < q.B; ++b) {
    const long long Tb = (q.h_len ? (long long)q.h_len[b] * q.rate : q.Tstride) + extra;
    tiles += (Tb + M_TILE - 1) / M_TILE;
  }
  const long long time_tiles = tiles;
  tiles *= L.n_tiles;
  dim3 grid((unsigned)std::min<long long>(tiles, q.sm_count > 0 ? q.sm_count : 148));
  // Cluster mode for the activated layers with 2 or 3 column tiles (C = 384, 768): the column tiles of a time tile form
  // one thread-block cluster and share ONE activation of every 32-channel chunk through distributed shared memory instead
  // of recomputing it per column tile (amp_tc.cuh).  Same arithmetic, same results.
  a.cl = 0;
  if (aw && !q.up && cluster_on() && (L.n_tiles == 2 || L.n_tiles == 3) && cw.Cin % KC == 0) {
    const long long sms = q.sm_count > 0 ? q.sm_count : 148;
    const long long ncl = std::min<long long>(time_tiles, sms / L.n_tiles);
    a.cl = L.n_tiles;
    grid = dim3((unsigned)(ncl * L.n_tiles));
  }
  const double samples = (p ? p->cur_sum_frames : 0.0) * q.rate;
  prof_begin(p, st, q.cls, 2.0 * cw.Cin * cw.Cout * cw.K * samples,
             samples * 2.0 * (cw.Cin + cw.Cout + (q.resid ? cw.Cout : 0) + (q.acc_in ? cw.Cout : 0)) +
                 2.0 * cw.Cin * cw.Cout * cw.K);
  int rc;
  const bool use_fir = aw && !q.up && L.n_tiles == 1 && L.n_tile <= fir::MAX_NTILE_F && cw.Cin <= fir_max_c() &&
                       cw.Cout <= fir_max_c() && hc <= 32;
  const bool use_nar = !use_fir && aw && !q.up && L.n_tiles == 1 && L.n_tile <= nar::MAX_NTILE_N && cw.Cin == cw.Cout &&
                       cw.Cin <= nar_max_c() && hc <= 32;
  if (use_nar) {
    // x tile = 16 TMA boxes {8 channels, 96 rows} (4 time segments x 4 channel groups); residual / running sum as D += R x I
    TcPlan* t = p ? static_cast<TcPlan*>(p->tc) : nullptr;
    CUtensorMap xm, tmpr, tmpq;
    rc = t ? get_map(t, q.x, cw.Cin, q.Tstride, q.B, &xm, 0, fir::XB) : make_map(q.x, cw.Cin, q.Tstride, q.B, &xm, 0, fir::XB);
    if (rc) return rc;
    const CUtensorMap* mr = &xm;
    const CUtensorMap* mq = &xm;
    a.xin = static_cast<const __nv_bfloat16*>(q.x);
    a.xgroups = cw.Cin / 8;
    const bool rm = residual_mma_on() && L.idw && (q.resid || q.acc_in);
    if (rm) {
      if (q.resid) {
        mr = &tmpr;
        rc = t ? get_map(t, q.resid, cw.Cout, a.Tstride, q.B, &tmpr, 0, 128) : make_map(q.resid, cw.Cout, a.Tstride, q.B, &tmpr, 0, 128);
      }
      if (!rc && q.acc_in) {
        mq = &tmpq;
        rc = t ? get_map(t, q.acc_in, cw.Cout, a.Tstride, q.B, &tmpq, q.acc_rows, 128)
               : make_map(q.acc_in, cw.Cout, a.Tstride, q.B, &tmpq, q.acc_rows, 128);
      }
      if (rc) return rc;
      a.idw = L.idw; a.nchr = (L.n_tile + 31) / 32;
      a.rmma_r = q.resid ? 1 : 0; a.rmma_q = q.acc_in ? 1 : 0;
      a.resid = nullptr; a.acc_in = nullptr;       // the epilogue warps see a plain conv
    }
    const bool plain = !(a.resid || a.acc_in);      // RM instantiation = plain-only epilogue
    if (hc <= 16) rc = plain ? launch_nar_inst<10, true>(xm, *mr, *mq, a, grid, st) : launch_nar_inst<10, false>(xm, *mr, *mq, a, grid, st);
    else rc = plain ? launch_nar_inst<11, true>(xm, *mr, *mq, a, grid, st) : launch_nar_inst<11, false>(xm, *mr, *mq, a, grid, st);
  } else if (use_fir) {
    // x tile = 16 TMA boxes {8 channels, 96 rows} (4 time segments x 4 channel groups)
    CUtensorMap fm;
    TcPlan* t = p ? static_cast<TcPlan*>(p->tc) : nullptr;
    if (t) rc = get_map(t, q.x, cw.Cin, q.Tstride, q.B, &fm, 0, fir::XB);
    else rc = make_map(q.x, cw.Cin, q.Tstride, q.B, &fm, 0, fir::XB);
    if (rc) return rc;
    a.wst = fir::W_STAGES_F;
    a.xin = static_cast<const __nv_bfloat16*>(q.x);
    a.xgroups = cw.Cin / 8;
    rc = (hc <= 16) ? launch_fir_inst<10>(fm, a, grid, st) : launch_fir_inst<11>(fm, a, grid, st);
  } else
  {
    // residual / running-sum add as identity MMAs (D += R x I) instead of loads + adds in the epilogue warps
    const bool rmma_on = residual_mma_on();
    CUtensorMap tmpr, tmpq;
    const CUtensorMap* mr = &map;       // `map` is the caller's own copy; tmpr / tmpq are ours
    const CUtensorMap* mq = &map;
    // Worth it unless the layer is bound by the MMA issue thread itself: narrow layers (small N, A-operand-fetch-bound
    // MMAs of ~50 cycles whatever N) with many taps got slower with 4-8 more MMAs and two more hand-shakes per chunk
    // (per-launch events, profiles/r01_rmma_ab.txt): threshold = 36 conv MMAs per tile.
    static const int rmma_max_mmas = [] { const char* e = getenv("BVG_RMMA_MAX_MMAS"); return e ? atoi(e) : 36; }();
    const bool rmma_pays = cw.Cin >= 192 || L.nch * cw.K * 4 <= rmma_max_mmas;
    if (rmma_on && rmma_pays && !q.up && L.idw && (q.resid || q.acc_in)) {
      TcPlan* t = p ? static_cast<TcPlan*>(p->tc) : nullptr;
      rc = 0;
      if (q.resid) {
        mr = &tmpr;
        rc = t ? get_map(t, q.resid, cw.Cout, a.Tstride, q.B, &tmpr, 0, 128) : make_map(q.resid, cw.Cout, a.Tstride, q.B, &tmpr, 0, 128);
      }
      if (!rc && q.acc_in) {
        mq = &tmpq;
        rc = t ? get_map(t, q.acc_in, cw.Cout, a.Tstride, q.B, &tmpq, q.acc_rows, 128)
               : make_map(q.acc_in, cw.Cout, a.Tstride, q.B, &tmpq, q.acc_rows, 128);
      }
      if (rc) return rc;
      a.idw = L.idw; a.nchr = (L.n_tile + 31) / 32;
      a.rmma_r = q.resid ? 1 : 0; a.rmma_q = q.acc_in ? 1 : 0;
      a.resid = nullptr; a.acc_in = nullptr;       // the epilogue warps see a plain conv
    }
  // run length per activation thread: 4 warps x (8L - 6) valid rows must cover 256 + 2*hc z rows
  if (!aw) rc = launch_inst<9, false>(map, *mr, *mq, a, grid, st);
  else if (hc <= 4) rc = launch_inst<9, true>(map, *mr, *mq, a, grid, st);     // 4*(8*9-6)  = 264 >= 256 + 2*4
  else if (hc <= 20) rc = launch_inst<10, true>(map, *mr, *mq, a, grid, st);   // 4*(8*10-6) = 296 >= 256 + 2*20 (2-way bank conflicts)
  else rc = launch_inst<11, true>(map, *mr, *mq, a, grid, st);                 // 4*(8*11-6) = 328 >= 256 + 2*25
  }
  prof_end(p, st);
  if (p) ++p->last_launches;
  return rc;
}

// ------------------------------------------------------------------------------ decode
// All activation buffers of a decode have Fs*rate rows per (utterance, channel group); `h_len` /
// `d_len` give the valid frames per utterance.  In the time-split (shard) mode Fs is the common
// frame stride of all ranks and the buffers are addressed through window-start pointers.
struct StageIO {
  const __nv_bfloat16* cur = nullptr;   // stage input, rate[i]
  __nv_bfloat16* xin = nullptr;         // ConvTranspose output (private)
  __nv_bfloat16* xs = nullptr;          // stage output, rate[i+1]
  int B = 1, Fs = 0;
  const int32_t* h_len = nullptr;
  const int* d_len = nullptr;
  int st_lo_f = 0, st_hi_f = 0x7fffff;  // frames of the window whose output rows are stored in xs
  int cur_rows = 0;                     // rows of `cur` that exist behind the pointer (0 = Fs*rate[i])
  int xs_rows = 0;                      // same for `xs`
  bool concurrent = true;
};

static int tc_prepare(bvg_plan* p, TcPlan* t, int B, int Fs) {
  int rc;
  if (t->maps.size() > kMapCacheMax) t->maps.clear();   // nothing holds a reference between decodes
  size_t max_elems = (size_t)p->C[0] * Fs;
  for (int i = 0; i < p->n_stages; ++i)
    max_elems = std::max(max_elems, (size_t)p->C[i + 1] * Fs * p->rate[i + 1]);
  const size_t buf_bytes = max_elems * B * sizeof(__nv_bfloat16);
  if ((rc = tc_ensure_ws(p, buf_bytes))) return rc;
  const size_t lat_need = (size_t)B * p->cfg.gpt_dim * Fs * sizeof(__nv_bfloat16);
  if (lat_need > t->lat_bytes) {
    BVG_CUDA(cudaDeviceSynchronize());
    if (t->lat_blk) BVG_CUDA(cudaFree(t->lat_blk));
    t->lat_blk = nullptr; t->lat_bytes = 0;
    BVG_CUDA(cudaMalloc(&t->lat_blk, lat_need));
    BVG_CUDA(cudaMemset(t->lat_blk, 0, lat_need));
    t->lat_bytes = lat_need;
    ++p->alloc_gen;
    t->maps.clear();
  }
  if (buf_bytes > t->cbuf_bytes) {
    BVG_CUDA(cudaDeviceSynchronize());
    for (void*& qb : t->cbuf) { if (qb) BVG_CUDA(cudaFree(qb)); qb = nullptr; }
    t->cbuf_bytes = 0;
    for (void*& qb : t->cbuf) {
      BVG_CUDA(cudaMalloc(&qb, buf_bytes));
      BVG_CUDA(cudaMemset(qb, 0, buf_bytes));       // see ensure_ws: stale rows must be finite
    }
    t->cbuf_bytes = buf_bytes;
    ++p->alloc_gen;
    t->maps.clear();
  }
  // scratch of the split layers: the largest split stage (channels x rows), one buffer per block stream
  size_t z_elems = 0;
  for (int i = 0; i < p->n_stages; ++i)
    if (split_layer(p->C[i + 1])) z_elems = std::max(z_elems, (size_t)p->C[i + 1] * Fs * p->rate[i + 1]);
  const size_t z_bytes = z_elems * B * sizeof(__nv_bfloat16);
  if (z_bytes > t->zbuf_bytes) {
    BVG_CUDA(cudaDeviceSynchronize());
    for (void*& qb : t->zbuf) { if (qb) BVG_CUDA(cudaFree(qb)); qb = nullptr; }
    t->zbuf_bytes = 0;
    for (void*& qb : t->zbuf) {
      BVG_CUDA(cudaMalloc(&qb, z_bytes));
      BVG_CUDA(cudaMemset(qb, 0, z_bytes));
    }
    t->zbuf_bytes = z_bytes;
    ++p->alloc_gen;
    t->maps.clear();
  }
  if (!t->ev_fork) {
    for (cudaStream_t& q : t->aux) BVG_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    BVG_CUDA(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
    for (cudaEvent_t& e : t->ev_last) BVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  return 0;
}

// latent -> blocked bf16 -> conv_pre + cond_layer add (models.py:222-228)
// parts: 1 = the re-blocking kernel (the only one that touches the caller's latent), 2 = conv_pre, 3 = both
static int tc_pre(bvg_plan* p, TcPlan* t, const void* latent, int latent_dtype, __nv_bfloat16* out, int B, int Fs,
                  const int32_t* h_len, const int* d_len, cudaStream_t st, int parts = 3) {
  int rc;
  CUtensorMap map;
  dim3 grid(ceil_div(Fs, 128), p->cfg.gpt_dim / 8, B);
  if (parts & 1) {
  prof_begin(p, st, 2, 0.0, (double)B * Fs * p->cfg.gpt_dim * 6.0);
  BVG_CUDA(launch_k(k_latent_blk, grid, dim3(128), 0, st, true, latent, latent_dtype, (__nv_bfloat16*)t->lat_blk, p->cfg.gpt_dim, Fs, d_len));
  prof_end(p, st);
  BVG_CUDA(cudaGetLastError());
  ++p->last_launches;
  }
  if (!(parts & 2)) return 0;
  if ((rc = get_map(t, t->lat_blk, p->cfg.gpt_dim, Fs, B, &map))) return rc;
  TcLaunch q;   // plain conv: the TMA tile feeds the MMA directly
  q.x = t->lat_blk; q.out = out; q.bias_b = p->condb + p->cond_off[0]; q.bias_b_stride = p->cond_total;
  q.dil = 1; q.B = B; q.Tstride = Fs; q.rate = 1; q.d_len = d_len; q.cls = 2;
  q.h_len = h_len; q.sm_count = p->sm_count;
  return launch_tc(p, map, t->pre, p->conv_pre, nullptr, q, st);
}

// one upsampling stage: ConvTranspose1d + cond add, three AMP blocks, mean (models.py:230-245)
static int tc_stage(bvg_plan* p, TcPlan* t, int i, const StageIO& io, cudaStream_t st) {
  int rc;
  CUtensorMap map;
  const int nk = p->cfg.num_kernels;
  const int B = io.B;
  const int Ci = p->C[i + 1], Ri = p->rate[i + 1], Ti = io.Fs * Ri;
  // with per-launch profiling on, the blocks are serialised on the caller's stream so that the
  // CUDA-event duration of a launch is that kernel alone (bvg_plan_set_profiling)
  const int nstreams = (p->profiling || !io.concurrent) ? 1 : std::min(nk, 3);
  cudaStream_t sj[3] = {st, t->aux[0], t->aux[1]};
  {
    // ConvTranspose1d + cond add (models.py:232-236) as a (k/u)-tap implicit GEMM on the input rate
    const int U = p->cfg.upsample_rates[i], KK = p->ups[i].K, Tin = io.Fs * p->rate[i];
    ConvW cw = p->ups[i];
    cw.Cout = U * Ci; cw.K = KK / U;
    TcLaunch qu;
    qu.x = io.cur; qu.out = io.xin; qu.dil = 1; qu.B = B; qu.Tstride = Tin; qu.out_tstride = Ti; qu.rate = p->rate[i];
    qu.d_len = io.d_len; qu.cls = 2; qu.up = U; qu.pad = (KK - U) / 2; qu.cphase = Ci;
    qu.h_len = io.h_len; qu.sm_count = p->sm_count;
    qu.bias_b = p->cfg.cond_in_each_up_layer ? p->condb + p->cond_off[i + 1] : nullptr;
    qu.bias_b_stride = p->cond_total;
    if ((rc = get_map(t, io.cur, p->C[i], Tin, B, &map, io.cur_rows))) return rc;
    if ((rc = launch_tc(p, map, t->ups[i], cw, nullptr, qu, st))) return rc;
  }
  if (nstreams > 1) {
    BVG_CUDA(cudaEventRecord(t->ev_fork, st));
    for (int js = 1; js < nstreams; ++js) BVG_CUDA(cudaStreamWaitEvent(sj[js], t->ev_fork, 0));
  }
  const int cls = (Ci >= 192) ? 0 : 1;
  for (int j = 0; j < nk; ++j) {
    const int n = i * nk + j;
    cudaStream_t sq = sj[j % nstreams];
    __nv_bfloat16* xr = (__nv_bfloat16*)t->cbuf[2 * (j % 3)];
    __nv_bfloat16* xt = (__nv_bfloat16*)t->cbuf[2 * (j % 3) + 1];
    const __nv_bfloat16* xcur = io.xin;
    for (int m = 0; m < BVG_MAX_DIL; ++m) {
      const int d = p->cfg.resblock_dilation_sizes[j][m];
      TcLaunch qa;
      qa.x = xcur; qa.out = xt; qa.dil = d; qa.B = B; qa.Tstride = Ti; qa.rate = Ri; qa.d_len = io.d_len; qa.cls = cls;
      qa.h_len = io.h_len; qa.sm_count = p->sm_count;
      qa.zbuf = split_layer(Ci) ? (__nv_bfloat16*)t->zbuf[j % 3] : nullptr;
      if ((rc = get_map(t, xcur, Ci, Ti, B, &map))) return rc;
      if ((rc = launch_tc(p, map, t->rb1[n][m], p->rb1[n][m], &p->rba[n][2 * m], qa, sq))) return rc;

      const bool last = (m == BVG_MAX_DIL - 1);
      TcLaunch qb;
      qb.x = xt; qb.resid = xcur; qb.dil = 1; qb.B = B; qb.Tstride = Ti; qb.rate = Ri; qb.d_len = io.d_len; qb.cls = cls;
      qb.h_len = io.h_len; qb.sm_count = p->sm_count;
      qb.zbuf = qa.zbuf;
      if (!last) {
        qb.out = xr;
      } else {
        // running sum over the blocks (models.py:239-245) is ordered: block j adds onto block j-1
        if (j > 0 && nstreams > 1) BVG_CUDA(cudaStreamWaitEvent(sq, t->ev_last[j - 1], 0));
        qb.out = io.xs;
        qb.acc_in = (j > 0) ? io.xs : nullptr;
        qb.acc_rows = io.xs_rows;
        qb.div = (j == nk - 1) ? (float)nk : 1.f;
        qb.st_lo = io.st_lo_f * Ri;
        qb.st_hi = (io.st_hi_f >= 0x7fffff) ? 0x7fffffff : io.st_hi_f * Ri;
      }
      if ((rc = get_map(t, xt, Ci, Ti, B, &map))) return rc;
      if ((rc = launch_tc(p, map, t->rb2[n][m], p->rb2[n][m], &p->rba[n][2 * m + 1], qb, sq))) return rc;
      if (last && nstreams > 1) BVG_CUDA(cudaEventRecord(t->ev_last[j], sq));
      xcur = xr;
    }
  }
  if (nstreams > 1) BVG_CUDA(cudaStreamWaitEvent(st, t->ev_last[nk - 1], 0));    // join
  return 0;
}

// activation_post + conv_post + tanh (models.py:248-250): the streaming Activation1d kernel, then a 24 -> 1 channel conv
// + tanh that reads its bf16 rows.  (The SIMT kernel this replaces took 1.1 ms of a 25 ms step; it still serves the
// shapes this one does not cover.)  `cur` may point into the middle of a stage buffer (time-split windows).
static bool post_is_split(const bvg_plan* p) {
  static const bool old_post = getenv("BVG_SIMT_POST") != nullptr;
  const int S = p->n_stages, Cp = p->C[S], Kp = p->conv_post.K;
  return !(old_post || Cp % 8 != 0 || Cp > 128 || Kp > 16 || p->conv_post.Cout != 1);
}

// parts: 1 = activation_post (plan-owned buffers only), 2 = conv_post + tanh (writes the caller's waveform), 3 = both
static int tc_post(bvg_plan* p, TcPlan* t, const __nv_bfloat16* cur, void* wav_out, int wav_dtype, int B, int Tmax,
                   const int* d_len, cudaStream_t st, int parts = 3) {
  int rc;
  const int S = p->n_stages, Cp = p->C[S], Kp = p->conv_post.K;
  if (!post_is_split(p)) return (parts & 2) ? simt_post_blk(p, cur, wav_out, wav_dtype, B, Tmax, d_len, st) : 0;
  const int Ts = Tmax * p->rate[S];
  __nv_bfloat16* zpost = (__nv_bfloat16*)t->cbuf[0];        // the block buffers are idle after the last stage
  TcLayer Lp;
  Lp.a2 = t->post_a2; Lp.nhb = t->post_nhb;
  ConvW cwp = p->conv_post;
  TcLaunch qp;
  qp.x = cur; qp.zbuf = zpost; qp.B = B; qp.Tstride = Ts; qp.rate = p->rate[S]; qp.d_len = d_len; qp.cls = 3;
  qp.sm_count = p->sm_count;
  if ((parts & 1) && (rc = launch_act_blk(p, cwp, &p->act_post, Lp, qp, st))) return rc;
  if (!(parts & 2)) return 0;
  dim3 gridp(ceil_div(Ts, 256), B);
  prof_begin(p, st, 3, 2.0 * Cp * Kp * p->cur_sum_frames * p->rate[S], (2.0 * Cp + 4.0) * p->cur_sum_frames * p->rate[S]);
  if (Kp == 7) {
    dim3 grid4(ceil_div(Ts, 256 * 4), B);
    BVG_CUDA(launch_k(k_conv_post_blk4<7>, grid4, dim3(256), 0, st, true, (const __nv_bfloat16*)zpost, (const float*)p->conv_post.wp,
                      (const float*)p->conv_post.bias, wav_out, wav_dtype, Cp / 8, Ts, d_len, p->rate[S]));
  } else {
    BVG_CUDA(launch_k(k_conv_post_blk, gridp, dim3(256), 0, st, true, (const __nv_bfloat16*)zpost, (const float*)p->conv_post.wp,
                      (const float*)p->conv_post.bias, wav_out, wav_dtype, Cp / 8, Kp, Ts, d_len, p->rate[S]));
  }
  prof_end(p, st);
  BVG_CUDA(cudaGetLastError());
  ++p->last_launches;
  return 0;
}

// Everything of a decode that only touches plan-owned buffers: conv_pre, the six stages, activation_post.
static int tc_middle(bvg_plan* p, TcPlan* t, int B, int Tmax, const int32_t* h_len, const int* d_len, cudaStream_t st) {
  int rc;
  __nv_bfloat16* cur = (__nv_bfloat16*)p->ws[0];
  if ((rc = tc_pre(p, t, nullptr, 0, cur, B, Tmax, h_len, d_len, st, 2))) return rc;
  for (int i = 0; i < p->n_stages; ++i) {
    StageIO io;
    io.cur = cur; io.xin = (__nv_bfloat16*)p->ws[1]; io.xs = cur;   // cur is dead once the ConvTranspose1d consumed it
    io.B = B; io.Fs = Tmax; io.h_len = h_len; io.d_len = d_len;
    if ((rc = tc_stage(p, t, i, io, st))) return rc;
  }
  return tc_post(p, t, cur, nullptr, 0, B, Tmax, d_len, st, 1);
}

// CUDA-graph replay of tc_middle.  A decode is 118 launches on three streams; on short utterances the launches, the
// fork / join events and the grid-launch latency between dependent kernels are a large part of the wall time.  The
// middle of a decode only reads and writes plan-owned buffers at fixed addresses, so for a given (B, Tmax, lengths) it
// is the same launch sequence every time: it is stream-captured (on a plan-owned stream: the caller's may be the
// legacy default stream, which cannot be captured) the SECOND time a shape is seen, instantiated, and replayed on the
// caller's stream afterwards.  The programmatic-dependent-launch edges survive the capture.  The kernels that touch
// the caller's pointers (latent re-blocking, speaker-conditioning GEMV, conv_post) stay ordinary launches around it.
static std::atomic<int> g_graphs{-1};
int tc_set_graphs(int on) {
  const int cur = g_graphs.load(), old = cur < 0 ? 1 : cur;
  g_graphs = on ? 1 : 0;
  return old;
}
static bool graphs_on() {
  if (g_graphs < 0) {
    const char* e = getenv("BVG_GRAPHS");
    g_graphs = (!e || atoi(e) != 0) ? 1 : 0;
  }
  return g_graphs != 0;
}
constexpr size_t kMaxGraphs = 64;

static GraphEntry* find_graph(TcPlan* t, int B, int Tmax, const int32_t* h_len) {
  for (GraphEntry& g : t->graphs) {
    if (g.B != B || g.Tmax != Tmax || g.ragged != (h_len != nullptr)) continue;
    if (h_len && memcmp(g.lens.data(), h_len, sizeof(int32_t) * B) != 0) continue;
    return &g;
  }
  return nullptr;
}

int tc_decode(bvg_plan* p, const void* latent, int latent_dtype, const int32_t* h_len, const int* d_len, int B,
              int Tmax, void* wav_out, int wav_dtype, cudaStream_t st) {
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  if (!t) return fail(BVG_ERR_STATE, "tcgen05 path: weights not packed");
  int rc;
  if ((rc = tc_prepare(p, t, B, Tmax))) return rc;
  if (t->graph_gen != p->alloc_gen) {          // some buffer moved since the graphs were captured
    for (GraphEntry& g : t->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    t->graphs.clear();
    t->graph_gen = p->alloc_gen;
  }
  __nv_bfloat16* cur = (__nv_bfloat16*)p->ws[0];
  if ((rc = tc_pre(p, t, latent, latent_dtype, cur, B, Tmax, h_len, d_len, st, 1))) return rc;
  bool done = false;
  if (graphs_on() && !p->profiling && post_is_split(p)) {
    GraphEntry* g = find_graph(t, B, Tmax, h_len);
    if (!g) {                                   // first sighting: run it eagerly (sets function attributes, fills caches)
      if (t->graphs.size() >= kMaxGraphs) {
        if (t->graphs.front().exec) cudaGraphExecDestroy(t->graphs.front().exec);
        t->graphs.erase(t->graphs.begin());
      }
      GraphEntry e;
      e.B = B; e.Tmax = Tmax; e.ragged = h_len != nullptr;
      if (h_len) e.lens.assign(h_len, h_len + B);
      t->graphs.push_back(std::move(e));
    } else if (!g->exec && !g->failed) {        // second sighting: capture
      if (!t->cap_stream) BVG_CUDA(cudaStreamCreateWithFlags(&t->cap_stream, cudaStreamNonBlocking));
      const int before = p->last_launches;
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaStreamBeginCapture(t->cap_stream, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        rc = tc_middle(p, t, B, Tmax, h_len, d_len, t->cap_stream);
        ce = cudaStreamEndCapture(t->cap_stream, &graph);
        if (rc == 0 && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g->exec, graph, 0);
        else if (ce == cudaSuccess) ce = cudaErrorUnknown;
        if (graph) cudaGraphDestroy(graph);
      }
      g->launches = p->last_launches - before;
      p->last_launches = before;
      if (ce != cudaSuccess || !g->exec) {       // capture is an optimisation: fall back to plain launches for this shape
        g->exec = nullptr;
        g->failed = true;
        cudaGetLastError();
      }
    }
    if (g && g->exec) {
      BVG_CUDA(cudaGraphLaunch(g->exec, st));
      p->last_launches += g->launches;
      done = true;
    }
  }
  if (!done && (rc = tc_middle(p, t, B, Tmax, h_len, d_len, st))) return rc;
  return tc_post(p, t, cur, wav_out, wav_dtype, B, Tmax, d_len, st, 2);
}

// ------------------------------------------------------------------------------ time split (P2P)
// BASELINE config 5 / SURVEY §8e: one long utterance split along time over R GPUs, one process per
// GPU.  Exchange granularity is the upsampling STAGE, not the layer: before stage i every rank
// needs HF[i] latent-frames worth of the previous stage's output beyond its own range (enough
// for the ConvTranspose skirt plus the 90-sample receptive field of the deepest AMP block), and
// gets them from its neighbours' exact rows by direct peer stores over NVLink (CUDA IPC mapped
// buffers) followed by a system-scope flag; inside a stage the window is decoded as a stand-alone
// utterance, whose window-end artefacts stay inside the halo rows and are never stored or sent.
// 6 exchanges per decode and ~6 % recomputed work at 8 x 7.5 s, against 39 % for whole-generator
// overlap-recompute (bvg_decode_shard) and 108 exchanges for per-layer halos.
constexpr int kShardMargin = 32;   // H: frames of margin on each side of the own range in every buffer

struct ShardState {
  bool ready = false;
  bvg_shard_geom g{};
  int Fs = 0, S = 0;
  int HF[kMaxStages + 2] = {};     // halo frames of stage i's input (i = 0..S-1) and of post (i = S)
  int HP = 0;                       // latent halo frames for conv_pre
  int h_win[kMaxStages + 2] = {};   // valid frames: [0] conv_pre window, [1+i] stage i window, [1+S] post window
  int* d_win = nullptr;
  int* flags = nullptr;             // [2 sides][16] epochs written by the neighbours; [32] = error flag
  void* nbr_ws[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // neighbour's ws[0], ws[1]
  int* nbr_flags[2] = {nullptr, nullptr};
  bool nbr_ipc[2] = {false, false};
  void* exported_ws[2] = {nullptr, nullptr};
  // error word of the decode in flight (k_halo_wait timeout), in mapped pinned memory: the host reads it without
  // touching the device, so every decode can check the previous one for free
  int* h_err = nullptr;
  int* d_err = nullptr;
};

static ShardState* shard_of(bvg_plan* p) {
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  return t ? static_cast<ShardState*>(t->shard) : nullptr;
}

__global__ void k_halo_send(const uint4* src_l, uint4* dst_l, const uint4* src_r, uint4* dst_r, int groups,
                            int n16, long long gstride16, int* flag_l, int* flag_r, int epoch) {
  const bool left = blockIdx.x == 0;
  const uint4* src = left ? src_l : src_r;
  uint4* dst = left ? dst_l : dst_r;
  int* flag = left ? flag_l : flag_r;
  if (!dst) return;
  for (long long i = threadIdx.x; i < (long long)groups * n16; i += blockDim.x) {
    const long long g = i / n16, r = i - g * n16;
    dst[g * gstride16 + r] = src[g * gstride16 + r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                       // halo rows are visible system-wide before the flag
    *reinterpret_cast<volatile int*>(flag) = epoch;
  }
}

__global__ void k_halo_wait(const int* flag_l, const int* flag_r, int epoch, int* err) {
  for (int side = 0; side < 2; ++side) {
    const volatile int* f = side == 0 ? flag_l : flag_r;
    if (!f) continue;
    long long spins = 0;
    while (*f < epoch) {
      __nanosleep(200);
      if (++spins > 20000000LL) {                              // ~4 s: give up instead of hanging the GPU
        *reinterpret_cast<volatile int*>(err) = 1 + side;
        __threadfence_system();
        return;
      }
    }
  }
  __threadfence_system();
}

// ws[0..1] are IPC-exported to the neighbours once a shard is set up: an ordinary decode must not re-allocate them
// under the peers' feet (bvg_api.cu::ensure_ws refuses to grow while this returns true)
bool tc_shard_pins_ws(const bvg_plan* p) {
  const TcPlan* t = static_cast<const TcPlan*>(p->tc);
  const ShardState* s = t ? static_cast<const ShardState*>(t->shard) : nullptr;
  return s && s->ready;
}

int tc_shard_setup(bvg_plan* p, const bvg_shard_geom* g, cudaStream_t st) {
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  if (!t) return fail(BVG_ERR_STATE, "shard setup: weights not loaded");
  if (!t->shard) t->shard = new ShardState();
  ShardState* s = static_cast<ShardState*>(t->shard);
  s->ready = false;                      // a new geometry may grow the workspace (and must then be re-exported)
  const int S = p->n_stages, own = g->f_end - g->f_begin;
  BVG_REQUIRE(g->f_begin >= 0 && own > 0 && g->f_end <= g->f_total, "bad shard range");
  // Back-to-back decodes need no host barrier between them: a rank can only finish decode n after its neighbours
  // have sent it their last halo of decode n, i.e. after every neighbour kernel that reads the stage buffer of the
  // OTHER parity has run; the first peer store of decode n+1 (end of phase 0) lands in stage buffer 1 while a lagging
  // neighbour is at most in its final phase S, which reads buffer S & 1 — disjoint for even S.
  BVG_REQUIRE(S % 2 == 0, "the time split needs an even number of upsampling stages (%d)", S);
  BVG_REQUIRE(g->own_max >= own, "own_max smaller than this shard");
  // receptive field of one stage at its output rate: deepest AMP block
  int kmax = 0, rows_out = 0;
  for (int j = 0; j < p->cfg.num_kernels; ++j) kmax = std::max(kmax, p->cfg.resblock_kernel_sizes[j]);
  for (int j = 0; j < p->cfg.num_kernels; ++j) {
    int r = 0;
    const int k = p->cfg.resblock_kernel_sizes[j];
    for (int m = 0; m < BVG_MAX_DIL; ++m) r += 5 + p->cfg.resblock_dilation_sizes[j][m] * (k - 1) / 2 + 5 + (k - 1) / 2;
    rows_out = std::max(rows_out, r);
  }
  for (int i = 0; i < S; ++i) {
    const int u = p->cfg.upsample_rates[i], kk = p->cfg.upsample_kernel_sizes[i];
    const int rows_in = (rows_out + u - 1) / u + kk / u + 1;
    s->HF[i] = (rows_in + p->rate[i] - 1) / p->rate[i];
  }
  s->HF[S] = (5 + 3 + p->rate[S] - 1) / p->rate[S] + 0;      // activation_post + conv_post
  if (s->HF[S] < 1) s->HF[S] = 1;
  s->HP = s->HF[0] + 4;                                        // conv_pre k=7 -> 3 frames + 1 spare
  BVG_REQUIRE(s->HP <= kShardMargin, "halo %d exceeds the buffer margin", s->HP);
  const bool hasl = g->f_begin > 0, hasr = g->f_end < g->f_total;
  BVG_REQUIRE(!hasl || g->own_left >= s->HP, "left neighbour shorter than the halo (%d frames)", s->HP);
  BVG_REQUIRE(!hasr || g->own_right >= s->HP, "right neighbour shorter than the halo (%d frames)", s->HP);
  BVG_REQUIRE(own >= s->HP, "shard shorter than the halo (%d frames)", s->HP);
  s->g = *g; s->S = S;
  s->Fs = g->own_max + 2 * kShardMargin;
  s->h_win[0] = (hasl ? s->HP : 0) + own + (hasr ? s->HP : 0);
  for (int i = 0; i <= S; ++i) s->h_win[1 + i] = (hasl ? s->HF[i] : 0) + own + (hasr ? s->HF[i] : 0);
  int rc;
  if ((rc = tc_prepare(p, t, 1, s->Fs))) return rc;
  if (!s->d_win) BVG_CUDA(cudaMalloc((void**)&s->d_win, sizeof(int) * (kMaxStages + 2)));
  if (!s->flags) {
    BVG_CUDA(cudaMalloc((void**)&s->flags, sizeof(int) * 64));
    BVG_CUDA(cudaMemsetAsync(s->flags, 0, sizeof(int) * 64, st));
  }
  if (!s->h_err) {
    BVG_CUDA(cudaHostAlloc((void**)&s->h_err, sizeof(int), cudaHostAllocMapped));
    *s->h_err = 0;
    BVG_CUDA(cudaHostGetDevicePointer((void**)&s->d_err, s->h_err, 0));
  }
  BVG_CUDA(cudaMemcpyAsync(s->d_win, s->h_win, sizeof(int) * (kMaxStages + 2), cudaMemcpyHostToDevice, st));
  BVG_CUDA(cudaStreamSynchronize(st));
  s->exported_ws[0] = p->ws[0];
  s->exported_ws[1] = p->ws[1];
  s->ready = true;
  return 0;
}

int tc_shard_halo_frames(bvg_plan* p) {
  ShardState* s = shard_of(p);
  return (s && s->ready) ? s->HP : -1;
}

int tc_shard_local_ptrs(bvg_plan* p, void** ws0, void** ws1, void** flags) {
  ShardState* s = shard_of(p);
  if (!s || !s->ready) return fail(BVG_ERR_STATE, "shard not set up");
  *ws0 = p->ws[0]; *ws1 = p->ws[1]; *flags = s->flags;
  return 0;
}

int tc_shard_export(bvg_plan* p, uint8_t* handles) {
  ShardState* s = shard_of(p);
  if (!s || !s->ready) return fail(BVG_ERR_STATE, "shard not set up");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  void* ptrs[3] = {p->ws[0], p->ws[1], s->flags};
  for (int i = 0; i < 3; ++i) {
    BVG_CUDA(cudaIpcGetMemHandle(&h, ptrs[i]));
    memcpy(handles + 64 * i, &h, 64);
  }
  return 0;
}

int tc_shard_connect(bvg_plan* p, int side, const uint8_t* handles, void* ws0, void* ws1, void* flags) {
  ShardState* s = shard_of(p);
  if (!s || !s->ready) return fail(BVG_ERR_STATE, "shard not set up");
  BVG_REQUIRE(side == 0 || side == 1, "side must be 0 (left) or 1 (right)");
  if (handles) {
    void* out[3];
    for (int i = 0; i < 3; ++i) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * i, 64);
      BVG_CUDA(cudaIpcOpenMemHandle(&out[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    ws0 = out[0]; ws1 = out[1]; flags = out[2];
    s->nbr_ipc[side] = true;
  }
  s->nbr_ws[side][0] = ws0; s->nbr_ws[side][1] = ws1; s->nbr_flags[side] = (int*)flags;
  return 0;
}

void tc_shard_free(TcPlan* t) {
  ShardState* s = static_cast<ShardState*>(t->shard);
  if (!s) return;
  for (int side = 0; side < 2; ++side)
    if (s->nbr_ipc[side]) {
      cudaIpcCloseMemHandle(s->nbr_ws[side][0]);
      cudaIpcCloseMemHandle(s->nbr_ws[side][1]);
      cudaIpcCloseMemHandle(s->nbr_flags[side]);
    }
  if (s->d_win) cudaFree(s->d_win);
  if (s->flags) cudaFree(s->flags);
  if (s->h_err) cudaFreeHost(s->h_err);
  delete s;
  t->shard = nullptr;
}

// phase 0: conv_pre + stage 0 (+ send halos for stage 1); phase i in 1..S-1: [wait] stage i (+ send);
// phase S: [wait] activation_post + conv_post + tanh, own samples -> wav_out.
int tc_shard_run(bvg_plan* p, int phase, const void* latent, int latent_dtype, const float* spk_emb, void* wav_out,
                 int wav_dtype, int epoch, int wait, cudaStream_t st) {
  TcPlan* t = static_cast<TcPlan*>(p->tc);
  ShardState* s = shard_of(p);
  if (!s || !s->ready) return fail(BVG_ERR_STATE, "shard not set up");
  if (p->ws[0] != s->exported_ws[0] || p->ws[1] != s->exported_ws[1])
    return fail(BVG_ERR_STATE, "workspace was re-allocated after bvg_shard_setup; set the shard up again");
  if (*reinterpret_cast<volatile int*>(s->h_err))
    return fail(BVG_ERR_STATE, "time-split decode: an earlier phase timed out waiting for the %s neighbour's halo",
                *s->h_err == 1 ? "left" : "right");
  const int S = s->S, H = kShardMargin, Fs = s->Fs;
  BVG_REQUIRE(phase >= 0 && phase <= S, "phase %d outside [0,%d]", phase, S);
  const int own = s->g.f_end - s->g.f_begin;
  const bool hasl = s->g.f_begin > 0, hasr = s->g.f_end < s->g.f_total;
  __nv_bfloat16* buf[2] = {(__nv_bfloat16*)p->ws[0], (__nv_bfloat16*)p->ws[1]};
  int rc;
  p->cur_sum_frames = own;
  if (phase == 0) {
    BVG_REQUIRE(latent && spk_emb, "phase 0 needs the latent window and the speaker embedding");
    p->last_launches = 0;
    if ((rc = compute_cond_bias(p, spk_emb, 1, st))) return rc;
    const int hlp = hasl ? s->HP : 0;
    if ((rc = tc_pre(p, t, latent, latent_dtype, buf[0] + (size_t)(H - hlp) * 8, 1, Fs, &s->h_win[0], s->d_win, st)))
      return rc;
  } else if (wait && (hasl || hasr)) {
    k_halo_wait<<<1, 1, 0, st>>>(hasl ? s->flags + phase : nullptr, hasr ? s->flags + 16 + phase : nullptr, epoch,
                                 s->d_err);
    BVG_CUDA(cudaGetLastError());
    ++p->last_launches;
  }
  if (phase < S) {
    const int i = phase;
    const int hl = hasl ? s->HF[i] : 0;
    StageIO io;
    io.cur = buf[i & 1] + (size_t)(H - hl) * p->rate[i] * 8;
    io.xin = (__nv_bfloat16*)p->ws[2];
    io.xs = buf[(i + 1) & 1] + (size_t)(H - hl) * p->rate[i + 1] * 8;
    io.B = 1; io.Fs = Fs; io.h_len = &s->h_win[1 + i]; io.d_len = s->d_win + 1 + i;
    // never touch the halo rows the neighbours write; at a true sequence end keep the store range
    // open so the zero rows past the end (conv padding for the next ConvTranspose1d) get written
    io.st_lo_f = hl; io.st_hi_f = hasr ? hl + own : 0x7fffff;
    io.cur_rows = (Fs - (H - hl)) * p->rate[i];
    io.xs_rows = (Fs - (H - hl)) * p->rate[i + 1];
    io.concurrent = true;
    if ((rc = tc_stage(p, t, i, io, st))) return rc;
    // send my boundary rows of this stage's output to the neighbours (input halo of the next phase)
    if (hasl || hasr) {
      const int hn = s->HF[i + 1], rho = p->rate[i + 1], groups = p->C[i + 1] / 8;
      const long long gstride16 = (long long)Fs * rho;       // uint4 (one 16 B row) per channel group
      const uint4* mine = reinterpret_cast<const uint4*>(buf[(i + 1) & 1]);
      const uint4* src_l = mine + (long long)H * rho;                          // my first hn frames
      const uint4* src_r = mine + (long long)(H + own - hn) * rho;             // my last hn frames
      uint4 *dst_l = nullptr, *dst_r = nullptr;
      int *fl = nullptr, *fr = nullptr;
      if (hasl) {
        if (!s->nbr_ws[0][0]) return fail(BVG_ERR_STATE, "left neighbour not connected");
        dst_l = reinterpret_cast<uint4*>(s->nbr_ws[0][(i + 1) & 1]) + (long long)(H + s->g.own_left) * rho;
        fl = s->nbr_flags[0] + 16 + (i + 1);                 // I am its RIGHT neighbour
      }
      if (hasr) {
        if (!s->nbr_ws[1][0]) return fail(BVG_ERR_STATE, "right neighbour not connected");
        dst_r = reinterpret_cast<uint4*>(s->nbr_ws[1][(i + 1) & 1]) + (long long)(H - hn) * rho;
        fr = s->nbr_flags[1] + (i + 1);                      // I am its LEFT neighbour
      }
      k_halo_send<<<2, 1024, 0, st>>>(src_l, dst_l, src_r, dst_r, groups, hn * rho, gstride16, fl, fr, epoch);
      BVG_CUDA(cudaGetLastError());
      ++p->last_launches;
    }
    return 0;
  }
  // post
  BVG_REQUIRE(wav_out, "final phase needs wav_out");
  const int hl = hasl ? s->HF[S] : 0;
  const size_t esz = wav_dtype == BVG_F32 ? 4 : 2;
  void* tmp = p->ws[3];
  if ((rc = tc_post(p, t, buf[S & 1] + (size_t)(H - hl) * p->rate[S] * 8, tmp, wav_dtype, 1, Fs, s->d_win + 1 + S, st)))
    return rc;
  BVG_CUDA(cudaMemcpyAsync(wav_out, (const char*)tmp + (size_t)hl * p->up_total * esz, (size_t)own * p->up_total * esz,
                           cudaMemcpyDeviceToDevice, st));
  return 0;
}

int tc_shard_error(bvg_plan* p, int clear) {
  ShardState* s = shard_of(p);
  if (!s || !s->ready) return -1;
  const int e = *reinterpret_cast<volatile int*>(s->h_err);   // mapped pinned word: no device access
  if (clear) *s->h_err = 0;
  return e;
}

// ------------------------------------------------------------------------------ per-op (tests)
int tc_amp_layer(const float* x, float* y, const float* resid, int B, int C_in, int C_out, int T, const float* w,
                 const float* bias, int k, int dilation, int act, const float* up_filter,
                 const float* down_filter, const float* alpha, const float* beta, int logscale, cudaStream_t st) {
  BVG_REQUIRE(C_in % 8 == 0 && C_out % 8 == 0, "tcgen05 per-op path: channels must be multiples of 8");
  const int hc = dilation * (k - 1) / 2;
  if (hc > 25) return fail(BVG_ERR_UNSUPPORTED, "tcgen05 path supports conv halos up to 25 samples");
  std::vector<void*> tmp;
  auto cleanup = [&]() { for (void* q : tmp) cudaFree(q); };
  auto alloc = [&](void** ptr, size_t bytes) { cudaError_t e = cudaMalloc(ptr, bytes); if (e == cudaSuccess) tmp.push_back(*ptr); return e; };
  ConvW cw; ActW aw;
  cw.Cin = C_in; cw.Cout = C_out; cw.K = k;
  const size_t nw = (size_t)C_out * C_in * k;
  __nv_bfloat16 *xb, *yb, *rb = nullptr, *zb = nullptr;
  float *a_dev = nullptr, *invb_dev = nullptr;
  if (alloc((void**)&cw.wp, nw * 4) || alloc((void**)&xb, (size_t)B * C_in * T * 2) ||
      alloc((void**)&yb, (size_t)B * C_out * T * 2) || alloc((void**)&a_dev, C_in * 4) || alloc((void**)&invb_dev, C_in * 4) ||
      (resid && alloc((void**)&rb, (size_t)B * C_out * T * 2)) ||
      (act && split_layer(C_in) && alloc((void**)&zb, (size_t)B * C_in * T * 2))) {
    cleanup();
    return fail(BVG_ERR_CUDA, "tc_amp_layer: allocation failed");
  }
  cw.bias = const_cast<float*>(bias);
  cudaMemsetAsync(yb, 0, (size_t)B * C_out * T * 2, st);
  tc_pack_conv_w(w, cw.wp, C_out, C_in, k, st);
  if (act) {
    tc_snake_params(alpha, beta, a_dev, invb_dev, C_in, logscale, st);
    aw.a = a_dev; aw.invb = invb_dev; aw.C = C_in;
    cudaMemcpyAsync(aw.up, up_filter, 48, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(aw.dn, down_filter, 48, cudaMemcpyDeviceToHost, st);
  }
  TcLayer L;
  std::vector<void*> owned;
  int rc = build_layer(owned, L, cw, act ? &aw : nullptr, st);
  for (void* q : owned) tmp.push_back(q);
  if (rc) { cleanup(); return rc; }
  dim3 g(ceil_div(T, 128), C_in / 8, B);
  k_cm_to_blk<<<g, 128, 0, st>>>(x, xb, C_in, T);
  if (resid) {
    dim3 g2(ceil_div(T, 128), C_out / 8, B);
    k_cm_to_blk<<<g2, 128, 0, st>>>(resid, rb, C_out, T);
  }
  cudaStreamSynchronize(st);   // filter taps on the host
  CUtensorMap map;
  if ((rc = make_map(xb, C_in, T, B, &map))) { cleanup(); return rc; }
  TcLaunch q;
  q.x = xb; q.resid = rb; q.out = yb; q.dil = dilation; q.B = B; q.Tstride = T; q.rate = 1; q.d_len = nullptr;
  q.zbuf = zb;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    q.sm_count = sms;
  }
  rc = launch_tc(nullptr, map, L, cw, act ? &aw : nullptr, q, st);
  if (!rc) {
    dim3 g3(ceil_div(T, 128), C_out / 8, B);
    k_blk_to_cm<<<g3, 128, 0, st>>>(yb, y, C_out, T);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(BVG_ERR_CUDA, "tc_amp_layer: %s", cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}

// ConvTranspose1d(C_in, C_out, k, stride u, padding (k-u)/2) through the SAME tcgen05 launch as the decode
// (models.py:157-163,232-233): (k/u)-tap implicit GEMM over the input rate, phase-scatter epilogue.  fp32 [B,C,T] in/out,
// operands rounded to bf16.  Test entry point: packs per call and synchronises.
int tc_conv_transpose(const float* x, float* y, int B, int C_in, int C_out, int T, const float* w, const float* bias,
                      int k, int u, cudaStream_t st) {
  BVG_REQUIRE(C_in % 8 == 0 && C_out % 8 == 0, "tcgen05 per-op path: channels must be multiples of 8");
  BVG_REQUIRE(k / u - 1 <= X_LEAD, "tcgen05 ConvTranspose1d: at most %d taps per phase", X_LEAD + 1);
  std::vector<void*> tmp;
  auto cleanup = [&]() { for (void* q : tmp) cudaFree(q); };
  auto alloc = [&](void** ptr, size_t bytes) { cudaError_t e = cudaMalloc(ptr, bytes); if (e == cudaSuccess) tmp.push_back(*ptr); return e; };
  ConvW cw;
  cw.Cin = C_in; cw.Cout = C_out; cw.K = k;
  const size_t nw = (size_t)C_out * C_in * k;
  __nv_bfloat16 *xb, *yb;
  if (alloc((void**)&cw.wp, nw * 4) || alloc((void**)&xb, (size_t)B * C_in * T * 2) ||
      alloc((void**)&yb, (size_t)B * C_out * T * u * 2)) {
    cleanup();
    return fail(BVG_ERR_CUDA, "tc_conv_transpose: allocation failed");
  }
  tc_pack_convtr_w(w, cw.wp, C_in, C_out, k, st);
  TcLayer L;
  std::vector<void*> owned;
  int rc = build_layer_tr(owned, L, cw, u, st);
  for (void* q : owned) tmp.push_back(q);
  if (rc) { cleanup(); return rc; }
  dim3 g(ceil_div(T, 128), C_in / 8, B);
  k_cm_to_blk<<<g, 128, 0, st>>>(x, xb, C_in, T);
  CUtensorMap map;
  if ((rc = make_map(xb, C_in, T, B, &map))) { cleanup(); return rc; }
  ConvW cg = cw;                       // the GEMM's view: u*C_out phase-major columns, k/u taps
  cg.Cout = u * C_out; cg.K = k / u; cg.bias = const_cast<float*>(bias);
  TcLaunch q;
  q.x = xb; q.out = yb; q.dil = 1; q.B = B; q.Tstride = T; q.out_tstride = T * u; q.rate = 1; q.d_len = nullptr;
  q.up = u; q.pad = (k - u) / 2; q.cphase = C_out; q.cls = 2;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    q.sm_count = sms;
  }
  rc = launch_tc(nullptr, map, L, cg, nullptr, q, st);
  if (!rc) {
    dim3 g3(ceil_div(T * u, 128), C_out / 8, B);
    k_blk_to_cm<<<g3, 128, 0, st>>>(yb, y, C_out, T * u);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(BVG_ERR_CUDA, "tc_conv_transpose: %s", cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}

}  // namespace bvg
