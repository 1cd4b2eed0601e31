// bvg_api.cu — C ABI of libbvg.so (include/bvg.h): plan lifetime, weight packing, decode
// orchestration of the fp32 SIMT path, per-op entry points.  The bf16 tcgen05 path lives in
// decode_tc.cu and is dispatched from here.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <type_traits>

#include "kernels_f32.cuh"
#include "plan.h"
#include "tc_api.h"

namespace bvg {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return status;
}

// PDL on the decode's launches: default on; BVG_PDL=0 or bvg_set_pdl(0) turns the launch attribute off
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load();
  if (v < 0) {
    const char* e = getenv("BVG_PDL");
    v = (!e || atoi(e) != 0) ? 1 : 0;
    g_pdl = v;
  }
  return v != 0;
}
int set_pdl(int on) {
  const int old = pdl_enabled() ? 1 : 0;
  g_pdl = on ? 1 : 0;
  return old;
}

cudaError_t func_attr_once(const void* fn, cudaFuncAttribute attr, int value) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, int>, int> done;   // (function, attribute, device) -> value set
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple(fn, (int)attr, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second == value) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, attr, value);
  if (e == cudaSuccess) done[key] = value;
  return e;
}

static int dev_alloc(bvg_plan* p, void** out, size_t bytes) {
  BVG_CUDA(cudaMalloc(out, bytes ? bytes : 4));
  p->owned.push_back(*out);
  return 0;
}

// ------------------------------------------------------------------------------ launches
struct Ctx {           // launch context: where to count launches / record profile events
  bvg_plan* p;
  cudaStream_t st;
  int cls;
  double elt;          // bytes per activation element on this path
};

static void conv_cost(const Ctx& c, const ConvArgs& a, int K, double* flops, double* bytes) {
  const double samples = (c.p ? c.p->cur_sum_frames : 0.0) * a.rate;
  *flops = 2.0 * a.Cin * a.Cout * K * samples;
  *bytes = samples * c.elt * (a.Cin + a.Cout + (a.resid ? a.Cout : 0) + (a.acc_in ? a.Cout : 0)) +
           (double)a.Cin * a.Cout * K * c.elt;
}

template <int K, bool ACT, int TY, int NC, int NT, int XL, bool FAST = false>
static int launch_conv(const ConvArgs& a, int B, const Ctx& c) {
  cudaStream_t st = c.st;
  constexpr int TX = 256 / TY, TT = TX * NT, COB = TY * NC, CK = 8;
  const int hc = a.dil * (K - 1) / 2;
  const int ZW = TT + 2 * hc, SW = 2 * ZW + 12, XW = ZW + 12;
  const size_t smem = sizeof(float) * (size_t)(CK * ZW + CK * K * COB + (ACT ? CK * (SW + XW) : 0));
  auto kern = k_conv_f32<K, ACT, TY, NC, NT, XL, FAST>;
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  dim3 grid(ceil_div(a.Tmax, TT), ceil_div(a.Cout, COB), B);
  double fl, by;
  conv_cost(c, a, K, &fl, &by);
  prof_begin(c.p, st, c.cls, fl, by);
  kern<<<grid, 256, smem, st>>>(a);
  prof_end(c.p, st);
  BVG_CUDA(cudaGetLastError());
  if (c.p) ++c.p->last_launches;
  return 0;
}

template <bool ACT>
static int launch_conv_k(int K, const ConvArgs& a, int B, const Ctx& c) {
  const bool narrow = a.Cout <= 32;
  switch (K) {
    case 3:
      return narrow ? launch_conv<3, ACT, 8, 4, 8, 0>(a, B, c)
                    : launch_conv<3, ACT, 16, 4, 8, 0>(a, B, c);
    case 7:
      return narrow ? launch_conv<7, ACT, 8, 4, 8, 0>(a, B, c)
                    : launch_conv<7, ACT, 16, 4, 8, 0>(a, B, c);
    case 11:
      return narrow ? launch_conv<11, ACT, 8, 4, 8, 0>(a, B, c)
                    : launch_conv<11, ACT, 16, 4, 8, 0>(a, B, c);
    default:
      return fail(BVG_ERR_UNSUPPORTED, "conv kernel size %d not supported (3, 7, 11)", K);
  }
}

template <bool BLK>
static int launch_convtr(const ConvTrArgs& a, int B, const Ctx& c) {
  cudaStream_t st = c.st;
  constexpr int TY = 16, NC = 4, NT = 8, TX = 256 / TY, TT = TX * NT, COB = TY * NC, CK = 8;
  const int QW = TT / a.U + a.KK / a.U + 2;
  const size_t smem = sizeof(float) * (size_t)(CK * QW + CK * a.KK * COB);
  auto kern = k_convtr_f32<TY, NC, NT, BLK>;
  BVG_CUDA(func_attr_once((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  dim3 grid(ceil_div(a.Tmax_out, TT), ceil_div(a.Cout, COB), B);
  const double in_samples = (c.p ? c.p->cur_sum_frames : 0.0) * a.rate_out / a.U;
  prof_begin(c.p, st, c.cls, 2.0 * a.Cin * a.Cout * a.KK * in_samples,
             in_samples * c.elt * (a.Cin + (double)a.Cout * a.U) + (double)a.Cin * a.Cout * a.KK * c.elt);
  kern<<<grid, 256, smem, st>>>(a);
  prof_end(c.p, st);
  BVG_CUDA(cudaGetLastError());
  if (c.p) ++c.p->last_launches;
  return 0;
}

static void fill_act(ActParams& ap, const ActW& w) {
  ap.a = w.a;
  ap.invb = w.invb;
  memcpy(ap.up, w.up, sizeof(ap.up));
  memcpy(ap.dn, w.dn, sizeof(ap.dn));
}

// ------------------------------------------------------------------------------ weights
struct TensorMap {
  std::map<std::string, const bvg_tensor_desc*> m;
  const bvg_tensor_desc* get(const std::string& n) const {
    auto it = m.find(n);
    return it == m.end() ? nullptr : it->second;
  }
};

static int need(const TensorMap& tm, const std::string& name, std::initializer_list<int64_t> shape,
                const float** out) {
  const bvg_tensor_desc* d = tm.get(name);
  if (!d) return fail(BVG_ERR_ARG, "missing tensor '%s'", name.c_str());
  if (d->dtype != BVG_F32) return fail(BVG_ERR_ARG, "tensor '%s' must be fp32", name.c_str());
  // full shape, not just the element count: a ConvTranspose1d weight handed over as [Cout,Cin,K] instead of
  // [Cin,Cout,K] would otherwise load silently and decode garbage.  Leading / trailing unit dims are tolerated
  // (filters are [1,1,12] in the state dict, alpha / beta are [C]).
  auto squeeze = [](const int64_t* s, int n, int64_t* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) if (s[i] != 1) out[m++] = s[i];
    return m;
  };
  int64_t hs[8], wsq[8], wraw[8];
  int nw = 0;
  for (int64_t v : shape) if (nw < 8) wraw[nw++] = v;
  if (d->ndim < 0 || d->ndim > 4) return fail(BVG_ERR_ARG, "tensor '%s': ndim %d", name.c_str(), d->ndim);
  const int nh = squeeze(d->shape, d->ndim, hs), nq = squeeze(wraw, nw, wsq);
  bool same = nh == nq;
  for (int i = 0; same && i < nh; ++i) same = hs[i] == wsq[i];
  if (!same) {
    std::string have, want;
    for (int i = 0; i < d->ndim; ++i) have += (i ? "," : "") + std::to_string((long long)d->shape[i]);
    for (int i = 0; i < nw; ++i) want += (i ? "," : "") + std::to_string((long long)wraw[i]);
    return fail(BVG_ERR_ARG, "tensor '%s' has shape [%s], expected [%s]", name.c_str(), have.c_str(), want.c_str());
  }
  *out = reinterpret_cast<const float*>(d->data);
  return 0;
}

static int load_conv(bvg_plan* p, const TensorMap& tm, const std::string& prefix, ConvW& cw,
                     int Cout, int Cin, int K, bool transposed, cudaStream_t st) {
  const float *w, *b;
  int rc;
  if (transposed) {
    if ((rc = need(tm, prefix + ".weight", {Cin, Cout, K}, &w))) return rc;
  } else {
    if ((rc = need(tm, prefix + ".weight", {Cout, Cin, K}, &w))) return rc;
  }
  if ((rc = need(tm, prefix + ".bias", {Cout}, &b))) return rc;
  cw.Cin = Cin; cw.Cout = Cout; cw.K = K;
  const size_t n = (size_t)Cout * Cin * K;
  if (!cw.wp) {
    if ((rc = dev_alloc(p, (void**)&cw.wp, n * sizeof(float)))) return rc;
    if ((rc = dev_alloc(p, (void**)&cw.bias, Cout * sizeof(float)))) return rc;
  }
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 4096);
  if (transposed) k_pack_convtr_w<<<blocks, 256, 0, st>>>(w, cw.wp, Cin, Cout, K);
  else k_pack_conv_w<<<blocks, 256, 0, st>>>(w, cw.wp, Cout, Cin, K);
  BVG_CUDA(cudaGetLastError());
  BVG_CUDA(cudaMemcpyAsync(cw.bias, b, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int load_act(bvg_plan* p, const TensorMap& tm, const std::string& prefix, ActW& aw, int C,
                    cudaStream_t st) {
  const float *al, *be, *uf, *df;
  int rc;
  if ((rc = need(tm, prefix + ".act.alpha", {C}, &al))) return rc;
  if ((rc = need(tm, prefix + ".act.beta", {C}, &be))) return rc;
  if ((rc = need(tm, prefix + ".upsample.filter", {12}, &uf))) return rc;
  if ((rc = need(tm, prefix + ".downsample.lowpass.filter", {12}, &df))) return rc;
  aw.C = C;
  if (!aw.a) {
    if ((rc = dev_alloc(p, (void**)&aw.a, C * sizeof(float)))) return rc;
    if ((rc = dev_alloc(p, (void**)&aw.invb, C * sizeof(float)))) return rc;
  }
  k_snake_params<<<ceil_div(C, 128), 128, 0, st>>>(al, be, aw.a, aw.invb, C, p->cfg.snake_logscale);
  BVG_CUDA(cudaGetLastError());
  BVG_CUDA(cudaMemcpyAsync(aw.up, uf, 12 * sizeof(float), cudaMemcpyDeviceToHost, st));
  BVG_CUDA(cudaMemcpyAsync(aw.dn, df, 12 * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

// ------------------------------------------------------------------------------ workspace
static int ensure_ws(bvg_plan* p, size_t bytes_per_buf) {
  if (bytes_per_buf <= p->ws_bytes) return 0;
  if (tc_shard_pins_ws(p))
    return fail(BVG_ERR_STATE, "this decode needs a larger workspace (%zu B per buffer, have %zu), but the workspace is "
                "exported to the neighbouring GPUs of a time-split decode: call bvg_shard_setup again afterwards, or use "
                "a separate plan", bytes_per_buf, p->ws_bytes);
  BVG_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < 4; ++i) {
    if (p->ws[i]) BVG_CUDA(cudaFree(p->ws[i]));
    p->ws[i] = nullptr;
  }
  p->ws_bytes = 0;
  // Zero-filled once: rows past an utterance's end are only ever written where a consumer's zero padding needs them,
  // and the banded-Toeplitz FIR MMAs (amp_nar.cuh) multiply stale rows by zero taps — which must not be NaN / Inf bit
  // patterns of uninitialised memory.  Everything the kernels write afterwards is finite.
  for (int i = 0; i < 4; ++i) {
    BVG_CUDA(cudaMalloc(&p->ws[i], bytes_per_buf));
    BVG_CUDA(cudaMemset(p->ws[i], 0, bytes_per_buf));
  }
  p->ws_bytes = bytes_per_buf;
  ++p->alloc_gen;
  return 0;
}

int upload_lengths(bvg_plan* p, const int32_t* lengths, int B, int Tmax, cudaStream_t st,
                   const int** d_out) {
  if (B > p->d_len_cap) {
    BVG_CUDA(cudaDeviceSynchronize());
    if (p->d_len) BVG_CUDA(cudaFree(p->d_len));
    for (int i = 0; i < kLenSlots; ++i) {
      if (p->h_len[i]) BVG_CUDA(cudaFreeHost(p->h_len[i]));
      BVG_CUDA(cudaMallocHost((void**)&p->h_len[i], sizeof(int) * B));
      if (!p->len_ev[i]) BVG_CUDA(cudaEventCreateWithFlags(&p->len_ev[i], cudaEventDisableTiming));
    }
    BVG_CUDA(cudaMalloc((void**)&p->d_len, sizeof(int) * B));
    p->d_len_cap = B;
    ++p->alloc_gen;
  }
  const int s = p->len_slot;
  p->len_slot = (s + 1) % kLenSlots;
  BVG_CUDA(cudaEventSynchronize(p->len_ev[s]));  // slot's previous copy has been consumed
  for (int b = 0; b < B; ++b) {
    const int L = lengths ? lengths[b] : Tmax;
    if (L < 1 || L > Tmax) return fail(BVG_ERR_ARG, "lengths[%d]=%d outside [1,%d]", b, L, Tmax);
    p->h_len[s][b] = L;
  }
  BVG_CUDA(cudaMemcpyAsync(p->d_len, p->h_len[s], sizeof(int) * B, cudaMemcpyHostToDevice, st));
  BVG_CUDA(cudaEventRecord(p->len_ev[s], st));
  *d_out = p->d_len;
  return 0;
}

int compute_cond_bias(bvg_plan* p, const float* spk_emb, int B, cudaStream_t st) {
  const size_t need_elems = (size_t)B * p->cond_total;
  if (need_elems > p->condb_elems) {
    BVG_CUDA(cudaDeviceSynchronize());
    if (p->condb) BVG_CUDA(cudaFree(p->condb));
    BVG_CUDA(cudaMalloc((void**)&p->condb, need_elems * sizeof(float)));
    p->condb_elems = need_elems;
    ++p->alloc_gen;
  }
  const int warps = B * p->cond_total;
  BVG_CUDA(launch_k(k_cond_bias, dim3(ceil_div(warps * 32, 256)), dim3(256), 0, st, false, (const float*)p->cond_W,
                    (const float*)p->cond_b, spk_emb, p->condb, p->cond_total, p->cfg.speaker_embedding_dim, B));
  ++p->last_launches;
  return 0;
}

// ------------------------------------------------------------------------------ fp32 decode
static int decode_f32(bvg_plan* p, const void* latent, int latent_dtype, const int* d_len, int B,
                      int Tmax, void* wav_out, int wav_dtype, cudaStream_t st) {
  int rc;
  size_t max_elems = (size_t)p->C[0] * Tmax;
  for (int i = 0; i < p->n_stages; ++i)
    max_elems = std::max(max_elems, (size_t)p->C[i + 1] * Tmax * p->rate[i + 1]);
  if ((rc = ensure_ws(p, max_elems * B * sizeof(float)))) return rc;
  float* bufs[4] = {(float*)p->ws[0], (float*)p->ws[1], (float*)p->ws[2], (float*)p->ws[3]};
  Ctx cx{p, st, 2, 4.0};

  // conv_pre + cond_layer add (models.py:226-228), reading the time-major latent directly
  float* cur = bufs[0];
  {
    ConvArgs a{};
    a.x = latent; a.x_dtype = latent_dtype; a.x_tstride = 0;
    a.wp = p->conv_pre.wp; a.bias = p->conv_pre.bias;
    a.bias_b = p->condb + p->cond_off[0]; a.bias_b_stride = p->cond_total;
    a.out = cur; a.out_dtype = BVG_F32; a.out_tstride = Tmax; a.div = 1.f;
    a.Cin = p->conv_pre.Cin; a.Cout = p->conv_pre.Cout; a.dil = 1;
    a.lengths = d_len; a.rate = 1; a.Tmax = Tmax;
    if ((rc = launch_conv<7, false, 16, 4, 8, 1>(a, B, cx))) return rc;
  }
  const int nk = p->cfg.num_kernels;
  for (int i = 0; i < p->n_stages; ++i) {
    const int Ci = p->C[i + 1], Ri = p->rate[i + 1], Ti = Tmax * Ri;
    // pick 4 distinct buffers: cur (input), xin, xr, xt; xs reuses cur's buffer once xin exists
    float* free3[3];
    int nf = 0;
    for (int q = 0; q < 4; ++q)
      if (bufs[q] != cur) free3[nf++] = bufs[q];
    float *xin = free3[0], *xr = free3[1], *xt = free3[2], *xs = cur;
    {
      ConvTrArgs a{};
      a.x = cur; a.x_tstride = Tmax * p->rate[i];
      a.wp = p->ups[i].wp; a.bias = p->ups[i].bias;
      a.bias_b = p->cfg.cond_in_each_up_layer ? p->condb + p->cond_off[i + 1] : nullptr;
      a.bias_b_stride = p->cond_total;
      a.out = xin; a.out_tstride = Ti;
      a.Cin = p->ups[i].Cin; a.Cout = Ci; a.KK = p->ups[i].K; a.U = p->cfg.upsample_rates[i];
      a.lengths = d_len; a.rate_out = Ri; a.Tmax_out = Ti;
      cx.cls = 2;
      if ((rc = launch_convtr<false>(a, B, cx))) return rc;
    }
    cx.cls = (Ci >= 192) ? 0 : 1;
    for (int j = 0; j < nk; ++j) {
      const int n = i * nk + j;
      const int K = p->cfg.resblock_kernel_sizes[j];
      const float* xcur = xin;
      for (int m = 0; m < BVG_MAX_DIL; ++m) {
        const int d = p->cfg.resblock_dilation_sizes[j][m];
        ConvArgs a{};
        a.x = xcur; a.x_tstride = Ti;
        a.wp = p->rb1[n][m].wp; a.bias = p->rb1[n][m].bias;
        a.out = xt; a.out_dtype = BVG_F32; a.out_tstride = Ti; a.div = 1.f;
        a.Cin = Ci; a.Cout = Ci; a.dil = d;
        a.lengths = d_len; a.rate = Ri; a.Tmax = Ti;
        fill_act(a.act, p->rba[n][2 * m]);
        if ((rc = launch_conv_k<true>(K, a, B, cx))) return rc;

        const bool last = (m == BVG_MAX_DIL - 1);
        ConvArgs c{};
        c.x = xt; c.x_tstride = Ti;
        c.wp = p->rb2[n][m].wp; c.bias = p->rb2[n][m].bias;
        c.resid = xcur;
        c.out_dtype = BVG_F32; c.out_tstride = Ti; c.div = 1.f;
        c.Cin = Ci; c.Cout = Ci; c.dil = 1;
        c.lengths = d_len; c.rate = Ri; c.Tmax = Ti;
        fill_act(c.act, p->rba[n][2 * m + 1]);
        if (!last) {
          c.out = xr;
        } else {
          c.out = xs;
          c.acc_in = (j > 0) ? xs : nullptr;
          c.div = (j == nk - 1) ? (float)nk : 1.f;
        }
        if ((rc = launch_conv_k<true>(K, c, B, cx))) return rc;
        xcur = xr;
      }
    }
    cur = xs;
  }
  // activation_post + conv_post + tanh (models.py:248-250)
  {
    const int S = p->n_stages;
    ConvArgs a{};
    a.x = cur; a.x_tstride = Tmax * p->rate[S];
    a.wp = p->conv_post.wp; a.bias = p->conv_post.bias;
    a.out = wav_out; a.out_dtype = wav_dtype; a.out_tstride = Tmax * p->rate[S]; a.div = 1.f;
    a.Cin = p->C[S]; a.Cout = 1; a.dil = 1;
    a.lengths = d_len; a.rate = p->rate[S]; a.Tmax = Tmax * p->rate[S];
    a.tanh_out = 1; a.zero_tail = 1;
    fill_act(a.act, p->act_post);
    cx.cls = 3;
    if ((rc = launch_conv<7, true, 1, 1, 2, 0>(a, B, cx))) return rc;
  }
  return 0;
}

// ---- SIMT kernels reused by the tcgen05 path (blocked bf16 activations) ----------------------
int tc_ensure_ws(bvg_plan* p, size_t bytes_per_buf) { return ensure_ws(p, bytes_per_buf); }

// ConvTranspose1d + cond add of stage i (models.py:232-236), blocked bf16 in/out
int simt_convtr_blk(bvg_plan* p, const void* x_blk, void* out_blk, int i, int B, int Tmax, const int* d_len,
                    cudaStream_t st) {
  Ctx cx{p, st, 2, 2.0};
  ConvTrArgs a{};
  a.x = x_blk; a.x_tstride = Tmax * p->rate[i];
  a.wp = p->ups[i].wp; a.bias = p->ups[i].bias;
  a.bias_b = p->cfg.cond_in_each_up_layer ? p->condb + p->cond_off[i + 1] : nullptr;
  a.bias_b_stride = p->cond_total;
  a.out = out_blk; a.out_tstride = Tmax * p->rate[i + 1];
  a.Cin = p->ups[i].Cin; a.Cout = p->C[i + 1]; a.KK = p->ups[i].K; a.U = p->cfg.upsample_rates[i];
  a.lengths = d_len; a.rate_out = p->rate[i + 1]; a.Tmax_out = Tmax * p->rate[i + 1];
  return launch_convtr<true>(a, B, cx);
}

// activation_post + conv_post + tanh (models.py:248-250) from the blocked bf16 stage output
int simt_post_blk(bvg_plan* p, const void* x_blk, void* wav, int wav_dtype, int B, int Tmax, const int* d_len,
                  cudaStream_t st) {
  const int S = p->n_stages;
  Ctx cx{p, st, 3, 2.0};
  ConvArgs a{};
  a.x = x_blk; a.x_tstride = Tmax * p->rate[S];
  a.wp = p->conv_post.wp; a.bias = p->conv_post.bias;
  a.out = wav; a.out_dtype = wav_dtype; a.out_tstride = Tmax * p->rate[S]; a.div = 1.f;
  a.Cin = p->C[S]; a.Cout = 1; a.dil = 1;
  a.lengths = d_len; a.rate = p->rate[S]; a.Tmax = Tmax * p->rate[S];
  a.tanh_out = 1; a.zero_tail = 1;
  fill_act(a.act, p->act_post);
  return launch_conv<7, true, 1, 1, 2, 2, true>(a, B, cx);
}

void tc_pack_conv_w(const float* w, float* wp, int Cout, int Cin, int K, cudaStream_t st) {
  const size_t n = (size_t)Cout * Cin * K;
  k_pack_conv_w<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, st>>>(w, wp, Cout, Cin, K);
}
void tc_pack_convtr_w(const float* w, float* wp, int Cin, int Cout, int K, cudaStream_t st) {
  const size_t n = (size_t)Cout * Cin * K;
  k_pack_convtr_w<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, st>>>(w, wp, Cin, Cout, K);
}
void tc_snake_params(const float* alpha, const float* beta, float* a, float* invb, int C, int logscale,
                     cudaStream_t st) {
  k_snake_params<<<ceil_div(C, 128), 128, 0, st>>>(alpha, beta, a, invb, C, logscale);
}

static int check_device(int device, int* sm_count) {
  cudaDeviceProp prop;
  BVG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(BVG_ERR_ARCH, "device %d is sm_%d%d; libbvg is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return 0;
}

static int check_device_cached(int device) {
  static std::mutex mu;
  static std::map<int, int> ok;           // device -> status of check_device
  std::lock_guard<std::mutex> lk(mu);
  auto it = ok.find(device);
  if (it != ok.end() && it->second == 0) return 0;
  const int rc = check_device(device, nullptr);
  if (rc == 0) ok[device] = 0;
  return rc;
}

}  // namespace bvg

using namespace bvg;

// =============================================================================== C ABI
extern "C" {

int bvg_version(void) { return BVG_VERSION; }
const char* bvg_last_error(void) { return last_error().c_str(); }

int bvg_plan_create(const bvg_config* cfg, int device, bvg_plan** out) {
  BVG_REQUIRE(cfg && out, "bvg_plan_create: null argument");
  BVG_REQUIRE(cfg->num_upsamples >= 1 && cfg->num_upsamples <= BVG_MAX_UPS, "num_upsamples out of range");
  BVG_REQUIRE(cfg->num_kernels >= 1 && cfg->num_kernels <= BVG_MAX_KERNELS, "num_kernels out of range");
  BVG_REQUIRE(cfg->gpt_dim % 8 == 0 && cfg->upsample_initial_channel % (8 << cfg->num_upsamples) == 0,
              "channel counts must stay multiples of 8 at every stage");
  int sm = 0, rc;
  BVG_CUDA(cudaSetDevice(device));
  if ((rc = check_device(device, &sm))) return rc;
  bvg_plan* p = new bvg_plan();
  p->cfg = *cfg;
  p->device = device;
  p->sm_count = sm;
  p->n_stages = cfg->num_upsamples;
  p->C[0] = cfg->upsample_initial_channel;
  p->rate[0] = 1;
  for (int i = 0; i < p->n_stages; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u < 1 || k % u != 0 || (k - u) % 2 != 0) {
      delete p;
      return fail(BVG_ERR_UNSUPPORTED, "upsample (u=%d,k=%d): need k %% u == 0 and even k-u", u, k);
    }
    p->C[i + 1] = p->C[i] / 2;
    p->rate[i + 1] = p->rate[i] * u;
  }
  p->up_total = p->rate[p->n_stages];
  p->cond_off[0] = 0;
  p->cond_total = p->C[0];
  for (int i = 0; i < p->n_stages; ++i) {
    p->cond_off[i + 1] = p->cond_total;
    if (cfg->cond_in_each_up_layer) p->cond_total += p->C[i + 1];
  }
  *out = p;
  return 0;
}

int bvg_plan_destroy(bvg_plan* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (void* q : p->owned) cudaFree(q);
  for (int i = 0; i < 4; ++i)
    if (p->ws[i]) cudaFree(p->ws[i]);
  if (p->condb) cudaFree(p->condb);
  if (p->d_len) cudaFree(p->d_len);
  for (int i = 0; i < kLenSlots; ++i) {
    if (p->h_len[i]) cudaFreeHost(p->h_len[i]);
    if (p->len_ev[i]) cudaEventDestroy(p->len_ev[i]);
  }
  for (const ProfRec& r : p->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  if (p->st_lat) cudaFree(p->st_lat);
  if (p->ev_h2d) cudaEventDestroy(p->ev_h2d);
  if (p->copy_st) cudaStreamDestroy(p->copy_st);
  if (p->st_emb) cudaFree(p->st_emb);
  if (p->st_wav) cudaFree(p->st_wav);
  tc_plan_free(p);
  delete p;
  return 0;
}

int bvg_plan_load_weights(bvg_plan* p, const bvg_tensor_desc* tensors, int n, void* stream) {
  BVG_REQUIRE(p && tensors && n > 0, "bvg_plan_load_weights: null argument");
  BVG_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  TensorMap tm;
  for (int i = 0; i < n; ++i) tm.m[tensors[i].name] = &tensors[i];
  const bvg_config& c = p->cfg;
  int rc;
  if ((rc = load_conv(p, tm, "conv_pre", p->conv_pre, p->C[0], c.gpt_dim, 7, false, st))) return rc;
  for (int i = 0; i < p->n_stages; ++i) {
    if ((rc = load_conv(p, tm, "ups." + std::to_string(i) + ".0", p->ups[i], p->C[i + 1], p->C[i],
                        c.upsample_kernel_sizes[i], true, st)))
      return rc;
    for (int j = 0; j < c.num_kernels; ++j) {
      const int nb = i * c.num_kernels + j;
      const std::string rb = "resblocks." + std::to_string(nb);
      const int K = c.resblock_kernel_sizes[j], Ci = p->C[i + 1];
      for (int m = 0; m < BVG_MAX_DIL; ++m) {
        if ((rc = load_conv(p, tm, rb + ".convs1." + std::to_string(m), p->rb1[nb][m], Ci, Ci, K, false, st))) return rc;
        if ((rc = load_conv(p, tm, rb + ".convs2." + std::to_string(m), p->rb2[nb][m], Ci, Ci, K, false, st))) return rc;
      }
      for (int m = 0; m < 2 * BVG_MAX_DIL; ++m)
        if ((rc = load_act(p, tm, rb + ".activations." + std::to_string(m), p->rba[nb][m], Ci, st))) return rc;
    }
  }
  const int Cl = p->C[p->n_stages];
  if ((rc = load_act(p, tm, "activation_post", p->act_post, Cl, st))) return rc;
  if ((rc = load_conv(p, tm, "conv_post", p->conv_post, 1, Cl, 7, false, st))) return rc;
  // speaker-conditioning 1x1 convs, concatenated
  const int D = c.speaker_embedding_dim;
  if (!p->cond_W) {
    if ((rc = dev_alloc(p, (void**)&p->cond_W, (size_t)p->cond_total * D * sizeof(float)))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->cond_b, (size_t)p->cond_total * sizeof(float)))) return rc;
  }
  for (int i = 0; i <= p->n_stages; ++i) {
    if (i > 0 && !c.cond_in_each_up_layer) break;
    const std::string nm = i == 0 ? "cond_layer" : "conds." + std::to_string(i - 1);
    const float *w, *b;
    if ((rc = need(tm, nm + ".weight", {p->C[i], D, 1}, &w))) return rc;
    if ((rc = need(tm, nm + ".bias", {p->C[i]}, &b))) return rc;
    BVG_CUDA(cudaMemcpyAsync(p->cond_W + (size_t)p->cond_off[i] * D, w, (size_t)p->C[i] * D * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    BVG_CUDA(cudaMemcpyAsync(p->cond_b + p->cond_off[i], b, (size_t)p->C[i] * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  }
  BVG_CUDA(cudaStreamSynchronize(st));   // act taps land in host memory; caller may free tensors
  if ((rc = tc_plan_pack(p, st))) return rc;
  BVG_CUDA(cudaStreamSynchronize(st));
  p->weights_loaded = true;
  return 0;
}

int bvg_decode(bvg_plan* p, const void* latent, int latent_dtype, const int32_t* lengths, int B,
               int Tmax, const float* spk_emb, void* wav_out, int wav_dtype, int precision,
               void* stream) {
  BVG_REQUIRE(p && latent && spk_emb && wav_out, "bvg_decode: null argument");
  if (!p->weights_loaded) return fail(BVG_ERR_STATE, "bvg_decode: no weights loaded");
  BVG_REQUIRE(B >= 1 && Tmax >= 1, "bvg_decode: B=%d Tmax=%d", B, Tmax);
  BVG_REQUIRE(latent_dtype >= BVG_F32 && latent_dtype <= BVG_F16, "bvg_decode: bad latent dtype");
  BVG_REQUIRE(wav_dtype >= BVG_F32 && wav_dtype <= BVG_I16, "bvg_decode: bad wav dtype");
  BVG_REQUIRE((size_t)B <= 65535, "bvg_decode: B too large");
  BVG_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  p->last_launches = 0;
  p->cur_sum_frames = 0;
  for (int b = 0; b < B; ++b) p->cur_sum_frames += lengths ? lengths[b] : Tmax;
  const int* d_len = nullptr;
  int rc;
  if ((rc = upload_lengths(p, lengths, B, Tmax, st, &d_len))) return rc;
  if ((rc = compute_cond_bias(p, spk_emb, B, st))) return rc;
  if (precision != BVG_PREC_F32 && precision != BVG_PREC_BF16) return fail(BVG_ERR_ARG, "bvg_decode: unknown precision %d", precision);
  if (p->ws_prec != precision) {
    // the four workspace buffers are shared by both paths: fp32 words read as bf16 pairs include NaN / Inf patterns,
    // which the bf16 kernels may multiply by zero weights (rows past an utterance's end) — start from zeros instead
    for (int i = 0; i < 4; ++i)
      if (p->ws[i]) BVG_CUDA(cudaMemsetAsync(p->ws[i], 0, p->ws_bytes, st));
    p->ws_prec = precision;
  }
  if (precision == BVG_PREC_F32)
    return decode_f32(p, latent, latent_dtype, d_len, B, Tmax, wav_out, wav_dtype, st);
  if (precision == BVG_PREC_BF16)
    return tc_decode(p, latent, latent_dtype, lengths, d_len, B, Tmax, wav_out, wav_dtype, st);
  return fail(BVG_ERR_ARG, "bvg_decode: unknown precision %d", precision);
}

static size_t dtype_size(int dt) { return dt == BVG_F32 ? 4 : 2; }

static int grow(void** buf, size_t* cap, size_t need_bytes) {
  if (need_bytes <= *cap) return 0;
  BVG_CUDA(cudaDeviceSynchronize());
  if (*buf) BVG_CUDA(cudaFree(*buf));
  *buf = nullptr; *cap = 0;
  BVG_CUDA(cudaMalloc(buf, need_bytes));
  *cap = need_bytes;
  return 0;
}

int bvg_decode_host(bvg_plan* p, const void* latent_host, int latent_dtype, const int32_t* lengths,
                    int B, int Tmax, const float* spk_emb, void* wav_out_host, int wav_dtype,
                    int precision, void* stream) {
  BVG_REQUIRE(p && latent_host && spk_emb && wav_out_host, "bvg_decode_host: null argument");
  BVG_REQUIRE(B >= 1 && Tmax >= 1, "bvg_decode_host: B=%d Tmax=%d", B, Tmax);
  BVG_REQUIRE(latent_dtype >= BVG_F32 && latent_dtype <= BVG_F16, "bad latent dtype");
  BVG_REQUIRE(wav_dtype >= BVG_F32 && wav_dtype <= BVG_I16, "bad wav dtype");
  BVG_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t lat_bytes = (size_t)B * Tmax * p->cfg.gpt_dim * dtype_size(latent_dtype);
  const size_t wav_bytes = (size_t)B * Tmax * p->up_total * dtype_size(wav_dtype);
  int rc;
  if ((rc = grow(&p->st_lat, &p->st_lat_bytes, lat_bytes))) return rc;
  if ((rc = grow(&p->st_wav, &p->st_wav_bytes, wav_bytes))) return rc;
  // The latents go up on the plan's own copy stream, beside whatever the caller has already queued on `st` — in
  // infer.py's order that is the speaker encoder of the same request (models.py:204), whose output `spk_emb` this
  // call consumes — and the decode waits for the copy by event.  The staging buffer is free: the previous call on this
  // plan ended with a stream synchronize.
  if (!p->copy_st) {
    BVG_CUDA(cudaStreamCreateWithFlags(&p->copy_st, cudaStreamNonBlocking));
    BVG_CUDA(cudaEventCreateWithFlags(&p->ev_h2d, cudaEventDisableTiming));
  }
  BVG_CUDA(cudaMemcpyAsync(p->st_lat, latent_host, lat_bytes, cudaMemcpyHostToDevice, p->copy_st));
  BVG_CUDA(cudaEventRecord(p->ev_h2d, p->copy_st));
  BVG_CUDA(cudaStreamWaitEvent(st, p->ev_h2d, 0));
  if ((rc = bvg_decode(p, p->st_lat, latent_dtype, lengths, B, Tmax, spk_emb, p->st_wav, wav_dtype,
                       precision, stream)))
    return rc;
  BVG_CUDA(cudaMemcpyAsync(wav_out_host, p->st_wav, wav_bytes, cudaMemcpyDeviceToHost, st));
  BVG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int bvg_receptive_field_frames(const bvg_plan* p) {
  if (!p) return -1;
  // per-side receptive field in latent frames: conv_pre 3; each ConvTranspose k/u - 1 input
  // samples; each stage the deepest AMP block: sum_d [5 + d(k-1)/2 + 5 + (k-1)/2]; post 5 + 3.
  double rf = 3.0;
  int kmax_i = 0;
  for (int j = 1; j < p->cfg.num_kernels; ++j)
    if (p->cfg.resblock_kernel_sizes[j] > p->cfg.resblock_kernel_sizes[kmax_i]) kmax_i = j;
  const int k = p->cfg.resblock_kernel_sizes[kmax_i];
  for (int i = 0; i < p->n_stages; ++i) {
    rf += (double)(p->cfg.upsample_kernel_sizes[i] / p->cfg.upsample_rates[i]) / p->rate[i];
    double s = 0;
    for (int m = 0; m < BVG_MAX_DIL; ++m)
      s += 5 + p->cfg.resblock_dilation_sizes[kmax_i][m] * (k - 1) / 2 + 5 + (k - 1) / 2;
    rf += s / p->rate[i + 1];
  }
  rf += 8.0 / p->up_total;
  return (int)rf + 2;
}

int bvg_decode_shard(bvg_plan* p, const void* latent, int latent_dtype, int f_begin, int f_end,
                     int f_total, int halo_l, int halo_r, const float* spk_emb, void* wav_out,
                     int wav_dtype, int precision, void* stream) {
  BVG_REQUIRE(p && latent && spk_emb && wav_out, "bvg_decode_shard: null argument");
  BVG_REQUIRE(0 <= f_begin && f_begin < f_end && f_end <= f_total, "bad shard [%d,%d) of %d", f_begin,
              f_end, f_total);
  BVG_REQUIRE(halo_l >= 0 && halo_r >= 0 && f_begin - halo_l >= 0 && f_end + halo_r <= f_total,
              "halos leave the utterance");
  const int rf = bvg_receptive_field_frames(p);
  BVG_REQUIRE(halo_l >= rf || f_begin - halo_l == 0,
              "left halo %d < receptive field %d and shard does not start at frame 0", halo_l, rf);
  BVG_REQUIRE(halo_r >= rf || f_end + halo_r == f_total,
              "right halo %d < receptive field %d and shard does not end at the last frame", halo_r, rf);
  BVG_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int L = halo_l + (f_end - f_begin) + halo_r;
  const size_t esz = dtype_size(wav_dtype);
  int rc;
  if ((rc = grow(&p->st_wav, &p->st_wav_bytes, (size_t)L * p->up_total * esz))) return rc;
  // overlap-recompute: decode the window as an utterance of its own; window ends that are not
  // true sequence ends only perturb samples inside the discarded halos.
  if ((rc = bvg_decode(p, latent, latent_dtype, nullptr, 1, L, spk_emb, p->st_wav, wav_dtype, precision,
                       stream)))
    return rc;
  BVG_CUDA(cudaMemcpyAsync(wav_out, (const char*)p->st_wav + (size_t)halo_l * p->up_total * esz,
                           (size_t)(f_end - f_begin) * p->up_total * esz, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int bvg_shard_setup(bvg_plan* p, const bvg_shard_geom* g, void* stream) {
  BVG_REQUIRE(p && g, "bvg_shard_setup: null argument");
  if (!p->weights_loaded) return fail(BVG_ERR_STATE, "bvg_shard_setup: no weights loaded");
  BVG_CUDA(cudaSetDevice(p->device));
  return tc_shard_setup(p, g, (cudaStream_t)stream);
}
int bvg_shard_halo_frames(bvg_plan* p) { return p ? tc_shard_halo_frames(p) : -1; }
int bvg_shard_export(bvg_plan* p, uint8_t* handles) {
  BVG_REQUIRE(p && handles, "bvg_shard_export: null argument");
  BVG_CUDA(cudaSetDevice(p->device));
  return tc_shard_export(p, handles);
}
int bvg_shard_connect(bvg_plan* p, int side, const uint8_t* handles) {
  BVG_REQUIRE(p && handles, "bvg_shard_connect: null argument");
  BVG_CUDA(cudaSetDevice(p->device));
  return tc_shard_connect(p, side, handles, nullptr, nullptr, nullptr);
}
int bvg_shard_local_ptrs(bvg_plan* p, void** ws0, void** ws1, void** flags) {
  BVG_REQUIRE(p && ws0 && ws1 && flags, "bvg_shard_local_ptrs: null argument");
  return tc_shard_local_ptrs(p, ws0, ws1, flags);
}
int bvg_shard_connect_ptr(bvg_plan* p, int side, void* ws0, void* ws1, void* flags) {
  BVG_REQUIRE(p && ws0 && ws1 && flags, "bvg_shard_connect_ptr: null argument");
  return tc_shard_connect(p, side, nullptr, ws0, ws1, flags);
}
int bvg_shard_run(bvg_plan* p, int phase, const void* latent, int latent_dtype, const float* spk_emb, void* wav_out,
                  int wav_dtype, int epoch, int wait, void* stream) {
  BVG_REQUIRE(p, "bvg_shard_run: null plan");
  BVG_REQUIRE(wav_dtype >= BVG_F32 && wav_dtype <= BVG_I16, "bvg_shard_run: bad wav dtype");
  BVG_CUDA(cudaSetDevice(p->device));
  return tc_shard_run(p, phase, latent, latent_dtype, spk_emb, wav_out, wav_dtype, epoch, wait, (cudaStream_t)stream);
}
int bvg_shard_error(bvg_plan* p) { return p ? tc_shard_error(p, 0) : -1; }
int bvg_shard_clear_error(bvg_plan* p) { return p ? tc_shard_error(p, 1) : -1; }

int64_t bvg_plan_workspace_bytes(const bvg_plan* p) {
  return p ? (int64_t)(4 * p->ws_bytes + tc_plan_workspace_bytes(p)) : -1;
}
int bvg_plan_last_launches(const bvg_plan* p) { return p ? p->last_launches : -1; }

int bvg_plan_set_profiling(bvg_plan* p, int enable) {
  BVG_REQUIRE(p, "bvg_plan_set_profiling: null plan");
  p->profiling = enable != 0;
  return 0;
}

int bvg_set_tc_fir_max_channels(int max_c) { return bvg::tc_set_fir_max_c(max_c); }
int bvg_set_tc_narrow_max_channels(int max_c) { return bvg::tc_set_nar_max_c(max_c); }
int bvg_set_tc_split_min_channels(int min_c) { return bvg::tc_set_split_min_c(min_c); }
int bvg_set_tc_residual_mma(int on) { return bvg::tc_set_residual_mma(on); }
int bvg_set_graphs(int on) { return bvg::tc_set_graphs(on); }
int bvg_set_tc_cluster(int on) { return bvg::tc_set_cluster(on); }
int bvg_set_pdl(int on) { return bvg::set_pdl(on); }

int bvg_plan_read_profile(bvg_plan* p, bvg_profile* out) {
  BVG_REQUIRE(p && out, "bvg_plan_read_profile: null argument");
  BVG_CUDA(cudaSetDevice(p->device));
  static const bool dump = getenv("BVG_PROF_DUMP") != nullptr;   // per-launch lines on stderr (tools/)
  int idx = 0;
  for (const ProfRec& r : p->prof) {
    BVG_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    BVG_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (dump)
      fprintf(stderr, "bvg_prof %d cls %d us %.1f gflop %.2f mbytes %.1f\n", idx++, r.cls, ms * 1e3, r.flops * 1e-9,
              r.bytes * 1e-6);
    p->prof_acc.ms[r.cls] += ms;
    p->prof_acc.flops[r.cls] += r.flops;
    p->prof_acc.bytes[r.cls] += r.bytes;
    p->prof_acc.launches[r.cls] += 1;
    p->ev_pool.push_back(r.e0);
    p->ev_pool.push_back(r.e1);
  }
  p->prof.clear();
  *out = p->prof_acc;
  p->prof_acc = bvg_profile{};
  return 0;
}

// ------------------------------------------------------------------------------ per-op
// stream-ordered temporaries of the per-op entry points: freed on every exit path
struct AsyncTmp {
  cudaStream_t st;
  void* ptrs[8];
  int n = 0;
  explicit AsyncTmp(cudaStream_t s) : st(s) {}
  cudaError_t alloc(void** out, size_t bytes) {
    if (n >= 8) return cudaErrorMemoryAllocation;
    cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 4, st);
    if (e == cudaSuccess) ptrs[n++] = *out;
    return e;
  }
  ~AsyncTmp() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
};

int bvg_activation1d(const void* x, void* y, int dtype, int B, int C, int T, const float* up_filter,
                     const float* down_filter, const float* alpha, const float* beta, int logscale,
                     void* stream) {
  BVG_REQUIRE(x && y && up_filter && down_filter && alpha && beta, "bvg_activation1d: null argument");
  BVG_REQUIRE(B >= 1 && C >= 1 && T >= 1 && B <= 65535 && C <= 65535, "bvg_activation1d: bad shape");
  BVG_REQUIRE(dtype == BVG_F32 || dtype == BVG_BF16 || dtype == BVG_F16, "bvg_activation1d: dtype %d", dtype);
  int dev, rc;
  BVG_CUDA(cudaGetDevice(&dev));
  if ((rc = check_device_cached(dev))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // one asynchronous launch on the caller's stream: no host copies, no synchronisation, graph-capturable
  dim3 grid(ceil_div(ceil_div(T, 4), 256), C, B);
  if (dtype == BVG_F32)
    k_act1d<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, C, T, up_filter, down_filter, alpha, beta, logscale);
  else if (dtype == BVG_BF16)
    k_act1d<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, C, T, up_filter,
                                                 down_filter, alpha, beta, logscale);
  else
    k_act1d<__half><<<grid, 256, 0, st>>>((const __half*)x, (__half*)y, C, T, up_filter, down_filter, alpha, beta,
                                          logscale);
  BVG_CUDA(cudaGetLastError());
  return 0;
}

int bvg_amp_layer(const float* x, float* y, const float* resid, int B, int C_in, int C_out, int T,
                  const float* w, const float* bias, int k, int dilation, int act,
                  const float* up_filter, const float* down_filter, const float* alpha,
                  const float* beta, int logscale, int precision, void* stream) {
  BVG_REQUIRE(x && y && w && bias, "bvg_amp_layer: null argument");
  BVG_REQUIRE(!act || (up_filter && down_filter && alpha && beta), "bvg_amp_layer: act needs parameters");
  BVG_REQUIRE(B >= 1 && T >= 1 && C_in % 8 == 0 && C_in >= 8 && C_out >= 1 && dilation >= 1,
              "bvg_amp_layer: bad shape (C_in must be a multiple of 8)");
  int dev, rc;
  BVG_CUDA(cudaGetDevice(&dev));
  if ((rc = check_device_cached(dev))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == BVG_PREC_BF16)
    return tc_amp_layer(x, y, resid, B, C_in, C_out, T, w, bias, k, dilation, act, up_filter,
                        down_filter, alpha, beta, logscale, st);
  BVG_REQUIRE(precision == BVG_PREC_F32, "bvg_amp_layer: unknown precision");
  // layer-level TEST entry point (bvg.h): packs the weights per call and reads the taps back to the host, so it
  // synchronises the stream once; the decode path packs at load time and never does
  AsyncTmp tmp(st);
  float *wp, *a = nullptr, *invb = nullptr;
  const size_t n = (size_t)C_out * C_in * k;
  BVG_CUDA(tmp.alloc((void**)&wp, n * sizeof(float)));
  k_pack_conv_w<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, st>>>(w, wp, C_out, C_in, k);
  BVG_CUDA(cudaGetLastError());
  ConvArgs ca{};
  if (act) {
    BVG_CUDA(cudaMemcpyAsync(ca.act.up, up_filter, 12 * sizeof(float), cudaMemcpyDeviceToHost, st));
    BVG_CUDA(cudaMemcpyAsync(ca.act.dn, down_filter, 12 * sizeof(float), cudaMemcpyDeviceToHost, st));
    BVG_CUDA(tmp.alloc((void**)&a, C_in * sizeof(float)));
    BVG_CUDA(tmp.alloc((void**)&invb, C_in * sizeof(float)));
    k_snake_params<<<ceil_div(C_in, 128), 128, 0, st>>>(alpha, beta, a, invb, C_in, logscale);
    BVG_CUDA(cudaGetLastError());
    BVG_CUDA(cudaStreamSynchronize(st));
    ca.act.a = a; ca.act.invb = invb;
  }
  ca.x = x; ca.x_tstride = T;
  ca.wp = wp; ca.bias = bias; ca.resid = resid;
  ca.out = y; ca.out_dtype = BVG_F32; ca.out_tstride = T; ca.div = 1.f;
  ca.Cin = C_in; ca.Cout = C_out; ca.dil = dilation;
  ca.lengths = nullptr; ca.rate = 1; ca.Tmax = T;
  Ctx cx{nullptr, st, 0, 4.0};
  return act ? launch_conv_k<true>(k, ca, B, cx) : launch_conv_k<false>(k, ca, B, cx);
}

int bvg_conv_transpose1d(const float* x, float* y, int B, int C_in, int C_out, int T, const float* w,
                         const float* bias, int k, int u, int precision, void* stream) {
  BVG_REQUIRE(x && y && w && bias, "bvg_conv_transpose1d: null argument");
  BVG_REQUIRE(B >= 1 && T >= 1 && C_in % 8 == 0 && C_out >= 1, "bvg_conv_transpose1d: bad shape");
  BVG_REQUIRE(u >= 1 && k % u == 0 && (k - u) % 2 == 0, "bvg_conv_transpose1d: need k %% u == 0, even k-u");
  int dev, rc;
  BVG_CUDA(cudaGetDevice(&dev));
  if ((rc = check_device(dev, nullptr))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == BVG_PREC_BF16) return tc_conv_transpose(x, y, B, C_in, C_out, T, w, bias, k, u, st);
  BVG_REQUIRE(precision == BVG_PREC_F32, "bvg_conv_transpose1d: unknown precision");
  AsyncTmp tmp(st);
  float* wp;
  const size_t n = (size_t)C_out * C_in * k;
  BVG_CUDA(tmp.alloc((void**)&wp, n * sizeof(float)));
  k_pack_convtr_w<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, st>>>(w, wp, C_in, C_out, k);
  BVG_CUDA(cudaGetLastError());
  ConvTrArgs a{};
  a.x = x; a.x_tstride = T; a.wp = wp; a.bias = bias;
  a.out = y; a.out_tstride = T * u;
  a.Cin = C_in; a.Cout = C_out; a.KK = k; a.U = u;
  a.lengths = nullptr; a.rate_out = u; a.Tmax_out = T * u;
  Ctx cx{nullptr, st, 2, 4.0};
  return launch_convtr<false>(a, B, cx);
}

}  // extern "C"
