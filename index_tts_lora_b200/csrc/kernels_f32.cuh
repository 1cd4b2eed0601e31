// kernels_f32.cuh — the fp32 SIMT "exactness" path (BVG_PREC_F32): every layer of
// BigVGAN.forward (indextts/BigVGAN/models.py:212-252) in plain FFMA with precise sinf/tanhf.
// It is the 1e-4 parity anchor for the tcgen05 path, not the throughput path.
//
// Layout: activations [B][C][Tstride] channel-major fp32 (the reference's), weights re-packed
// tap-major [C_in][K][C_out] so a thread's C_out run is contiguous.
#pragma once
#include "common.cuh"

namespace bvg {

// ---------------------------------------------------------------------------------------
// Anti-aliased activation pieces (closed form):
//   y[2m]   = 2*sum_{i<6} up[11-2i] * x[clamp(m-3+i)]          resample.py:25-33
//   y[2m+1] = 2*sum_{i<6} up[10-2i] * x[clamp(m-2+i)]
//   s[n]    = y[n] + invb * sin(a*y[n])^2                       activations.py:109-122
//   z[m]    = sum_k dn[k] * s[clamp(2m+k-5, 0, 2T-1)]           filter.py:87-96
// `xw` must hold x[clamp(j)] for j = xlo .. ; n is already clamped to [0, 2T).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float __tanhf_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.TANH, abs err ~5e-4 (bf16 path only)
  return y;
}

template <bool FAST = false>
__device__ __forceinline__ float snake_up_sample(const float* __restrict__ xw, int xlo, int n,
                                                 const ActParams& ap, float a, float invb) {
  const int m = n >> 1;
  float y = 0.f;
  if ((n & 1) == 0) {
    const float* p = xw + (m - 3 - xlo);
#pragma unroll
    for (int i = 0; i < 6; ++i) y = fmaf(ap.up[11 - 2 * i], p[i], y);
  } else {
    const float* p = xw + (m - 2 - xlo);
#pragma unroll
    for (int i = 0; i < 6; ++i) y = fmaf(ap.up[10 - 2 * i], p[i], y);
  }
  y *= 2.0f;
  const float sn = FAST ? __sinf(y * a) : sinf(y * a);   // FAST: MUFU.SIN (bf16 path only)
  return y + invb * (sn * sn);
}

// ---------------------------------------------------------------------------------------
// Fused [Activation1d ->] dilated Conv1d [+ cond bias] [+ residual] [+ running sum] [tanh]
// One launch = one `xt = conv(act(x))` step of AMPBlock1.forward (models.py:68-71), or
// conv_pre (:226-228), or activation_post+conv_post+tanh (:248-250).
// ---------------------------------------------------------------------------------------
struct ConvArgs {
  const void* x;        // channel-major fp32 [B][Cin][x_tstride], or (TM_IN) time-major [B][Tmax][Cin]
  int x_dtype;          // TM_IN only: dtype of the latent
  int x_tstride;
  const float* wp;      // [Cin][K][Cout]
  const float* bias;    // [Cout]
  const float* bias_b;  // [B][bias_b_stride] speaker-conditioning add (models.py:228,236) or null
  int bias_b_stride;
  const float* resid;   // [B][Cout][out_tstride] or null              (models.py:72)
  const float* acc_in;  // running sum over the 3 AMP blocks or null    (models.py:239-244)
  void* out;            // [B][Cout][out_tstride]
  int out_dtype;
  int out_tstride;
  float div;            // final division (num_kernels, models.py:245) or 1
  int Cin, Cout, dil;
  const int* lengths;   // device int32[B] latent frames, or null
  int rate;             // samples of this layer per latent frame
  int Tmax;             // frames_max * rate
  int tanh_out;         // models.py:250
  int zero_tail;        // write zeros for t in [T_b, Tmax)
  ActParams act;
};

// XL: layout of x — 0 channel-major fp32 [B][Cin][T]; 1 time-major latent [B][Tmax][Cin] (any
// dtype); 2 blocked bf16 [B][Cin/8][T][8] (the tcgen05 path's layout, used by conv_post there).
template <int K, bool ACT, int TY, int NC, int NT, int XL, bool FAST = false>
__global__ void __launch_bounds__(256) k_conv_f32(const ConvArgs a) {
  constexpr bool TM_IN = (XL == 1);
  constexpr int TX = 256 / TY;
  constexpr int TT = TX * NT;
  constexpr int COB = TY * NC;
  constexpr int CK = 8;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int b = blockIdx.z;
  const int T = a.lengths ? a.lengths[b] * a.rate : a.Tmax;
  const int t0 = blockIdx.x * TT;
  const int co0 = blockIdx.y * COB;
  if (t0 >= T) {
    if (a.zero_tail) {
      for (int idx = tid; idx < COB * TT; idx += 256) {
        const int co = co0 + idx / TT, t = t0 + idx % TT;
        if (co < a.Cout && t < a.Tmax)
          st_dyn(a.out, ((size_t)b * a.Cout + co) * a.out_tstride + t, a.out_dtype, 0.f);
      }
    }
    return;
  }
  const int hc = a.dil * (K - 1) / 2;
  const int ZW = TT + 2 * hc;     // z[m], m = t0-hc .. t0+TT+hc-1
  const int SW = 2 * ZW + 12;     // s[n], n = 2(t0-hc)-6 ..
  const int XW = ZW + 12;         // x[j], j = t0-hc-6 ..
  const int n_lo = 2 * (t0 - hc) - 6;
  const int x_lo = t0 - hc - 6;

  extern __shared__ float sm[];
  float* zs = sm;                     // [CK][ZW]
  float* ws = zs + CK * ZW;           // [CK][K][COB]
  float* ss = ws + CK * K * COB;      // [CK][SW]   (ACT)
  float* xs = ss + (ACT ? CK * SW : 0);  // [CK][XW] (ACT)

  float acc[NT][NC];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[i][q] = 0.f;

  const float* xf = reinterpret_cast<const float*>(a.x);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x);

  for (int ci0 = 0; ci0 < a.Cin; ci0 += CK) {
    // ---- stage weights [CK][K][COB] (zero beyond Cout)
    for (int idx = tid; idx < CK * K * COB; idx += 256) {
      const int o = idx % COB, cj = idx / COB;
      const int co = co0 + o;
      ws[idx] = (co < a.Cout) ? __ldg(a.wp + ((size_t)(ci0) * K + cj) * a.Cout + co) : 0.f;
    }
    if (ACT) {
      // ---- x window with replicate clamp (resample.py:28)
      for (int idx = tid; idx < CK * XW; idx += 256) {
        int c, i;
        if (XL == 2) { c = idx % CK; i = idx / CK; } else { c = idx / XW; i = idx % XW; }
        int t = x_lo + i;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        if (XL == 2)
          xs[c * XW + i] = __bfloat162float(xb[(((size_t)b * (a.Cin >> 3) + (ci0 >> 3)) * a.x_tstride + t) * 8 + c]);
        else
          xs[c * XW + i] = __ldg(xf + ((size_t)b * a.Cin + ci0 + c) * a.x_tstride + t);
      }
      __syncthreads();
      // ---- s = snake(upsample(x)) on the clamped 2x grid
      for (int idx = tid; idx < CK * SW; idx += 256) {
        const int c = idx / SW, i = idx % SW;
        int n = n_lo + i;
        n = n < 0 ? 0 : (n > 2 * T - 1 ? 2 * T - 1 : n);
        ss[idx] = snake_up_sample<FAST>(xs + c * XW, x_lo, n, a.act, __ldg(a.act.a + ci0 + c),
                                        __ldg(a.act.invb + ci0 + c));
      }
      __syncthreads();
      // ---- z = downsample(s); conv zero padding outside [0,T) (utils.py:59)
      for (int idx = tid; idx < CK * ZW; idx += 256) {
        const int c = idx / ZW, i = idx % ZW;
        const int m = t0 - hc + i;
        float z = 0.f;
        if (m >= 0 && m < T) {
          const float* sp = ss + c * SW + (2 * m - 5 - n_lo);
#pragma unroll
          for (int k = 0; k < 12; ++k) z = fmaf(a.act.dn[k], sp[k], z);
        }
        zs[idx] = z;
      }
    } else {
      for (int idx = tid; idx < CK * ZW; idx += 256) {
        int c, i;
        if (XL != 0) { c = idx % CK; i = idx / CK; } else { c = idx / ZW; i = idx % ZW; }
        const int m = t0 - hc + i;
        float z = 0.f;
        if (m >= 0 && m < T) {
          if (TM_IN)
            z = ld_dyn(a.x, ((size_t)b * a.Tmax + m) * a.Cin + ci0 + c, a.x_dtype);
          else if (XL == 2)
            z = __bfloat162float(xb[(((size_t)b * (a.Cin >> 3) + (ci0 >> 3)) * a.x_tstride + m) * 8 + c]);
          else
            z = __ldg(xf + ((size_t)b * a.Cin + ci0 + c) * a.x_tstride + m);
        }
        zs[c * ZW + i] = z;
      }
    }
    __syncthreads();
    // ---- implicit GEMM over (ci, tap)
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
      const float* zr = zs + c * ZW + tx;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float wv[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) wv[q] = ws[(c * K + j) * COB + ty * NC + q];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const float zv = zr[TX * i + j * a.dil];
#pragma unroll
          for (int q = 0; q < NC; ++q) acc[i][q] = fmaf(zv, wv[q], acc[i][q]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    const int co = co0 + ty * NC + q;
    if (co >= a.Cout) continue;
    float bsum = __ldg(a.bias + co);
    if (a.bias_b) bsum += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int t = t0 + tx + TX * i;
      if (t >= a.Tmax) continue;
      const size_t o = ((size_t)b * a.Cout + co) * a.out_tstride + t;
      if (t >= T) {
        if (a.zero_tail) st_dyn(a.out, o, a.out_dtype, 0.f);
        continue;
      }
      float v = acc[i][q] + bsum;
      if (a.resid) v = v + __ldg(a.resid + o);
      if (a.acc_in) v = a.acc_in[o] + v;
      if (a.div != 1.0f) v = v / a.div;
      if (a.tanh_out) v = FAST ? __tanhf_fast(v) : tanhf(v);
      st_dyn(a.out, o, a.out_dtype, v);
    }
  }
}

// ---------------------------------------------------------------------------------------
// ConvTranspose1d(C_in, C_out, k, stride=u, padding=(k-u)/2) + cond bias (models.py:232-236)
//   out[co,n] = b[co] + sum_ci sum_kk x[ci,j] * W[ci,co,kk],  n = j*u - p + kk
// Polyphase: for output n, r=(n+p)%u, q=(n+p)/u: taps kk = r + m*u with j = q - m, m < k/u.
// ---------------------------------------------------------------------------------------
struct ConvTrArgs {
  const void* x;        // fp32 [B][Cin][x_tstride]   (BLK: bf16 [B][Cin/8][x_tstride][8])
  int x_tstride;
  const float* wp;      // [Cin][KK][Cout]
  const float* bias;
  const float* bias_b;
  int bias_b_stride;
  void* out;            // fp32 [B][Cout][out_tstride] (BLK: bf16 [B][Cout/8][out_tstride][8])
  int out_tstride;
  int Cin, Cout, KK, U;
  const int* lengths;
  int rate_out;         // output samples per latent frame
  int Tmax_out;
};

template <int TY, int NC, int NT, bool BLK>
__global__ void __launch_bounds__(256) k_convtr_f32(const ConvTrArgs a) {
  const float* xf = reinterpret_cast<const float*>(a.x);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x);
  constexpr int TX = 256 / TY;
  constexpr int TT = TX * NT;
  constexpr int COB = TY * NC;
  constexpr int CK = 8;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int b = blockIdx.z;
  const int T_out = a.lengths ? a.lengths[b] * a.rate_out : a.Tmax_out;
  const int T_in = T_out / a.U;
  const int n0 = blockIdx.x * TT;
  if (n0 >= T_out) return;
  const int co0 = blockIdx.y * COB;
  const int p = (a.KK - a.U) / 2;
  const int M = a.KK / a.U;
  const int q_lo = (n0 + p) / a.U - (M - 1);
  const int QW = (n0 + TT - 1 + p) / a.U - q_lo + 1;

  extern __shared__ float sm[];
  float* xs = sm;               // [CK][QW]
  float* ws = xs + CK * QW;     // [CK][KK][COB]

  float acc[NT][NC];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[i][q] = 0.f;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += CK) {
    for (int idx = tid; idx < CK * a.KK * COB; idx += 256) {
      const int o = idx % COB, ck = idx / COB;
      const int co = co0 + o;
      ws[idx] = (co < a.Cout) ? __ldg(a.wp + ((size_t)ci0 * a.KK + ck) * a.Cout + co) : 0.f;
    }
    for (int idx = tid; idx < CK * QW; idx += 256) {
      int c, i;
      if (BLK) { c = idx % CK; i = idx / CK; } else { c = idx / QW; i = idx % QW; }
      const int j = q_lo + i;
      float v = 0.f;
      if (j >= 0 && j < T_in)
        v = BLK ? __bfloat162float(xb[(((size_t)b * (a.Cin >> 3) + (ci0 >> 3)) * a.x_tstride + j) * 8 + c])
                : __ldg(xf + ((size_t)b * a.Cin + ci0 + c) * a.x_tstride + j);
      xs[c * QW + i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        const int n = n0 + tx + TX * i;
        const int r = (n + p) % a.U, q = (n + p) / a.U;
        for (int m = 0; m < M; ++m) {
          const float xv = xs[c * QW + (q - m - q_lo)];
          const float* wr = ws + (c * a.KK + r + m * a.U) * COB + ty * NC;
#pragma unroll
          for (int qq = 0; qq < NC; ++qq) acc[i][qq] = fmaf(xv, wr[qq], acc[i][qq]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int qq = 0; qq < NC; ++qq) {
    const int co = co0 + ty * NC + qq;
    if (co >= a.Cout) continue;
    float bsum = __ldg(a.bias + co);
    if (a.bias_b) bsum += __ldg(a.bias_b + (size_t)b * a.bias_b_stride + co);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int n = n0 + tx + TX * i;
      if (n >= T_out) continue;
      if (BLK)
        reinterpret_cast<__nv_bfloat16*>(a.out)[(((size_t)b * (a.Cout >> 3) + (co >> 3)) * a.out_tstride + n) * 8 + (co & 7)] =
            __float2bfloat16_rn(acc[i][qq] + bsum);
      else
        reinterpret_cast<float*>(a.out)[((size_t)b * a.Cout + co) * a.out_tstride + n] = acc[i][qq] + bsum;
    }
  }
  if (BLK) {
    // rows [T_out, T_out+32) are zeroed for the consumer's halo reads (ragged batches)
    for (int idx = tid; idx < COB * 32; idx += 256) {
      const int co = co0 + idx % COB, n = T_out + idx / COB;
      if (co < a.Cout && n < a.Tmax_out && n0 + TT >= T_out)
        reinterpret_cast<__nv_bfloat16*>(a.out)[(((size_t)b * (a.Cout >> 3) + (co >> 3)) * a.out_tstride + n) * 8 + (co & 7)] =
            __float2bfloat16_rn(0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Stand-alone Activation1d on [B][C][T] — the replacement of the reference's native op
// anti_alias_activation_forward (alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-181)
// with the torch path's edge rule (replicate the ACTIVATED edge sample).
// One thread -> 4 consecutive outputs; x window of 16, s window of 18, all in registers.
// ---------------------------------------------------------------------------------------
template <typename T_io>
__device__ __forceinline__ void act1d_body(const T_io* __restrict__ x, T_io* __restrict__ y, int C, int T,
                                           const ActParams& ap, float a, float invb) {
  const int m0 = 4 * (blockIdx.x * 256 + threadIdx.x);
  if (m0 >= T) return;
  const int c = blockIdx.y, b = blockIdx.z;
  const T_io* xr = x + ((size_t)b * C + c) * T;
  T_io* yr = y + ((size_t)b * C + c) * T;
  float xw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int t = m0 - 6 + i;
    t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    xw[i] = ld_as_float<T_io>(xr + t);
  }
  float s[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) {
    int n = 2 * m0 - 5 + i;
    n = n < 0 ? 0 : (n > 2 * T - 1 ? 2 * T - 1 : n);
    s[i] = snake_up_sample(xw, m0 - 6, n, ap, a, invb);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (m0 + r >= T) break;
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) z = fmaf(ap.dn[k], s[2 * r + k], z);
    if constexpr (sizeof(T_io) == 4) yr[m0 + r] = z;
    else if constexpr (std::is_same<T_io, __nv_bfloat16>::value) yr[m0 + r] = __float2bfloat16_rn(z);
    else yr[m0 + r] = __float2half_rn(z);
  }
}

// The op boundary (bvg_activation1d): everything the reference op takes as a tensor stays a DEVICE pointer — the 12
// taps of each filter, alpha and beta — so the call neither copies to the host nor synchronises; like the reference's
// launch (anti_alias_activation_cuda.cu:209) it is one asynchronous kernel on the caller's stream and can be
// captured into a CUDA graph.  a = exp(alpha), 1/(exp(beta)+1e-9) are evaluated per thread (activations.py:116-120).
template <typename T_io>
__global__ void __launch_bounds__(256) k_act1d(const T_io* __restrict__ x, T_io* __restrict__ y, int C, int T,
                                               const float* __restrict__ up_filter,
                                               const float* __restrict__ down_filter,
                                               const float* __restrict__ alpha, const float* __restrict__ beta,
                                               int logscale) {
  ActParams ap;
  ap.a = nullptr; ap.invb = nullptr;
#pragma unroll
  for (int i = 0; i < 12; ++i) { ap.up[i] = __ldg(up_filter + i); ap.dn[i] = __ldg(down_filter + i); }
  const float al = __ldg(alpha + blockIdx.y), be = __ldg(beta + blockIdx.y);
  const float a = logscale ? expf(al) : al;
  const float invb = 1.0f / ((logscale ? expf(be) : be) + 1e-9f);
  act1d_body<T_io>(x, y, C, T, ap, a, invb);
}

// ---------------------------------------------------------------------------------------
// Small helpers
// ---------------------------------------------------------------------------------------
// a = exp(alpha) (or alpha), invb = 1/(exp(beta)+1e-9) (activations.py:116-120)
__global__ void k_snake_params(const float* alpha, const float* beta, float* a, float* invb, int C,
                               int logscale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float al = alpha[i], be = beta[i];
  a[i] = logscale ? expf(al) : al;
  invb[i] = 1.0f / ((logscale ? expf(be) : be) + 1e-9f);
}

// Conv1d weight [Cout][Cin][K] -> [Cin][K][Cout]
__global__ void k_pack_conv_w(const float* w, float* wp, int Cout, int Cin, int K) {
  const size_t n = (size_t)Cout * Cin * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const size_t r = i / Cout;
    const int j = r % K;
    const int ci = r / K;
    wp[i] = w[((size_t)co * Cin + ci) * K + j];
  }
}

// ConvTranspose1d weight [Cin][Cout][K] -> [Cin][K][Cout]
__global__ void k_pack_convtr_w(const float* w, float* wp, int Cin, int Cout, int K) {
  const size_t n = (size_t)Cout * Cin * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int co = i % Cout;
    const size_t r = i / Cout;
    const int j = r % K;
    const int ci = r / K;
    wp[i] = w[((size_t)ci * Cout + co) * K + j];
  }
}

// Speaker-conditioning 1x1 convs (models.py:194-199,228,236) for ALL layers at once:
// condb[b][c] = bc[c] + sum_d Wc[c][d] * emb[b][d];  one warp per (b,c).
__global__ void k_cond_bias(const float* __restrict__ Wc, const float* __restrict__ bc,
                            const float* __restrict__ emb, float* __restrict__ condb, int Ctot,
                            int D, int B) {
  pdl_launch_dependents();
  pdl_wait();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= Ctot * B) return;
  const int b = w / Ctot, c = w % Ctot;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(Wc[(size_t)c * D + d], emb[(size_t)b * D + d], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) condb[(size_t)b * Ctot + c] = s + bc[c];
}

}  // namespace bvg
