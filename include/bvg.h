/*
 * bvg.h — C ABI of libbvg.so, the B200-native (sm_100a) BigVGAN waveform decoder.
 *
 * Drop-in boundary for ONE path of CreateIntelligens/index-tts-lora:
 *     wav, _ = self.bigvgan(latent, cond_input)          indextts/infer.py:748, :888
 * i.e. indextts/BigVGAN/models.py:203-252 (BigVGAN.forward) minus the ECAPA speaker
 * encoder (models.py:204), which stays in PyTorch on the host side.
 *
 * Conventions (all entry points):
 *   - plain C, no torch / pybind types; loaded with ctypes;
 *   - returns 0 on success, a negative bvg_status otherwise; never throws, never falls
 *     back to another backend or to the CPU.  bvg_last_error() gives the message
 *     (thread-local);
 *   - the CALLER owns every tensor it passes; a plan owns only its packed weights and its
 *     grow-only activation workspace;
 *   - all device work is enqueued on the caller's stream (`stream` is a cudaStream_t passed
 *     as void*; NULL = legacy default stream); no internal synchronisation except where a
 *     function says so (`*_host` variants and bvg_plan_load_weights);
 *   - one plan per (device, config); plans are independent, so 8 plans can be driven from
 *     8 host threads or from one thread round-robin;
 *   - requires compute capability 10.x (B200); anything else → BVG_ERR_ARCH.
 *
 * Tensor layouts are the reference's: activations [B, C, T] channel-major contiguous,
 * latents [B, T, gpt_dim] time-major (models.py:215-222 transposes; we consume time-major
 * directly), Conv1d weights [C_out, C_in, k], ConvTranspose1d weights [C_in, C_out, k].
 */
#ifndef BVG_H_
#define BVG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVG_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum bvg_status {
  BVG_OK = 0,
  BVG_ERR_ARG = -1,     /* bad argument / shape / missing tensor */
  BVG_ERR_CUDA = -2,    /* a CUDA runtime call failed */
  BVG_ERR_ARCH = -3,    /* device is not sm_100 */
  BVG_ERR_STATE = -4,   /* plan has no weights loaded, etc. */
  BVG_ERR_UNSUPPORTED = -5
} bvg_status;

typedef enum bvg_dtype { BVG_F32 = 0, BVG_BF16 = 1, BVG_F16 = 2, BVG_I16 = 3 } bvg_dtype;

/* Arithmetic path.  BVG_PREC_F32: SIMT FFMA everywhere (the 1e-4 exactness path).
 * BVG_PREC_BF16: bf16 operands on tcgen05/TMEM tensor cores, fp32 accumulation, fp32
 * FIR/snake math (the throughput path; waveform SNR >= 40 dB vs the fp32 reference). */
typedef enum bvg_precision { BVG_PREC_F32 = 0, BVG_PREC_BF16 = 1 } bvg_precision;

#define BVG_MAX_UPS 8
#define BVG_MAX_KERNELS 4
#define BVG_MAX_DIL 3

/* Mirrors the `bigvgan:` block of finetune_models/config.yaml:88-146 — only the keys
 * BigVGAN.__init__ reads (models.py:144-199). */
typedef struct bvg_config {
  int32_t gpt_dim;                  /* 1280  (conv_pre input channels, models.py:151) */
  int32_t upsample_initial_channel; /* 1536 */
  int32_t num_upsamples;            /* 6 */
  int32_t upsample_rates[BVG_MAX_UPS];        /* 4,4,4,4,2,2 */
  int32_t upsample_kernel_sizes[BVG_MAX_UPS]; /* 8,8,4,4,4,4 */
  int32_t num_kernels;              /* 3 */
  int32_t resblock_kernel_sizes[BVG_MAX_KERNELS];             /* 3,7,11 */
  int32_t resblock_dilation_sizes[BVG_MAX_KERNELS][BVG_MAX_DIL]; /* 1,3,5 each */
  int32_t speaker_embedding_dim;    /* 512 */
  int32_t cond_in_each_up_layer;    /* 1 */
  int32_t snake_logscale;           /* 1 */
} bvg_config;

/* A named tensor in the reference's state-dict naming AFTER weight-norm folding
 * (models.py:254-262): "conv_pre.weight", "ups.0.0.weight", "resblocks.3.convs1.0.bias",
 * "resblocks.3.activations.0.act.alpha", "...upsample.filter",
 * "...downsample.lowpass.filter", "activation_post.act.beta", "conv_post.weight",
 * "cond_layer.weight", "conds.2.bias", ...  `data` is a DEVICE pointer, fp32, contiguous. */
typedef struct bvg_tensor_desc {
  const char* name;
  const void* data;
  int32_t dtype; /* bvg_dtype; only BVG_F32 accepted */
  int32_t ndim;
  int64_t shape[4];
} bvg_tensor_desc;

typedef struct bvg_plan bvg_plan;

int bvg_version(void);
const char* bvg_last_error(void);

/* Replaces BigVGAN.__init__ (models.py:132-199) for the generator half. */
int bvg_plan_create(const bvg_config* cfg, int device, bvg_plan** out);
int bvg_plan_destroy(bvg_plan* plan);

/* Replaces load_state_dict + remove_weight_norm + dtype cast (infer.py:392-409): copies and
 * re-packs the folded fp32 weights into the plan's own buffers (fp32 tap-major for the SIMT
 * path, bf16 UMMA core-matrix tiles for the tcgen05 path) and precomputes exp(alpha),
 * 1/(exp(beta)+1e-9).  Synchronises `stream` before returning (the caller may free its
 * tensors afterwards). */
int bvg_plan_load_weights(bvg_plan* plan, const bvg_tensor_desc* tensors, int n, void* stream);

/* Replaces BigVGAN.forward (models.py:212-252) after the speaker encoder.
 *   latent   device, [B, Tmax, gpt_dim], dtype latent_dtype (F32 / BF16 / F16)
 *   lengths  HOST int32[B], valid latent frames per utterance (NULL = all Tmax).  Each
 *            utterance is decoded with true sequence-edge rules at ITS length, so a ragged
 *            batch equals per-utterance reference runs (the reference has no masking).
 *   spk_emb  device fp32 [B, speaker_embedding_dim] — ECAPA output (models.py:204,212)
 *   wav_out  device, [B, 1, Tmax * prod(upsample_rates)], dtype wav_dtype:
 *            F32/BF16/F16 = tanh output (models.py:250);  I16 = clamp(32767*wav) as in
 *            infer.py:892.  Samples beyond an utterance's length are written as 0.
 * The bf16 (tcgen05) path takes at most 512 utterances per call (the per-launch tile table lives in shared
 * memory); more returns BVG_ERR_UNSUPPORTED — split the batch.
 */
int bvg_decode(bvg_plan* plan, const void* latent, int latent_dtype, const int32_t* lengths,
               int B, int Tmax, const float* spk_emb, void* wav_out, int wav_dtype,
               int precision, void* stream);

/* Same, with HOST buffers for the bulk data (pinned for async copies): H2D of the latent (on a copy
 * stream owned by the plan, so that it runs beside work already queued on `stream`, e.g. the speaker
 * encoder of the same request; the decode waits for it by event), decode, D2H of wav, then synchronises `stream`.  `spk_emb` stays a DEVICE pointer (the ECAPA
 * encoder runs on the device).  This is the end-to-end call bench.py times as `e2e`. */
int bvg_decode_host(bvg_plan* plan, const void* latent_host, int latent_dtype,
                    const int32_t* lengths, int B, int Tmax, const float* spk_emb,
                    void* wav_out_host, int wav_dtype, int precision, void* stream);

/* Time-split decode of one long utterance shard (BASELINE config 5): decodes latent frames
 * [f_begin, f_end) of an utterance of f_total frames given the shard's latent WITH
 * `halo_l`/`halo_r` extra frames actually present on each side (overlap-recompute: the halos
 * are consumed, true-edge rules apply only where the shard touches frame 0 / f_total).
 *   latent   device [halo_l + (f_end-f_begin) + halo_r, gpt_dim]
 *   wav_out  device [(f_end-f_begin) * prod(upsample_rates)]
 */
int bvg_decode_shard(bvg_plan* plan, const void* latent, int latent_dtype, int f_begin,
                     int f_end, int f_total, int halo_l, int halo_r, const float* spk_emb,
                     void* wav_out, int wav_dtype, int precision, void* stream);

/* ---- time split with NVLink P2P halo exchange (one process per GPU; bf16 path) --------------
 * A long utterance is cut into contiguous frame ranges, one per GPU.  Every rank calls
 * bvg_shard_setup with its range and its neighbours' sizes, exports three CUDA-IPC handles
 * (bvg_shard_export) that the host exchanges by any means (torch.distributed object gather),
 * connects to its neighbours (bvg_shard_connect) and then runs the S+1 phases of bvg_shard_run
 * in order.  Between phases the ranks exchange receptive-field halos of the stage output by
 * direct peer stores over NVLink plus a system-scope flag per (side, stage); `wait` != 0 makes a
 * phase first wait (on the device) for its neighbours' flags of this `epoch` (use a new, larger
 * epoch for every decode and put a host barrier between decodes).  With all "ranks" in ONE
 * process on one device (tests) connect by pointer and run phase p of every rank before phase
 * p+1 of any, with wait = 0. */
typedef struct bvg_shard_geom {
  int32_t f_begin, f_end, f_total; /* this rank decodes latent frames [f_begin, f_end) of f_total */
  int32_t own_left, own_right;     /* frame counts of the left / right neighbour (0 if none) */
  int32_t own_max;                 /* largest frame count over all ranks (common buffer stride) */
} bvg_shard_geom;
int bvg_shard_setup(bvg_plan* plan, const bvg_shard_geom* geom, void* stream);
/* latent frames of context the caller must supply on every side that is not a sequence end */
int bvg_shard_halo_frames(bvg_plan* plan);
int bvg_shard_export(bvg_plan* plan, uint8_t* handles /* 3 x 64 bytes */);
int bvg_shard_connect(bvg_plan* plan, int side /* 0 left, 1 right */, const uint8_t* handles);
int bvg_shard_local_ptrs(bvg_plan* plan, void** ws0, void** ws1, void** flags);
int bvg_shard_connect_ptr(bvg_plan* plan, int side, void* ws0, void* ws1, void* flags);
/* phase 0 needs `latent` = frames [f_begin - halo_l, f_end + halo_r) (halo = bvg_shard_halo_frames
 * on non-end sides) and spk_emb [1, D]; the last phase (num_upsamples) writes the rank's own
 * (f_end - f_begin) * prod(upsample_rates) samples to wav_out. */
int bvg_shard_run(bvg_plan* plan, int phase, const void* latent, int latent_dtype, const float* spk_emb,
                  void* wav_out, int wav_dtype, int epoch, int wait, void* stream);
/* 0 = fine, 1/2 = a phase timed out (~4 s) waiting for the left/right neighbour's halo rows; the waveform of that
 * decode is then invalid.  Reads a mapped host word: no device access, no synchronisation — the word is final once
 * the stream the phases ran on has been synchronised.  Every later bvg_shard_run fails with BVG_ERR_STATE until
 * bvg_shard_clear_error (which returns the code it cleared). */
int bvg_shard_error(bvg_plan* plan);
int bvg_shard_clear_error(bvg_plan* plan);

/* Number of latent frames of context each side that makes bvg_decode_shard exact
 * (receptive field of the generator in latent frames, rounded up). */
int bvg_receptive_field_frames(const bvg_plan* plan);

/* Bytes of workspace currently held / kernels launched by the last decode (for bench.py's
 * gpu_launches). */
int64_t bvg_plan_workspace_bytes(const bvg_plan* plan);
int bvg_plan_last_launches(const bvg_plan* plan);

/* Optional per-launch device timing (CUDA events on the caller's stream around every kernel of
 * a decode), accumulated per class until read.  While it is on, the three AMP blocks of a stage,
 * which normally run on three streams, are serialised on the caller's stream so that each
 * event pair brackets exactly one kernel.  Classes: 0 = fused AMP layers of the
 * tensor-bound stages (C >= 192), 1 = fused AMP layers of the small-channel stages, 2 =
 * conv_pre / ConvTranspose1d / cond, 3 = activation_post + conv_post + tanh. */
#define BVG_PROFILE_CLASSES 4
typedef struct bvg_profile {
  double ms[BVG_PROFILE_CLASSES];     /* summed device time */
  double flops[BVG_PROFILE_CLASSES];  /* algorithmic FLOPs (2*Cin*Cout*k per output sample, valid samples) */
  double bytes[BVG_PROFILE_CLASSES];  /* algorithmic HBM bytes (in + out [+ resid] + weights once) */
  int32_t launches[BVG_PROFILE_CLASSES];
} bvg_profile;
int bvg_plan_set_profiling(bvg_plan* plan, int enable);
/* Synchronises the events recorded so far, adds them up, clears the accumulators. */
int bvg_plan_read_profile(bvg_plan* plan, bvg_profile* out);

/* Experimental: narrow activated AMP layers (C_in, C_out <= max_c, at most 64) take k_amp_fir, which runs both
 * kaiser-sinc FIRs of Activation1d on the tensor cores (csrc/amp_fir.cuh), instead of k_amp_tc.  Process-wide;
 * 0 (the default, or the BVG_FIR_MAX_C environment variable at first use) keeps every layer on k_amp_tc.
 * Returns the previous value. */
int bvg_set_tc_fir_max_channels(int max_c);

/* bf16 path: square activated AMP layers with C <= max_c channels (at most 96) take k_amp_nar (csrc/amp_nar.cuh): both
 * kaiser-sinc FIRs of Activation1d as banded-Toeplitz tcgen05 MMAs streamed block by block through TMEM rings, SnakeBeta
 * on the CUDA cores, the dilated conv as in k_amp_tc.  Experimental (slower than k_amp_tc as measured).
 * Process-wide.  Default 0 = every layer on k_amp_tc (or the BVG_NAR_MAX_C environment variable at first use); 96 = all
 * three narrow stages.  Returns the previous value. */
int bvg_set_tc_narrow_max_channels(int max_c);

/* bf16 path: AMP layers with C_in >= min_c run in split form — Activation1d once per layer in a streaming kernel
 * (csrc/act_blk.cuh) into an L2-sized scratch buffer, then the dilated conv as the same tcgen05 kernel without its
 * activation role — instead of the fused kernel, which repeats the activation for every 256-wide column tile of the
 * conv (3x at C = 768, 2x at C = 384).  The z rows, and therefore the output, are bit-identical in both forms.
 * Process-wide; 0 = every layer fused.  Default: the BVG_SPLIT_MIN_C environment variable at first use, else the
 * built-in threshold.  Returns the previous value. */
int bvg_set_tc_split_min_channels(int min_c);

/* bf16 path: the `+ x` of a conv pair (models.py:72) and the running sum over the AMP blocks (:239-245) are accumulated
 * by the tensor core as D += R x I (TMA-staged rows, identity weight tile) on the layers where that is faster than
 * loading and adding the rows in the epilogue warps (every C >= 192 layer; narrow layers with <= 36 conv MMAs per
 * tile).  on = 0 keeps every add in the epilogue.  Both forms add the same bf16 rows to the same fp32 accumulator; only
 * the order of the additions differs.  Process-wide; default on (or BVG_RMMA=0 at first use).  Returns the previous value. */
int bvg_set_tc_residual_mma(int on);

/* bf16 path, launch latency.  (1) Programmatic dependent launch: every kernel of a decode is launched with the
 * programmatic-stream-serialization attribute, releases its successor first thing (griddepcontrol.launch_dependents)
 * and waits for its predecessor (griddepcontrol.wait) only after its own prologue (TMEM allocation, mbarrier init, tile
 * prefix table), so the prologue and the grid-launch latency of kernel n+1 overlap kernel n wherever SMs are free.
 * (2) CUDA graphs: the part of a decode that only touches plan-owned buffers (conv_pre ... activation_post, 114 of the
 * 118 launches on three streams) is stream-captured the second time a (B, Tmax, lengths) shape is seen and replayed
 * afterwards; the launches that touch caller pointers stay ordinary launches, so latent / waveform addresses may change
 * from call to call.  Results are bit-identical either way.  Process-wide, default on (BVG_PDL=0 / BVG_GRAPHS=0 at
 * first use turn them off).  Both return the previous value. */
int bvg_set_pdl(int on);
int bvg_set_graphs(int on);

/* bf16 path, wide stages: the activated AMP layers whose output needs 2 or 3 column tiles (C = 384, 768) are launched as
 * thread-block clusters of that many CTAs.  The CTAs of a cluster convolve the same time tile; each activates every
 * n-th 32-channel chunk and writes the z tile into the shared-memory ring of ALL of them (st.shared::cluster), so
 * Activation1d runs once per chunk instead of once per column tile; ring slots are released by multicast
 * tcgen05.commit.  Same arithmetic per element, bit-identical results.  Process-wide, default on (BVG_CLUSTER=0 at first
 * use turns it off).  Returns the previous value. */
int bvg_set_tc_cluster(int on);

/* ---- per-op entry points (tests, and the reference's own native-op boundary) ---------- */

/* Supersedes anti_alias_activation_cuda.forward(input, up_filter, down_filter, alpha, beta)
 * (alias_free_activation/cuda/anti_alias_activation.cpp:19-23, .cu:214-256) with the exact
 * edge semantics of the torch Activation1d (alias_free_torch/act.py:24-29).
 *   x, y [B, C, T] contiguous, dtype F32/BF16/F16;  up/down filter fp32[12]; alpha, beta
 *   fp32[C] device;  logscale: 1 = parameters are log-scale (exp applied inside). */
int bvg_activation1d(const void* x, void* y, int dtype, int B, int C, int T,
                     const float* up_filter, const float* down_filter, const float* alpha,
                     const float* beta, int logscale, void* stream);

/* One fused AMPBlock1 layer (models.py:65-74, one `xt = conv(act(x))` step):
 *   y = [acc +] conv1d(act1d(x); w, bias, dilation, padding=(k-1)*d/2) [+ resid]
 * x, y, resid [B, C, T] fp32;  w [C_out=C, C_in=C, k] fp32; precision selects the SIMT or the
 * tcgen05 path (the latter rounds operands to bf16).  act==0 skips the activation (plain
 * Conv1d, used for conv_pre-like layers). */
int bvg_amp_layer(const float* x, float* y, const float* resid, int B, int C_in, int C_out,
                  int T, const float* w, const float* bias, int k, int dilation, int act,
                  const float* up_filter, const float* down_filter, const float* alpha,
                  const float* beta, int logscale, int precision, void* stream);

/* ConvTranspose1d(C_in, C_out, k, stride=u, padding=(k-u)/2) (models.py:157-163), fp32
 * [B,C_in,T] -> [B,C_out,T*u];  w [C_in, C_out, k].  precision BVG_PREC_BF16 runs the decode's tcgen05 launch
 * ((k/u)-tap implicit GEMM, operands rounded to bf16; channels must be multiples of 8). */
int bvg_conv_transpose1d(const float* x, float* y, int B, int C_in, int C_out, int T,
                         const float* w, const float* bias, int k, int u, int precision,
                         void* stream);

/* ---- speaker encoder -------------------------------------------------------------------------
 * ECAPA_TDNN.forward(mel_ref, lens=None) -> [B, emb_dim] (indextts/BigVGAN/ECAPA_TDNN.py:543-581, called at
 * models.py:204) as hand-written fp32 CUDA kernels (csrc/ecapa.cu).  The caller (index_tts_lora_b200/ecapa_native.py)
 * owns the weight tensors and hands over device pointers in kernel-ready form:
 *   conv weights   re-laid out as [C_in][k][C_out] fp32 (from the checkpoint's [C_out][C_in][k]);
 *   BatchNorm      folded to scale = weight / sqrt(running_var + eps), shift = bias - running_mean * scale (eval mode);
 *   asp_ctx_w      the [att][2 * mfa] slice of asp.tdnn's 1x1 weight that multiplies the broadcast mean | std context;
 *   asp_tdnn.w     its remaining [mfa][1][att] part.
 * Pointers must stay valid for the life of the handle.  Same error convention as the plan API. */
typedef struct bvg_ecapa_tdnn {
  const float* w;        /* [cin][k][cout] */
  const float* bias;     /* [cout] */
  const float* bn_scale; /* [cout] or NULL (plain conv) */
  const float* bn_shift;
  int32_t cin, cout, k, dil, relu;
} bvg_ecapa_tdnn;
typedef struct bvg_ecapa_block {          /* SERes2NetBlock, ECAPA_TDNN.py:341-426 */
  bvg_ecapa_tdnn tdnn1, res2[7], tdnn2;
  const float *se_w1, *se_b1, *se_w2, *se_b2;   /* [se][C], [se], [C][se], [C] */
} bvg_ecapa_block;
typedef struct bvg_ecapa_desc {
  int32_t in_channels, channels, scale, se_channels, att_channels, mfa_channels, emb_dim;
  bvg_ecapa_tdnn block0;
  bvg_ecapa_block blocks[3];
  bvg_ecapa_tdnn mfa, asp_tdnn, asp_conv;
  const float* asp_ctx_w;                 /* [att][2 * mfa] */
  const float *asp_bn_scale, *asp_bn_shift;   /* [2 * mfa] */
  const float *fc_w, *fc_b;               /* [emb][2 * mfa], [emb] */
} bvg_ecapa_desc;
typedef struct bvg_ecapa bvg_ecapa;
int bvg_ecapa_create(const bvg_ecapa_desc* desc, int device, bvg_ecapa** out);
int bvg_ecapa_destroy(bvg_ecapa* enc);
/* mel [B, Tm, in_channels] contiguous (F32 / BF16 / F16) on the device -> emb_out fp32 [B, emb_dim]; asynchronous on
 * `stream`; the ~45 launches are replayed from a CUDA graph per (B, Tm). */
int bvg_ecapa_forward(bvg_ecapa* enc, const void* mel, int mel_dtype, int B, int Tm, float* emb_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BVG_H_ */
