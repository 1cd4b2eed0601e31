// tc_api.h — internal interface of the bf16 tcgen05/TMEM path (decode_tc.cu).
#pragma once
#include "plan.h"

namespace bvg {
int tc_plan_pack(bvg_plan* p, cudaStream_t st);           // bf16 UMMA weight tiles
void tc_plan_free(bvg_plan* p);
int64_t tc_plan_workspace_bytes(const bvg_plan* p);
int tc_decode(bvg_plan* p, const void* latent, int latent_dtype, const int32_t* h_len,
              const int* d_len, int B, int Tmax, void* wav_out, int wav_dtype, cudaStream_t st);
int tc_amp_layer(const float* x, float* y, const float* resid, int B, int C_in, int C_out, int T,
                 const float* w, const float* bias, int k, int dilation, int act,
                 const float* up_filter, const float* down_filter, const float* alpha,
                 const float* beta, int logscale, cudaStream_t st);
int tc_conv_transpose(const float* x, float* y, int B, int C_in, int C_out, int T, const float* w, const float* bias,
                      int k, int u, cudaStream_t st);
void tc_pack_convtr_w(const float* w, float* wp, int Cin, int Cout, int K, cudaStream_t st);
int tc_set_fir_max_c(int v);
int tc_set_nar_max_c(int v);
int tc_set_split_min_c(int v);
int tc_set_residual_mma(int on);
int tc_set_graphs(int on);
int tc_set_cluster(int on);
int set_pdl(int on);
// time-split P2P decode (decode_tc.cu)
int tc_shard_setup(bvg_plan* p, const bvg_shard_geom* g, cudaStream_t st);
int tc_shard_halo_frames(bvg_plan* p);
int tc_shard_local_ptrs(bvg_plan* p, void** ws0, void** ws1, void** flags);
int tc_shard_export(bvg_plan* p, uint8_t* handles);
int tc_shard_connect(bvg_plan* p, int side, const uint8_t* handles, void* ws0, void* ws1, void* flags);
int tc_shard_run(bvg_plan* p, int phase, const void* latent, int latent_dtype, const float* spk_emb, void* wav_out,
                 int wav_dtype, int epoch, int wait, cudaStream_t st);
int tc_shard_error(bvg_plan* p, int clear);
bool tc_shard_pins_ws(const bvg_plan* p);
// helpers implemented in bvg_api.cu (SIMT kernels reused by the tcgen05 path on the blocked layout)
int tc_ensure_ws(bvg_plan* p, size_t bytes_per_buf);
int simt_convtr_blk(bvg_plan* p, const void* x_blk, void* out_blk, int stage, int B, int Tmax, const int* d_len,
                    cudaStream_t st);
int simt_post_blk(bvg_plan* p, const void* x_blk, void* wav, int wav_dtype, int B, int Tmax, const int* d_len,
                  cudaStream_t st);
void tc_pack_conv_w(const float* w, float* wp, int Cout, int Cin, int K, cudaStream_t st);
void tc_snake_params(const float* alpha, const float* beta, float* a, float* invb, int C, int logscale,
                     cudaStream_t st);
// shared with bvg_api.cu
int upload_lengths(bvg_plan* p, const int32_t* lengths, int B, int Tmax, cudaStream_t st,
                   const int** d_out);
int compute_cond_bias(bvg_plan* p, const float* spk_emb, int B, cudaStream_t st);
}  // namespace bvg
