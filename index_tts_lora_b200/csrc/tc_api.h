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
// shared with bvg_api.cu
int upload_lengths(bvg_plan* p, const int32_t* lengths, int B, int Tmax, cudaStream_t st,
                   const int** d_out);
int compute_cond_bias(bvg_plan* p, const float* spk_emb, int B, cudaStream_t st);
}  // namespace bvg
