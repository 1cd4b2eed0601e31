// plan.h — internal plan structure behind the opaque bvg_plan of include/bvg.h.
#pragma once
#include <vector>

#include "common.cuh"

namespace bvg {

struct ConvW {          // one dense conv layer, fp32 tap-major + (bf16 path) UMMA-tiled weights
  float* wp = nullptr;  // [Cin][K][Cout]
  float* bias = nullptr;
  void* wtc = nullptr;  // bf16 tcgen05 tiles (tc path), layout in amp_tc.cuh
  int Cin = 0, Cout = 0, K = 0;
};

struct ActW {
  float* a = nullptr;
  float* invb = nullptr;
  float up[12], dn[12];
  int C = 0;
};

constexpr int kMaxStages = BVG_MAX_UPS;
constexpr int kMaxBlocks = BVG_MAX_UPS * BVG_MAX_KERNELS;
constexpr int kLenSlots = 8;

struct ProfRec {
  cudaEvent_t e0, e1;
  int cls;
  double flops, bytes;
};

}  // namespace bvg

struct bvg_plan {
  bvg_config cfg;
  int device = 0;
  int sm_count = 0;
  int n_stages = 0;
  int C[bvg::kMaxStages + 1];     // C[0] = conv_pre out; C[i+1] = channels after ups[i]
  int rate[bvg::kMaxStages + 1];  // samples per latent frame at that point (rate[0] = 1)
  int up_total = 1;

  bvg::ConvW conv_pre, conv_post;
  bvg::ConvW ups[bvg::kMaxStages];
  bvg::ConvW rb1[bvg::kMaxBlocks][BVG_MAX_DIL], rb2[bvg::kMaxBlocks][BVG_MAX_DIL];
  bvg::ActW rba[bvg::kMaxBlocks][2 * BVG_MAX_DIL], act_post;

  float* cond_W = nullptr;  // [cond_total][D]
  float* cond_b = nullptr;  // [cond_total]
  int cond_total = 0;
  int cond_off[bvg::kMaxStages + 1];  // [0] = cond_layer, [i+1] = conds[i]

  bool weights_loaded = false;
  void* tc = nullptr;   // bvg::TcPlan (decode_tc.cu)

  // grow-only workspace
  void* ws[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t ws_bytes = 0;  // per buffer
  int ws_prec = -1;     // precision path that last wrote the workspace
  float* condb = nullptr;
  size_t condb_elems = 0;
  int* d_len = nullptr;
  int d_len_cap = 0;
  int* h_len[bvg::kLenSlots] = {};
  cudaEvent_t len_ev[bvg::kLenSlots] = {};
  int h_len_cap = 0;
  int len_slot = 0;
  // host-variant staging
  void* st_lat = nullptr;  size_t st_lat_bytes = 0;
  cudaStream_t copy_st = nullptr;       // bvg_decode_host: the H2D of the latents runs beside the caller's queued work
  cudaEvent_t ev_h2d = nullptr;
  float* st_emb = nullptr; size_t st_emb_elems = 0;
  void* st_wav = nullptr;  size_t st_wav_bytes = 0;

  unsigned long long alloc_gen = 1;   // bumped whenever a plan-owned device buffer is re-allocated (captured graphs die)
  int last_launches = 0;
  double cur_sum_frames = 0;   // sum of valid latent frames of the decode being enqueued
  bool profiling = false;
  std::vector<bvg::ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  bvg_profile prof_acc = {};
  std::vector<void*> owned;  // every cudaMalloc'd weight buffer
};

namespace bvg {
// Bracket one kernel launch with events when profiling is on (bvg_plan_set_profiling).
inline void prof_begin(bvg_plan* p, cudaStream_t st, int cls, double flops, double bytes) {
  if (!p || !p->profiling) return;
  ProfRec r;
  for (cudaEvent_t* e : {&r.e0, &r.e1}) {
    if (!p->ev_pool.empty()) { *e = p->ev_pool.back(); p->ev_pool.pop_back(); }
    else cudaEventCreate(e);
  }
  r.cls = cls; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, st);
  p->prof.push_back(r);
}
inline void prof_end(bvg_plan* p, cudaStream_t st) {
  if (!p || !p->profiling || p->prof.empty()) return;
  cudaEventRecord(p->prof.back().e1, st);
}
}  // namespace bvg
