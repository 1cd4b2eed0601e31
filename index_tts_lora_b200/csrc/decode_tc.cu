// decode_tc.cu — bf16 tcgen05/TMEM path (placeholder until the kernel lands).
#include "tc_api.h"
namespace bvg {
int tc_plan_pack(bvg_plan*, cudaStream_t) { return 0; }
void tc_plan_free(bvg_plan*) {}
int64_t tc_plan_workspace_bytes(const bvg_plan*) { return 0; }
int tc_decode(bvg_plan*, const void*, int, const int32_t*, const int*, int, int, void*, int, cudaStream_t) {
  return fail(BVG_ERR_UNSUPPORTED, "bf16 tcgen05 path not built yet");
}
int tc_amp_layer(const float*, float*, const float*, int, int, int, int, const float*, const float*, int,
                 int, int, const float*, const float*, const float*, const float*, int, cudaStream_t) {
  return fail(BVG_ERR_UNSUPPORTED, "bf16 tcgen05 path not built yet");
}
}  // namespace bvg
