// common.cuh — shared host/device helpers for libbvg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/bvg.h"

namespace bvg {

// thread-local last-error string behind bvg_last_error()
std::string& last_error();
int fail(int status, const char* fmt, ...);

#define BVG_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::bvg::fail(BVG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                              \
  } while (0)

#define BVG_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::bvg::fail(BVG_ERR_ARG, __VA_ARGS__);  \
  } while (0)

// 12-tap anti-aliasing filters + per-channel snake parameters of ONE Activation1d
// (alias_free_torch/act.py:9-29).  Taps travel by value in kernel params (constant bank).
struct ActParams {
  const float* a;     // [C]  exp(alpha)              (activations.py:116-119)
  const float* invb;  // [C]  1 / (exp(beta) + 1e-9)  (activations.py:120)
  float up[12];       // resample.py:19-22  (NOT pre-multiplied by the ratio)
  float dn[12];       // filter.py:82-83
};

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }

__device__ __forceinline__ float ld_dyn(const void* base, size_t idx, int dtype) {
  if (dtype == BVG_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == BVG_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}

__device__ __forceinline__ void st_dyn(void* base, size_t idx, int dtype, float v) {
  if (dtype == BVG_F32) {
    reinterpret_cast<float*>(base)[idx] = v;
  } else if (dtype == BVG_BF16) {
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  } else if (dtype == BVG_F16) {
    reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
  } else {  // BVG_I16: infer.py:892  clamp(32767*wav, -32767, 32767) then int16 truncation (:911)
    float s = fminf(fmaxf(32767.0f * v, -32767.0f), 32767.0f);
    reinterpret_cast<int16_t*>(base)[idx] = static_cast<int16_t>(s);
  }
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Timing knock-outs (BVG_DBG bits) change the arithmetic of the kernels; they exist only in builds made with
// -DBVG_EXPERIMENTS (tools/), never in the shipped library.
#ifdef BVG_EXPERIMENTS
#define BVG_DBGBIT(a, bit) (((a).dbg & (bit)) != 0)
#else
#define BVG_DBGBIT(a, bit) false
#endif

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still running; it calls pdl_wait() before it touches anything the
// predecessor wrote (or writes anything the predecessor may still read), and every kernel calls
// pdl_launch_dependents() first thing so that its successor's CTAs can take idle SMs and run their prologue (TMEM
// allocation, mbarrier init, tile prefix table, tensor-map prefetch) early.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                             int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl && pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {                       // thread-block cluster of `cluster` CTAs along x (grid.x is a multiple of it)
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
  return launch_kc(kern, grid, block, smem, st, pdl, 0, static_cast<Args&&>(args)...);
}

// cudaFuncSetAttribute is per (function, device): a plan per GPU in one process, or a module moved between GPUs,
// must opt every device in.  Thread-safe; returns a cudaError_t.
cudaError_t func_attr_once(const void* fn, cudaFuncAttribute attr, int value);

}  // namespace bvg
