// Shared helpers for the yad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/yad_b200.h"

namespace yad {

void set_error(const char* fmt, ...);

#define YAD_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      yad::set_error(__VA_ARGS__);               \
      return YAD_ERR_ARG;                        \
    }                                            \
  } while (0)

#define YAD_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      yad::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return YAD_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define YAD_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      yad::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return YAD_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// element load/store with fp32 math
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == YAD_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == YAD_ACT_LRELU02) return v > 0.0f ? v : 0.2f * v;
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// driver entry point resolved by yad_init (no link-time libcuda dependency)
void* tensor_map_encode_fn();
int sm_count();

// ---- programmatic dependent launch (PDL).  The inference step is a chain of ~60 short kernels on one stream; with the
// cudaLaunchAttributeProgrammaticStreamSerialization attribute kernel N+1 is launched while kernel N still runs, executes its
// prologue (barrier init, TMEM allocation, tensor-map prefetch, constant tables) and blocks in pdl_wait() until kernel N has
// completed and flushed its memory.  Rules every kernel launched through launch_pdl follows:
//   * pdl_wait() before the first access to any buffer another kernel writes or reads (only launch-constant tables before it);
//   * pdl_trigger() only AFTER the CTA holds every resource it will ever need (TMEM columns in particular): a dependent CTA that
//     became resident earlier could otherwise hold the columns this CTA is waiting for while itself waiting for this grid.
// YAD_PDL=0 in the environment launches everything fully serialised.
int pdl_mode();
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // Measured (B200, 512-clip step): launched kernel by kernel the CNN stage drops from 2.106 to 2.019 ms with the attribute;
  // replayed from a CUDA graph the step got 0.04 ms SLOWER with programmatic edges (3.66 vs 3.62 ms: graph nodes already
  // follow each other within ~1 us).  Default (YAD_PDL=1): the attribute only outside stream capture; YAD_PDL=2: also inside.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  int mode = pdl_mode();
  if (mode == 1 && cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) mode = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = mode ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace yad
