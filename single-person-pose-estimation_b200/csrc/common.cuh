// Shared helpers for libhgb200 (error plumbing, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/hgb200.h"

namespace hgb {

void set_error(const char* fmt, ...);

#define HGB_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::hgb::set_error(__VA_ARGS__);                               \
      return HGB_ERR_INVALID;                                      \
    }                                                              \
  } while (0)

#define HGB_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::hgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                       __FILE__, __LINE__);                                          \
      return HGB_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define HGB_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::hgb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                       __FILE__, __LINE__);                                          \
      return HGB_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

extern int g_debug[32];

// Programmatic dependent launch: every kernel of the training step is launched with
// programmaticStreamSerialization, calls pdl_trigger() first thing and pdl_wait() before its first
// global-memory access, so launch latency and per-kernel setup (mbarrier init, TMEM allocation,
// descriptor prefetch) overlap the tail of the previous kernel.  g_debug[6] = 1 disables it.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_debug[6] ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace hgb
