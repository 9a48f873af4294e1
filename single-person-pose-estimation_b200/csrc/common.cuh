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

extern int g_debug[16];

}  // namespace hgb
