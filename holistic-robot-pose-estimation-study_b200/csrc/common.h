// Shared helpers for libhrp_b200: error reporting across the C ABI, launch checks, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/hrp_b200.h"

namespace hrp {

std::string& last_error();                                  // thread-local
int fail(int code, const char* fmt, ...);                   // records message, returns code

#define HRP_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::hrp::fail(HRP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                                       \
  } while (0)

#define HRP_CHECK_LAUNCH(what)                                                                      \
  do {                                                                                              \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess)                                                                          \
      return ::hrp::fail(HRP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e));    \
  } while (0)

#define HRP_TRY(expr)                    \
  do {                                   \
    int _s = (expr);                     \
    if (_s != HRP_OK) return _s;         \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();   // SMs of the current device (cached)

}  // namespace hrp
