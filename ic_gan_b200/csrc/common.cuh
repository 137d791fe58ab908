// Library-internal helpers: error reporting, launch checks, dtype access.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/icgan_b200.h"

namespace icgan {

void set_error(const char* fmt, ...);

#define ICGAN_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      icgan::set_error(__VA_ARGS__);  \
      return -1;                      \
    }                                 \
  } while (0)

#define ICGAN_CUDA(expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      icgan::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                     \
    }                                                                                  \
  } while (0)

#define ICGAN_LAUNCH_CHECK()                                                            \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      icgan::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                      \
    }                                                                                   \
  } while (0)

__device__ __forceinline__ float ld_as_float(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void st_from_float(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs must opt every one of them in.  Returns true the
// first time it is called for the current device with a given (function-local static) mask.
inline bool first_use_on_this_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace icgan
