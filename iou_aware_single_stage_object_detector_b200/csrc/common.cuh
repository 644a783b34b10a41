// Shared host/device helpers for libiou_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/iou_b200.h"

namespace iou {

void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define IOU_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return ::iou::fail(IOU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                         cudaGetErrorString(e__), __FILE__, __LINE__);              \
  } while (0)

#define IOU_REQUIRE(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) return ::iou::fail(IOU_ERR_INVALID, __VA_ARGS__);                  \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(IOU_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return IOU_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int next_pow2_host(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// order-preserving map float -> uint32 (ascending)
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t b) {
  b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

}  // namespace iou
