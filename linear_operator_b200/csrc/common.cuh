// common.cuh -- shared helpers for liblob_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "lob_b200.h"

namespace lob {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launch_count;

inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LOB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return LOB_OK;
}

#define LOB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return ::lob::fail(LOB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define LOB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != LOB_OK) return _s; \
  } while (0)

#define LOB_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) return ::lob::fail(LOB_ERR_ARG, std::string(msg)); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t dsize(int dtype) { return dtype == LOB_F64 ? 8 : 4; }

constexpr int kNumSMs = 148;  // B200

// --------------------------------------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum broadcast to all threads; `scratch` holds >= 32 doubles.  Deterministic.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += scratch[i];
  return t;
}

template <typename T>
__device__ __forceinline__ T ldg_stream(const T* p) {
  return __ldg(p);
}

}  // namespace lob

// dtype dispatch: calls `fn.template operator()<T>()`-style lambdas via a macro
#define LOB_DISPATCH_DTYPE(dtype, ...)                              \
  do {                                                              \
    if ((dtype) == LOB_F32) {                                       \
      using scalar_t = float;                                       \
      __VA_ARGS__                                                   \
    } else if ((dtype) == LOB_F64) {                                \
      using scalar_t = double;                                      \
      __VA_ARGS__                                                   \
    } else {                                                        \
      return ::lob::fail(LOB_ERR_ARG, "dtype must be LOB_F32 or LOB_F64"); \
    }                                                               \
  } while (0)
