// Shared helpers for the alive_knn CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "alive_knn.h"

namespace alive {

void set_error(const char* fmt, ...);

#define ALIVE_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::alive::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)

#define ALIVE_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::alive::set_error(__VA_ARGS__);    \
      return -1;                          \
    }                                     \
  } while (0)

static inline cudaStream_t as_stream(alive_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Ordering used everywhere a top-k is taken (mirrors torch.topk: NaN ranks above
// everything; ties resolve to the lowest frame index).
__device__ __forceinline__ bool score_better(float sa, long long ia, float sb, long long ib) {
  const bool na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return ia < ib;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace alive
