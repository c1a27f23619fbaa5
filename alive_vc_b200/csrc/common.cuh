// Shared helpers for the alive_knn CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <utility>

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

// internal variants of two ABI entry points used by the one-call pipeline (api.cu): the query pack also
// zeroes the per-item fallback counters, so that alive_knn_finish needs no memset node of its own
// (`item_frames`, `stride_b`: the n frames are n / item_frames batch items stride_b elements apart - the
// [B, D, T] query batch of one call in ONE launch; 0 = a single item)
int pack_impl(const float* x, int64_t n, int32_t d, int64_t stride_n, int64_t stride_d, float* raw, float* norms,
              uint16_t* packed, float* err, uint32_t* stats, int32_t* zero_words, int32_t n_zero,
              alive_stream_t stream, int64_t item_frames = 0, int64_t stride_b = 0, uint16_t* lo = nullptr,
              float* err2 = nullptr, int32_t format = ALIVE_KNN_FORMAT_BF16);
// `after_query_pack`: the launch directly follows the query pack of the same call in `stream`; kernels
// that can use it start early (programmatic dependent launch) and wait for the pack on the device
int search_impl(const uint16_t* q_packed, const uint16_t* lib_packed, const alive_knn_plan_t* plan, float* cand_score,
                int32_t* cand_idx, int after_query_pack, alive_stream_t stream);
int finish_impl(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists, int32_t k,
                const float* q_raw, const float* q_norm, const float* q_err, const float* lib_raw,
                const float* lib_norm, const uint32_t* lib_stats, int64_t n, int32_t d, int32_t r_max,
                int64_t idx_base, float alpha, float* out, float* top_score, int64_t* top_idx, int32_t* sel_n,
                int32_t* fb_list, int32_t* fb_count, int32_t items, int zero_counts, const uint16_t* q_packed,
                uint16_t* qc_packed, float* c_cut, int32_t* c_cnt, int32_t rows_c, alive_stream_t stream);
// second screen pass for uncertified queries (search_sm100.cu, select.cu; driven by api.cu); `items` independent
// (query batch, library) pairs, rows_c compact slots per item; qc_lo / lib_lo: the second bf16 planes (refined
// accumulation) or NULL
int collect_impl(const uint16_t* qc_packed, const uint16_t* lib_packed, const alive_knn_plan_t* plan,
                 const int32_t* active_rows, const float* cut, int32_t* cnt, int32_t* idx, int32_t cap,
                 const uint16_t* qc_lo, const uint16_t* lib_lo, alive_stream_t stream);
int refine_prep_impl(const float* cand_score, const int32_t* cand_idx, int32_t lists, int32_t k, const int32_t* fb_list,
                     const int32_t* fb_count, int32_t t_item, int32_t items, int32_t rows_c, const float* q_raw,
                     const float* q_norm, const float* q_err, const float* q_err2, const uint16_t* q_lo, uint16_t* qc_lo,
                     const float* lib_raw, const float* lib_norm, const uint32_t* lib_stats, int32_t d, float* c_cut,
                     alive_stream_t stream);
int collect_rescore_impl(const int32_t* fb_list, const int32_t* fb_count, int32_t t_item, int32_t items, int32_t rows_c,
                         const int32_t* c_cnt, const int32_t* c_idx, int32_t c_cap, int32_t k, const float* q_raw,
                         const float* q_norm, const float* lib_raw, const float* lib_norm, int64_t n_total, int32_t d,
                         float alpha, float* out, float* top_score, int64_t* top_idx, int64_t idx_base, int32_t* fb2_list,
                         int32_t* fb2_count, alive_stream_t stream);

// programmatic dependent launch (sm_90+): `launch_dependents` lets the next kernel of the stream be
// scheduled while this one still runs, `wait` blocks until the kernels it depends on have completed and
// their writes are visible.  Both are no-ops when the launch carries no programmatic dependency.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch with the programmatic-stream-serialization attribute: the kernel may be scheduled while its
// predecessor in the stream still runs.  Every kernel launched this way executes pdl_wait() before it
// touches global memory, so ordering is unchanged - only launch latency is hidden.  ALIVE_KNN_PDL=0
// turns the attribute off (A/B runs).
inline bool chain_pdl_enabled() {
  static const bool on = !(getenv("ALIVE_KNN_PDL") && atoi(getenv("ALIVE_KNN_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = chain_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static inline cudaStream_t as_stream(alive_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// One-time setup that is PER DEVICE (cudaFuncSetAttribute, device allocations): one process may drive
// several GPUs, from several threads.  Bit i of `done` = the setup has run on device i; two threads racing
// on the same device both run `f` (the setups are idempotent).  Devices >= 64 run it on every call.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <class F> int run(F&& f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
      set_error("cudaGetDevice failed");
      return -2;
    }
    if (dev >= 0 && dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return 0;
    const int rc = f();
    if (rc == 0 && dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return rc;
  }
};

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
