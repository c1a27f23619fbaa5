// K4 - gather + mean + blend, and the backward scatter.
//
// Reference lines replaced: module/common.py:107-109 (voice_library.py:31-33):
//   result = stack([reference[n][best.indices[n]]]).mean(dim=2); transpose;
//   result * (1 - alpha) + input * alpha
// Bit-exact model of torch's CPU arithmetic: the k RAW frames are summed
// sequentially in descending-score order in float32, divided by k (true division), and the
// blend uses two separately rounded products followed by one addition (no FMA contraction).
//
// HBM-bound: algorithmic bytes per query frame = k*d*4 read + d*4 write (15,360 B at k=4,
// d=768; +d*4 for the query when alpha != 0).  One CTA of d/4 threads per query frame, one
// float4 per thread per gathered frame, fully coalesced 3 KB rows.
#include <cuda_runtime.h>

#include "common.cuh"

namespace alive {
namespace {

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float blend1(float acc, float kf, float a1, float q, float a0) {
  return __fadd_rn(__fmul_rn(__fdiv_rn(acc, kf), a1), __fmul_rn(q, a0));
}
__device__ __forceinline__ float4 finish4(float4 acc, float kf, float a1, float4 q, float a0) {
  return make_float4(blend1(acc.x, kf, a1, q.x, a0), blend1(acc.y, kf, a1, q.y, a0),
                     blend1(acc.z, kf, a1, q.z, a0), blend1(acc.w, kf, a1, q.w, a0));
}

// rows_lo/rows_n: frames [row_lo, row_lo+n) live in lib_raw (sharded libraries); a frame outside
// contributes zeros in gather_rows_kernel and is a caller error in gather_mean_kernel.
__global__ void gather_mean_kernel(const float* __restrict__ lib_raw, long long n, int d,
                                   const long long* __restrict__ top_idx, int t, int k,
                                   const float* __restrict__ q_raw, float a1, float a0, float* __restrict__ out) {
  const int q = blockIdx.x;
  const int j = threadIdx.x * 4;
  if (j >= d) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = 0; r0 < k; r0 += 8) {
    // all (up to 8) row loads are issued before the order-preserving sequential sum
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k) {
        long long idx = top_idx[static_cast<size_t>(q) * k + r0 + u];
        idx = idx < 0 ? 0 : (idx >= n ? n - 1 : idx);   // never read out of bounds
        v[u] = *reinterpret_cast<const float4*>(lib_raw + static_cast<size_t>(idx) * d + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k) acc = (r0 + u == 0) ? v[u] : add4(acc, v[u]);
    }
  }
  const float4 qv = *reinterpret_cast<const float4*>(q_raw + static_cast<size_t>(q) * d + j);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(q) * d + j) = finish4(acc, static_cast<float>(k), a1, qv, a0);
}

__global__ void gather_rows_kernel(const float* __restrict__ lib_raw, long long n, int d, long long row_lo,
                                   const long long* __restrict__ top_idx, int t, int k, float* __restrict__ rows) {
  const int qr = blockIdx.x;   // q*k + r
  const int j = threadIdx.x * 4;
  if (j >= d) return;
  const long long idx = top_idx[qr] - row_lo;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx >= 0 && idx < n) v = *reinterpret_cast<const float4*>(lib_raw + static_cast<size_t>(idx) * d + j);
  *reinterpret_cast<float4*>(rows + static_cast<size_t>(qr) * d + j) = v;
}

__global__ void mean_blend_kernel(const float* __restrict__ rows, int t, int k, int d,
                                  const float* __restrict__ q_raw, float a1, float a0, float* __restrict__ out) {
  const int q = blockIdx.x;
  const int j = threadIdx.x * 4;
  if (j >= d) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < k; ++r) {
    const float4 v = *reinterpret_cast<const float4*>(rows + (static_cast<size_t>(q) * k + r) * d + j);
    acc = (r == 0) ? v : add4(acc, v);
  }
  const float4 qv = *reinterpret_cast<const float4*>(q_raw + static_cast<size_t>(q) * d + j);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(q) * d + j) = finish4(acc, static_cast<float>(k), a1, qv, a0);
}

// Sharded gather over PEER memory: shard r of the raw library lives on GPU r and is mapped into
// this process (CUDA IPC over NVLink); frame i belongs to the shard with bounds[r] <= i < bounds[r+1].
// Same arithmetic as gather_mean_kernel, so every rank computes the bit-identical result without
// any collective after the top-k merge.
__global__ void gather_mean_peers_kernel(const float* const* __restrict__ shard_raw,
                                         const long long* __restrict__ bounds, int shards, int d,
                                         const long long* __restrict__ top_idx, int t, int k,
                                         const float* __restrict__ q_raw, float a1, float a0,
                                         float* __restrict__ out) {
  const int q = blockIdx.x;
  const int j = threadIdx.x * 4;
  if (j >= d) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = 0; r0 < k; r0 += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k) {
        long long idx = top_idx[static_cast<size_t>(q) * k + r0 + u];
        const long long n_total = bounds[shards];
        idx = idx < 0 ? 0 : (idx >= n_total ? n_total - 1 : idx);
        int s = 0;
        while (s + 1 < shards && idx >= bounds[s + 1]) ++s;
        const float* base = shard_raw[s];
        v[u] = *reinterpret_cast<const float4*>(base + static_cast<size_t>(idx - bounds[s]) * d + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k) acc = (r0 + u == 0) ? v[u] : add4(acc, v[u]);
    }
  }
  const float4 qv = *reinterpret_cast<const float4*>(q_raw + static_cast<size_t>(q) * d + j);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(q) * d + j) = finish4(acc, static_cast<float>(k), a1, qv, a0);
}

__global__ void scatter_grad_kernel(const float* __restrict__ grad_out, const long long* __restrict__ top_idx,
                                    int t, int k, int d, float scale, float* __restrict__ grad_rows, long long n) {
  const int qr = blockIdx.x;   // q*k + r
  const int q = qr / k;
  const long long idx = top_idx[qr];
  if (idx < 0 || idx >= n) return;
  for (int j = threadIdx.x; j < d; j += blockDim.x)
    atomicAdd(grad_rows + static_cast<size_t>(idx) * d + j, scale * grad_out[static_cast<size_t>(q) * d + j]);
}

inline int threads_for(int d) { return ((d / 4 + 31) / 32) * 32; }

}  // namespace
}  // namespace alive

extern "C" int alive_knn_gather_mean(const float* lib_raw, int64_t n, int32_t d, const int64_t* top_idx, int32_t t,
                                     int32_t k, const float* q_raw, float alpha, float* out, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(lib_raw && top_idx && q_raw && out, "alive_knn_gather_mean: NULL argument");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_gather_mean: d must be a multiple of 4, <= 4096");
  ALIVE_REQUIRE(k >= 1 && n >= 1, "alive_knn_gather_mean: bad sizes");
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(lib_raw) | reinterpret_cast<uintptr_t>(q_raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "alive_knn_gather_mean: buffers must be 16-byte aligned");
  if (t <= 0) return 0;
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  gather_mean_kernel<<<t, threads_for(d), 0, as_stream(stream)>>>(lib_raw, n, d, reinterpret_cast<const long long*>(top_idx),
                                                                   t, k, q_raw, a1, alpha, out);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_gather_rows(const float* lib_raw, int64_t n, int32_t d, int64_t row_lo, const int64_t* top_idx,
                                     int32_t t, int32_t k, float* rows, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(lib_raw && top_idx && rows, "alive_knn_gather_rows: NULL argument");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_gather_rows: d must be a multiple of 4, <= 4096");
  if (t <= 0) return 0;
  gather_rows_kernel<<<t * k, threads_for(d), 0, as_stream(stream)>>>(lib_raw, n, d, row_lo,
                                                                       reinterpret_cast<const long long*>(top_idx), t, k, rows);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_mean_blend(const float* rows, int32_t t, int32_t k, int32_t d, const float* q_raw, float alpha,
                                    float* out, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(rows && q_raw && out, "alive_knn_mean_blend: NULL argument");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_mean_blend: d must be a multiple of 4, <= 4096");
  if (t <= 0) return 0;
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  mean_blend_kernel<<<t, threads_for(d), 0, as_stream(stream)>>>(rows, t, k, d, q_raw, a1, alpha, out);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_scatter_grad(const float* grad_out, const int64_t* top_idx, int32_t t, int32_t k, int32_t d,
                                      float scale, float* grad_rows, int64_t n, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(grad_out && top_idx && grad_rows, "alive_knn_scatter_grad: NULL argument");
  if (t <= 0) return 0;
  scatter_grad_kernel<<<t * k, 256, 0, as_stream(stream)>>>(grad_out, reinterpret_cast<const long long*>(top_idx), t, k, d,
                                                             scale, grad_rows, n);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_gather_mean_peers(const float* const* shard_raw, const int64_t* bounds, int32_t shards,
                                           int32_t d, const int64_t* top_idx, int32_t t, int32_t k, const float* q_raw,
                                           float alpha, float* out, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(shard_raw && bounds && top_idx && q_raw && out, "alive_knn_gather_mean_peers: NULL argument");
  ALIVE_REQUIRE(shards >= 1 && shards <= 64, "alive_knn_gather_mean_peers: shards must be in [1,64]");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_gather_mean_peers: d must be a multiple of 4, <= 4096");
  ALIVE_REQUIRE(k >= 1, "alive_knn_gather_mean_peers: bad k");
  if (t <= 0) return 0;
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  gather_mean_peers_kernel<<<t, threads_for(d), 0, as_stream(stream)>>>(
      shard_raw, reinterpret_cast<const long long*>(bounds), shards, d, reinterpret_cast<const long long*>(top_idx), t, k,
      q_raw, a1, alpha, out);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
