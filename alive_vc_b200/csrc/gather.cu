// K4 - gather + mean + blend, the multi-GPU merge + peer-memory gather, and the backward scatter.
//
// Reference lines replaced: module/common.py:107-109 (voice_library.py:31-33):
//   result = stack([reference[n][best.indices[n]]]).mean(dim=2); transpose;
//   result * (1 - alpha) + input * alpha
// Bit-exact model of torch's CPU arithmetic: the k RAW frames are summed
// sequentially in descending-score order in float32, divided by k (true division), and the
// blend uses two separately rounded products followed by one addition (no FMA contraction).
//
// HBM-bound: algorithmic bytes per query frame = k*d*4 read + d*4 write (15,360 B at k=4, d=768).
// The query row of the blend is only fetched when it can change the result: alpha != 0, a
// non-finite query (0 * inf = NaN in the reference) or a mean of exactly zero (sign of zero).
//
// gather_mean_warp_kernel: ONE WARP per query frame, persistent grid: lane l owns float4 l, l+32, ...
// of the row, so every gathered frame is read as kVec fully coalesced 512-byte requests and ALL k*kVec
// requests of a query are in flight before the first add (24 x 16 B per lane at k=4, d=768); the
// neighbour indices of the warp's NEXT query are fetched while the current one is gathered.
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
// mean * a1 + q * a0 WITHOUT reading q: exact when a0 == 0, q is finite and the mean is not a zero
// (x + (+-0) == x for x != 0); the caller checks the first two, `zero` reports the third
__device__ __forceinline__ float4 finish4_noq(float4 acc, float kf, float a1, bool* zero) {
  const float4 m = make_float4(__fmul_rn(__fdiv_rn(acc.x, kf), a1), __fmul_rn(__fdiv_rn(acc.y, kf), a1),
                               __fmul_rn(__fdiv_rn(acc.z, kf), a1), __fmul_rn(__fdiv_rn(acc.w, kf), a1));
  *zero = m.x == 0.f || m.y == 0.f || m.z == 0.f || m.w == 0.f;
  return m;
}

// frame `idx` of a library that is either one local block (shards == 0: base0, n rows) or split over
// `shards` blocks (device array shard_raw, frame i in shard s iff bounds[s] <= i < bounds[s+1])
struct FrameSource {
  const float* base0;
  long long n;
  const float* const* shard_raw;
  const long long* bounds;
  int shards;
};
__device__ __forceinline__ const float* frame_row(const FrameSource& fs, long long idx, int d) {
  if (fs.shards == 0) {
    idx = idx < 0 ? 0 : (idx >= fs.n ? fs.n - 1 : idx);             // never read out of bounds
    return fs.base0 + static_cast<size_t>(idx) * d;
  }
  const long long n_total = fs.bounds[fs.shards];
  idx = idx < 0 ? 0 : (idx >= n_total ? n_total - 1 : idx);
  int s = 0;
  while (s + 1 < fs.shards && idx >= fs.bounds[s + 1]) ++s;
  return fs.shard_raw[s] + static_cast<size_t>(idx - fs.bounds[s]) * d;
}

constexpr int kGatherWarps = 4;      // 128-thread CTAs: ~160 registers per thread at d = 768 -> 3 CTAs per SM

// One query frame by one warp.  Lane j < k holds neighbour index j (`my_idx`; k <= 32).  The frames are
// gathered in batches of kBatch rows (all kBatch * kVec 16-byte loads of a lane are issued before the first add;
// kBatch * kVec <= 24 keeps the batch in registers: 4 rows at d = 768, i.e. the whole k = 4 neighbourhood).
// kEagerQ: the query row is known to be needed (alpha != 0): it is requested together with the first batch of frames
// instead of after the sum.
template <int kVec, bool kEagerQ>
__device__ __forceinline__ void gather_query(const FrameSource& fs, int d, int k, long long my_idx,
                                             const float* __restrict__ q_row, bool need_q, float a1, float a0,
                                             float* __restrict__ out_row, int lane) {
  constexpr int kBatch = (24 / kVec) < 1 ? 1 : ((24 / kVec) > 8 ? 8 : (24 / kVec));
  float4 acc[kVec];
  float4 qv[kEagerQ ? kVec : 1];
  const float4* q4 = reinterpret_cast<const float4*>(q_row);
  if constexpr (kEagerQ) {
#pragma unroll
    for (int c = 0; c < kVec; ++c) qv[c] = __ldg(q4 + lane + 32 * c);
  }
  for (int r0 = 0; r0 < k; r0 += kBatch) {
    float4 v[kBatch][kVec];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const long long id = __shfl_sync(0xffffffffu, my_idx, (r0 + u) & 31);
      if (r0 + u < k) {
        const float4* row = reinterpret_cast<const float4*>(frame_row(fs, id, d));
#pragma unroll
        for (int c = 0; c < kVec; ++c) v[u][c] = __ldg(row + lane + 32 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (r0 + u < k) {
#pragma unroll
        for (int c = 0; c < kVec; ++c) acc[c] = (r0 + u == 0) ? v[u][c] : add4(acc[c], v[u][c]);
      }
    }
  }
  const float kf = static_cast<float>(k);
  float4* o4 = reinterpret_cast<float4*>(out_row);
#pragma unroll
  for (int c = 0; c < kVec; ++c) {
    float4 r;
    if constexpr (kEagerQ) {
      r = finish4(acc[c], kf, a1, qv[c], a0);
    } else {
      bool zero = false;
      r = finish4_noq(acc[c], kf, a1, &zero);
      if (need_q || zero) r = finish4(acc[c], kf, a1, q4[lane + 32 * c], a0);
    }
    o4[lane + 32 * c] = r;
  }
}

template <int kVec, bool kEagerQ>
__global__ void __launch_bounds__(kGatherWarps * 32, kEagerQ ? 2 : 3)
gather_mean_warp_kernel(const FrameSource fs, int d, const long long* __restrict__ top_idx, int t, int k,
                        const float* __restrict__ q_raw, const float* __restrict__ q_norm, float a1, float a0,
                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * kGatherWarps;
  // lane j < k keeps neighbour j of the current query; the next query's are requested one iteration ahead
  long long my_idx = 0;
  float my_qn = 1.f;
  int q = warp_global;
  if (q < t) {
    if (lane < k) my_idx = top_idx[static_cast<size_t>(q) * k + lane];
    if (q_norm) my_qn = q_norm[q];
  }
  while (q < t) {
    const int q_next = q + n_warps;
    long long nx_idx = 0;
    float nx_qn = 1.f;
    if (q_next < t) {
      if (lane < k) nx_idx = top_idx[static_cast<size_t>(q_next) * k + lane];
      if (q_norm) nx_qn = q_norm[q_next];
    }
    const bool need_q = a0 != 0.f || q_norm == nullptr || !isfinite(my_qn);
    gather_query<kVec, kEagerQ>(fs, d, k, my_idx, q_raw + static_cast<size_t>(q) * d, need_q, a1, a0,
                                out + static_cast<size_t>(q) * d, lane);
    my_idx = nx_idx;
    my_qn = nx_qn;
    q = q_next;
  }
}

// any d (multiple of 4): one CTA of d/4 threads per query frame
__global__ void gather_mean_cta_kernel(const FrameSource fs, int d, const long long* __restrict__ top_idx, int t, int k,
                                       const float* __restrict__ q_raw, const float* __restrict__ q_norm, float a1,
                                       float a0, float* __restrict__ out) {
  const int q = blockIdx.x;
  const int j = threadIdx.x * 4;
  if (j >= d) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = 0; r0 < k; r0 += 8) {
    // all (up to 8) row loads are issued before the order-preserving sequential sum
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k)
        v[u] = *reinterpret_cast<const float4*>(frame_row(fs, top_idx[static_cast<size_t>(q) * k + r0 + u], d) + j);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (r0 + u < k) acc = (r0 + u == 0) ? v[u] : add4(acc, v[u]);
    }
  }
  const bool need_q = a0 != 0.f || q_norm == nullptr || !isfinite(q_norm[q]);
  bool zero = false;
  float4 r = finish4_noq(acc, static_cast<float>(k), a1, &zero);
  if (need_q || zero)
    r = finish4(acc, static_cast<float>(k), a1, *reinterpret_cast<const float4*>(q_raw + static_cast<size_t>(q) * d + j), a0);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(q) * d + j) = r;
}

inline int threads_for(int d) { return ((d / 4 + 31) / 32) * 32; }

int launch_gather(const FrameSource& fs, int d, const long long* top_idx, int t, int k, const float* q_raw,
                  const float* q_norm, float alpha, float* out, cudaStream_t stream) {
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  static const int force_cta = getenv("ALIVE_KNN_GATHER_CTA") ? atoi(getenv("ALIVE_KNN_GATHER_CTA")) : 0;   // A/B runs
  const int kvec = d / 128;
  const bool warp_ok = !force_cta && d % 128 == 0 && k <= 32 &&
                       (kvec == 1 || kvec == 2 || kvec == 4 || kvec == 6 || kvec == 8 || kvec == 12);
  if (warp_ok) {
    int dev = 0, sms = 148;
    ALIVE_CHECK_CUDA(cudaGetDevice(&dev));
    ALIVE_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static const int per_sm = getenv("ALIVE_KNN_GATHER_CTAS_PER_SM") ? atoi(getenv("ALIVE_KNN_GATHER_CTAS_PER_SM")) : 3;
    int grid = (t + kGatherWarps - 1) / kGatherWarps;
    if (grid > sms * per_sm) grid = sms * per_sm;
    const bool eager = alpha != 0.f || q_norm == nullptr;
#define ALIVE_GATHER_CASE(V)                                                                                                  \
  case V:                                                                                                                      \
    if (eager)                                                                                                                 \
      gather_mean_warp_kernel<V, true><<<grid, kGatherWarps * 32, 0, stream>>>(fs, d, top_idx, t, k, q_raw, q_norm, a1, alpha, out);  \
    else                                                                                                                       \
      gather_mean_warp_kernel<V, false><<<grid, kGatherWarps * 32, 0, stream>>>(fs, d, top_idx, t, k, q_raw, q_norm, a1, alpha, out); \
    break;
    switch (kvec) {
      ALIVE_GATHER_CASE(1)
      ALIVE_GATHER_CASE(2)
      ALIVE_GATHER_CASE(4)
      ALIVE_GATHER_CASE(6)
      ALIVE_GATHER_CASE(8)
      ALIVE_GATHER_CASE(12)
    }
#undef ALIVE_GATHER_CASE
  } else {
    gather_mean_cta_kernel<<<t, threads_for(d), 0, stream>>>(fs, d, top_idx, t, k, q_raw, q_norm, a1, alpha, out);
  }
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
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

// ---------------------------------------------------------------------------------------------
// Multi-GPU: merge of the per-rank exact top-k lists.  Every rank's list of a query arrives as one RECORD
// (what ONE all-gather moved): idx [t, k] int64 followed by score [t, k] float32, records
// `record_stride` bytes apart.  Entries with idx < 0 are padding (a shard with fewer than k frames).
// Total order: score descending, NaN first, global frame index ascending (common.cuh score_better).
// ---------------------------------------------------------------------------------------------
struct Records {
  const unsigned char* base;
  long long stride;        // bytes between the records of consecutive ranks
  int ranks, t, k;
  __device__ __forceinline__ long long idx(int r, int q, int j) const {
    return reinterpret_cast<const long long*>(base + r * stride)[static_cast<size_t>(q) * k + j];
  }
  __device__ __forceinline__ float score(int r, int q, int j) const {
    return reinterpret_cast<const float*>(base + r * stride + static_cast<size_t>(t) * k * 8)[static_cast<size_t>(q) * k + j];
  }
};

// one warp: global top-k of query q from the ranks*k candidates -> sel_i/sel_s[0..k) (shared or global)
__device__ __forceinline__ void warp_merge_query(const Records& rec, int q, float* cs, long long* ci,
                                                 float* sel_s, long long* sel_i, int lane) {
  const int n = rec.ranks * rec.k;
  for (int e = lane; e < n; e += 32) {
    cs[e] = rec.score(e / rec.k, q, e % rec.k);
    ci[e] = rec.idx(e / rec.k, q, e % rec.k);
  }
  __syncwarp();
  float ps = 0.f;
  long long pi = -1;
  for (int r = 0; r < rec.k; ++r) {
    float bs = 0.f;
    long long bi = -1;
    bool have = false;
    for (int e = lane; e < n; e += 32) {
      const float s = cs[e];
      const long long i = ci[e];
      if (i < 0) continue;
      if (r > 0 && !score_better(ps, pi, s, i)) continue;     // not strictly after the previous winner
      if (!have || score_better(s, i, bs, bi)) {
        bs = s;
        bi = i;
        have = true;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int oh = __shfl_xor_sync(0xffffffffu, static_cast<int>(have), o);
      if (oh && (!have || score_better(os, oi, bs, bi))) {
        bs = os;
        bi = oi;
        have = true;
      }
    }
    ps = bs;
    pi = bi;
    if (lane == 0) {
      sel_s[r] = have ? bs : -INFINITY;
      sel_i[r] = have ? bi : -1;
    }
  }
  __syncwarp();
}

constexpr int kMergeWarps = 4;
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_records_kernel(const Records rec, float* __restrict__ top_score, long long* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kMergeWarps + warp;
  const int n = rec.ranks * rec.k;
  long long* ci = reinterpret_cast<long long*>(smem_raw) + warp * n;
  float* cs = reinterpret_cast<float*>(reinterpret_cast<long long*>(smem_raw) + kMergeWarps * n) + warp * n;
  if (q >= rec.t) return;
  warp_merge_query(rec, q, cs, ci, top_score + static_cast<size_t>(q) * rec.k, top_idx + static_cast<size_t>(q) * rec.k, lane);
}

// Fused final step of the row-sharded match: for query rows [row0, row0 + rows) merge the per-rank lists AND
// gather + mean + blend the k winning raw frames from whichever GPU owns them (peer memory over NVLink, or the
// local shard) - one kernel, no collective after the all-gather of the records.  One warp per query.
template <int kVec>
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_gather_kernel(const Records rec, int row0, int rows, const FrameSource fs, int d, const float* __restrict__ q_raw,
                    const float* __restrict__ q_norm, float a1, float a0, float* __restrict__ out,
                    float* __restrict__ top_score, long long* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = rec.ranks * rec.k;
  long long* ci = reinterpret_cast<long long*>(smem_raw) + warp * n;
  float* cs = reinterpret_cast<float*>(reinterpret_cast<long long*>(smem_raw) + kMergeWarps * n) + warp * n;
  // selections live behind the candidate area: [kMergeWarps][k] int64 then [kMergeWarps][k] float
  unsigned char* tail = smem_raw + static_cast<size_t>(kMergeWarps) * n * 12;
  tail += (16 - (reinterpret_cast<uintptr_t>(tail) & 15)) & 15;
  long long* sel_i = reinterpret_cast<long long*>(tail) + warp * rec.k;
  float* sel_s = reinterpret_cast<float*>(reinterpret_cast<long long*>(tail) + kMergeWarps * rec.k) + warp * rec.k;
  for (int r = blockIdx.x * kMergeWarps + warp; r < rows; r += gridDim.x * kMergeWarps) {
    const int q = row0 + r;
    warp_merge_query(rec, q, cs, ci, sel_s, sel_i, lane);
    long long my_idx = 0;
    if (lane < rec.k) {
      my_idx = sel_i[lane];
      if (top_idx) top_idx[static_cast<size_t>(r) * rec.k + lane] = my_idx;
      if (top_score) top_score[static_cast<size_t>(r) * rec.k + lane] = sel_s[lane];
    }
    const bool need_q = a0 != 0.f || q_norm == nullptr || !isfinite(q_norm[q]);
    gather_query<kVec, false>(fs, d, rec.k, my_idx, q_raw + static_cast<size_t>(q) * d, need_q, a1, a0,
                              out + static_cast<size_t>(r) * d, lane);
    __syncwarp();
  }
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

int check_records(const void* records, int64_t record_stride, int32_t ranks, int32_t t, int32_t k, const char* who) {
  ALIVE_REQUIRE(records != nullptr, "%s: NULL records", who);
  ALIVE_REQUIRE(ranks >= 1 && ranks <= 64 && k >= 1 && k <= ALIVE_KNN_MAX_K && t >= 0, "%s: bad sizes", who);
  ALIVE_REQUIRE(record_stride >= static_cast<int64_t>(t) * k * 12 && record_stride % 8 == 0,
                "%s: record stride must be a multiple of 8 and hold t*k*12 bytes", who);
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 7) == 0, "%s: records must be 8-byte aligned", who);
  return 0;
}

}  // namespace
}  // namespace alive

extern "C" int alive_knn_gather_mean(const float* lib_raw, int64_t n, int32_t d, const int64_t* top_idx, int32_t t,
                                     int32_t k, const float* q_raw, const float* q_norm, float alpha, float* out,
                                     alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(lib_raw && top_idx && q_raw && out, "alive_knn_gather_mean: NULL argument");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_gather_mean: d must be a multiple of 4, <= 4096");
  ALIVE_REQUIRE(k >= 1 && n >= 1, "alive_knn_gather_mean: bad sizes");
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(lib_raw) | reinterpret_cast<uintptr_t>(q_raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "alive_knn_gather_mean: buffers must be 16-byte aligned");
  if (t <= 0) return 0;
  const FrameSource fs{lib_raw, static_cast<long long>(n), nullptr, nullptr, 0};
  return launch_gather(fs, d, reinterpret_cast<const long long*>(top_idx), t, k, q_raw, q_norm, alpha, out, as_stream(stream));
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
                                           const float* q_norm, float alpha, float* out, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(shard_raw && bounds && top_idx && q_raw && out, "alive_knn_gather_mean_peers: NULL argument");
  ALIVE_REQUIRE(shards >= 1 && shards <= 64, "alive_knn_gather_mean_peers: shards must be in [1,64]");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 4096, "alive_knn_gather_mean_peers: d must be a multiple of 4, <= 4096");
  ALIVE_REQUIRE(k >= 1, "alive_knn_gather_mean_peers: bad k");
  if (t <= 0) return 0;
  const FrameSource fs{nullptr, 0, shard_raw, reinterpret_cast<const long long*>(bounds), shards};
  return launch_gather(fs, d, reinterpret_cast<const long long*>(top_idx), t, k, q_raw, q_norm, alpha, out, as_stream(stream));
}

extern "C" int alive_knn_merge_records(const void* records, int64_t record_stride, int32_t ranks, int32_t t, int32_t k,
                                       float* top_score, int64_t* top_idx, alive_stream_t stream) {
  using namespace alive;
  int rc = check_records(records, record_stride, ranks, t, k, "alive_knn_merge_records");
  if (rc) return rc;
  ALIVE_REQUIRE(top_score && top_idx, "alive_knn_merge_records: NULL argument");
  if (t <= 0) return 0;
  const size_t smem = static_cast<size_t>(kMergeWarps) * ranks * k * 12;
  ALIVE_REQUIRE(smem <= 48 * 1024, "alive_knn_merge_records: ranks*k too large");
  const Records rec{static_cast<const unsigned char*>(records), static_cast<long long>(record_stride), ranks, t, k};
  merge_records_kernel<<<(t + kMergeWarps - 1) / kMergeWarps, kMergeWarps * 32, smem, as_stream(stream)>>>(
      rec, top_score, reinterpret_cast<long long*>(top_idx));
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_merge_gather(const void* records, int64_t record_stride, int32_t ranks, int32_t t, int32_t k,
                                      int32_t row0, int32_t rows, const float* const* shard_raw, const int64_t* bounds,
                                      int32_t shards, int32_t d, const float* q_raw, const float* q_norm, float alpha,
                                      float* out, float* top_score, int64_t* top_idx, alive_stream_t stream) {
  using namespace alive;
  int rc = check_records(records, record_stride, ranks, t, k, "alive_knn_merge_gather");
  if (rc) return rc;
  if (rows == 0) return 0;                 // (a rank whose slice of the query frames is empty)
  ALIVE_REQUIRE(shard_raw && bounds && q_raw && out, "alive_knn_merge_gather: NULL argument");
  ALIVE_REQUIRE(shards >= 1 && shards <= 64, "alive_knn_merge_gather: shards must be in [1,64]");
  ALIVE_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= t, "alive_knn_merge_gather: rows [%d, %d) outside [0, %d)", row0,
                row0 + rows, t);
  const int kvec = d / 128;
  ALIVE_REQUIRE(d % 128 == 0 && (kvec == 1 || kvec == 2 || kvec == 4 || kvec == 6 || kvec == 8 || kvec == 12),
                "alive_knn_merge_gather: d must be 128, 256, 512, 768, 1024 or 1536 (got %d)", d);
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(q_raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "alive_knn_merge_gather: buffers must be 16-byte aligned");
  if (rows <= 0) return 0;
  const size_t smem = static_cast<size_t>(kMergeWarps) * ranks * k * 12 + 16 + static_cast<size_t>(kMergeWarps) * k * 12;
  ALIVE_REQUIRE(smem <= 48 * 1024, "alive_knn_merge_gather: ranks*k too large");
  const Records rec{static_cast<const unsigned char*>(records), static_cast<long long>(record_stride), ranks, t, k};
  const FrameSource fs{nullptr, 0, shard_raw, reinterpret_cast<const long long*>(bounds), shards};
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  int grid = (rows + kMergeWarps - 1) / kMergeWarps;
  if (grid > 148 * 8) grid = 148 * 8;
#define ALIVE_MG_CASE(V)                                                                                            \
  case V:                                                                                                            \
    merge_gather_kernel<V><<<grid, kMergeWarps * 32, smem, as_stream(stream)>>>(                                     \
        rec, row0, rows, fs, d, q_raw, q_norm, a1, alpha, out, top_score, reinterpret_cast<long long*>(top_idx));    \
    break;
  switch (kvec) {
    ALIVE_MG_CASE(1)
    ALIVE_MG_CASE(2)
    ALIVE_MG_CASE(4)
    ALIVE_MG_CASE(6)
    ALIVE_MG_CASE(8)
    ALIVE_MG_CASE(12)
  }
#undef ALIVE_MG_CASE
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
