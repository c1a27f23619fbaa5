// K2b/K3 - certificate + prune, exact rescoring, exact scan, multi-GPU merge.
//
// Reference lines mirrored: module/common.py:102-105 (voice_library.py:26-29):
//   normalise each frame first (x / |x|, float32 IEEE division), then the dot product,
//   then torch.topk (largest, sorted, NaN first).  The dot is accumulated in fp64 and
//   rounded once, i.e. it is the correctly rounded value the reference's fp32 sgemm
//   approximates; ties resolve to the lowest frame index.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace alive {
namespace {

constexpr int kListLen = ALIVE_KNN_LIST_LEN;
constexpr int kMaxK = ALIVE_KNN_MAX_K;
constexpr int kMaxRMax = 256;
// Slack added to the screening error bound for the fp32 accumulation inside the tensor core.  Model, measured on the
// B200 (tools/gpu_tc_error.py, tests/test_gpu_parity.py::test_tensor_core_accumulation_error_model): every
// tcgen05.mma (K = 16) adds its 16 exact bf16 x bf16 products to the accumulator and TRUNCATES the sum to float32
// (the error is one-sided, never positive) - at most one ulp of the accumulator, |acc| < 2, per instruction.  Observed
// over 10^7 pairs with scores up to 1.0: max 2.1e-6 at d = 768 (48 instructions), 4.6e-6 at d = 1536.  The bound
// used is TWICE the model: n_mma * 2^-22 (d = 768: 1.1e-5), plus the final roundings.
__host__ __device__ inline float accum_slack(int n_mma) { return static_cast<float>(n_mma) * 0x1p-22f + 2e-7f; }

__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// prune (one warp per query): certificate + survivor compaction.  Returns the survivor count,
// or -1 when the query must go to the exact scan.  `sel` may point to global or shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int prune_query(const float* __restrict__ sc, const int* __restrict__ ix, int lists,
                                           int k, float qe, float qn, const unsigned int* __restrict__ lib_stats,
                                           int r_max, int* sel, int lane, int d) {
  const int entries = lists * kListLen;
  // tau: upper bound on every screened score that any list dropped (= max of list minima;
  // a list that never filled has minimum -inf and dropped nothing)
  float tau = -INFINITY;
  for (int l = lane; l < lists; l += 32) tau = fmaxf(tau, sc[l * kListLen + kListLen - 1]);
  tau = warp_max_f32(tau);

  // S_k: k-th largest screened score, by k passes under the total order (score desc, pos asc)
  float prev_s = INFINITY;
  int prev_p = -1;
  float sk = -INFINITY;
  for (int r = 0; r < k; ++r) {
    float best_s = -INFINITY;
    int best_p = 0x7fffffff;
    for (int e = lane; e < entries; e += 32) {
      const float s = sc[e];
      const bool after_prev = (s < prev_s) || (s == prev_s && e > prev_p);
      if (after_prev && (s > best_s || (s == best_s && e < best_p))) {
        best_s = s;
        best_p = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best_s, o);
      const int op = __shfl_xor_sync(0xffffffffu, best_p, o);
      if (os > best_s || (os == best_s && op < best_p)) {
        best_s = os;
        best_p = op;
      }
    }
    prev_s = best_s;
    prev_p = best_p;
    sk = best_s;
  }

  const float le = __uint_as_float(lib_stats[0]);
  const float eps = (le + qe + le * qe + accum_slack(d / 16)) * 1.00001f;
  const float cut = sk - 2.0f * eps - 1e-7f;
  if (lib_stats[1] != 0u || !(qn > 0.f) || !isfinite(qn) || !(sk > -INFINITY) || !(cut > tau)) return -1;

  int count = 0;
  for (int e0 = 0; e0 < entries; e0 += 32) {
    const int e = e0 + lane;
    const bool keep = e < entries && sc[e] >= cut && ix[e] >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = count + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < r_max) sel[pos] = ix[e];
    count += __popc(m);
  }
  if (count > r_max || count < k) return -1;
  return count;
}

__global__ void __launch_bounds__(256)
prune_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int t, int lists, int k,
             const float* __restrict__ q_err, const float* __restrict__ q_norm,
             const unsigned int* __restrict__ lib_stats, int r_max, int* __restrict__ sel_idx,
             int* __restrict__ sel_n, int* __restrict__ fb_list, int* __restrict__ fb_count, int d) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= t) return;
  const size_t entries = static_cast<size_t>(lists) * kListLen;
  const int count = prune_query(cand_score + q * entries, cand_idx + q * entries, lists, k, q_err[q], q_norm[q],
                                lib_stats, r_max, sel_idx + static_cast<size_t>(q) * r_max, lane, d);
  if (lane == 0) {
    sel_n[q] = count;
    if (count < 0) fb_list[atomicAdd(fb_count, 1)] = q;
  }
}

// ------------------------------------------------------------------------------------------
// exact similarity of one library frame against a normalised query held in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot_norm_f64(const float* __restrict__ qh, const float* __restrict__ row,
                                               float nrm, int d, int lane) {
  // the frame's chunks are fetched 6 at a time BEFORE any arithmetic: one DRAM round trip per
  // 768 channels instead of one per 128 (this runs at low occupancy, latency is the cost);
  // the accumulation order (j ascending per lane) does not change
  constexpr int kBatch = 6;
  double acc = 0.0;
  for (int j0 = lane * 4; j0 < d; j0 += kBatch * 128) {
    float4 r[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      if (j0 + i * 128 < d) r[i] = *reinterpret_cast<const float4*>(row + j0 + i * 128);
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      if (j0 + i * 128 < d) {
        const float4 a = *reinterpret_cast<const float4*>(qh + j0 + i * 128);
        acc += static_cast<double>(a.x) * static_cast<double>(__fdiv_rn(r[i].x, nrm));
        acc += static_cast<double>(a.y) * static_cast<double>(__fdiv_rn(r[i].y, nrm));
        acc += static_cast<double>(a.z) * static_cast<double>(__fdiv_rn(r[i].z, nrm));
        acc += static_cast<double>(a.w) * static_cast<double>(__fdiv_rn(r[i].w, nrm));
      }
    }
  }
  return warp_sum_f64(acc);
}

// top-k of (sc[i], id[i]) i<n by k selection rounds, executed by one warp
__device__ __forceinline__ void warp_select_topk(const float* sc, const long long* id, int n, int k,
                                                 float* out_s, long long* out_i, long long base, int lane) {
  float ps = 0.f;
  long long pi = -1;
  bool first = true;
  for (int r = 0; r < k; ++r) {
    float bs = 0.f;
    long long bi = -1;
    bool have = false;
    for (int e = lane; e < n; e += 32) {
      const float s = sc[e];
      const long long i = id[e];
      if (i < 0) continue;
      if (!first && !score_better(ps, pi, s, i)) continue;   // not strictly after the previous winner
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
    first = false;
    if (lane == 0) {
      out_s[r] = have ? bs : -INFINITY;
      out_i[r] = have ? bi + base : -1;
    }
    if (!have) {   // fewer than k valid entries (caller guarantees this cannot happen)
      for (int r2 = r + 1; r2 < k; ++r2)
        if (lane == 0) {
          out_s[r2] = -INFINITY;
          out_i[r2] = -1;
        }
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// rescore: one CTA (4 warps) per query
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rescore_kernel(const float* __restrict__ q_raw, const float* __restrict__ q_norm, int t,
               const float* __restrict__ lib_raw, const float* __restrict__ lib_norm, int d,
               const int* __restrict__ sel_idx, const int* __restrict__ sel_n, int r_max, int k,
               long long idx_base, float* __restrict__ top_score, long long* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qh = reinterpret_cast<float*>(smem_raw);                         // [d]
  long long* cid = reinterpret_cast<long long*>(qh + d);                  // [r_max]
  float* csc = reinterpret_cast<float*>(cid + r_max);                     // [r_max]
  const int q = blockIdx.x;
  const int n_sel = sel_n[q];
  if (n_sel < 0) return;   // exact scan handles it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float qn = q_norm[q];
  for (int j = threadIdx.x; j < d; j += blockDim.x) qh[j] = __fdiv_rn(q_raw[static_cast<size_t>(q) * d + j], qn);
  __syncthreads();
  for (int c = warp; c < n_sel; c += 4) {
    const int idx = sel_idx[static_cast<size_t>(q) * r_max + c];
    const double acc = dot_norm_f64(qh, lib_raw + static_cast<size_t>(idx) * d, lib_norm[idx], d, lane);
    if (lane == 0) {
      csc[c] = static_cast<float>(acc);
      cid[c] = idx;
    }
  }
  __syncthreads();
  if (warp == 0)
    warp_select_topk(csc, cid, n_sel, k, top_score + static_cast<size_t>(q) * k,
                     top_idx + static_cast<size_t>(q) * k, idx_base, lane);
}

// ------------------------------------------------------------------------------------------
// finish: prune + exact rescoring + top-k + (optional) gather-mean-blend, one CTA per query.
// Replaces three launches (prune, rescore, gather) and the survivor list round trip.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float blend_exact(float acc, float kf, float a1, float q, float a0) {
  return __fadd_rn(__fmul_rn(__fdiv_rn(acc, kf), a1), __fmul_rn(q, a0));
}

// out[j..j+3] = mean of the k raw rows (sequential fp32 sum, true division) blended with q.  The query row is
// only fetched when it can change the result (`need_q`: alpha != 0 or a non-finite query; or a mean of exactly
// zero, where the sign of 0 * q matters) - see gather.cu.
__device__ __forceinline__ void gather_mean_row(const float* __restrict__ lib_raw, long long n, int d,
                                                const long long* idx, long long idx_base, int k,
                                                const float* __restrict__ q_row, bool need_q, float a1, float a0,
                                                float* __restrict__ out_row, int tid, int nthreads) {
  for (int j = tid * 4; j < d; j += nthreads * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < k; r0 += 8) {
      // issue up to 8 row loads before the (order-preserving) sequential sum
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r0 + u < k) {
          long long i = idx[r0 + u] - idx_base;
          i = i < 0 ? 0 : (i >= n ? n - 1 : i);
          v[u] = *reinterpret_cast<const float4*>(lib_raw + static_cast<size_t>(i) * d + j);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r0 + u < k) {
          if (r0 + u == 0) acc = v[u];
          else acc = make_float4(__fadd_rn(acc.x, v[u].x), __fadd_rn(acc.y, v[u].y), __fadd_rn(acc.z, v[u].z),
                                 __fadd_rn(acc.w, v[u].w));
        }
      }
    }
    const float kf = static_cast<float>(k);
    float4 r = make_float4(__fmul_rn(__fdiv_rn(acc.x, kf), a1), __fmul_rn(__fdiv_rn(acc.y, kf), a1),
                           __fmul_rn(__fdiv_rn(acc.z, kf), a1), __fmul_rn(__fdiv_rn(acc.w, kf), a1));
    if (need_q || r.x == 0.f || r.y == 0.f || r.z == 0.f || r.w == 0.f) {
      const float4 qv = *reinterpret_cast<const float4*>(q_row + j);
      r = make_float4(blend_exact(acc.x, kf, a1, qv.x, a0), blend_exact(acc.y, kf, a1, qv.y, a0),
                      blend_exact(acc.z, kf, a1, qv.z, a0), blend_exact(acc.w, kf, a1, qv.w, a0));
    }
    *reinterpret_cast<float4*>(out_row + j) = r;
  }
}

#ifdef ALIVE_FINISH_TIMING
__device__ unsigned long long g_finish_t[16];
#define ALIVE_FT(i)                                                         \
  do {                                                                      \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                              \
      g_finish_t[i] = static_cast<unsigned long long>(clock64());          \
    }                                                                       \
  } while (0)
#else
#define ALIVE_FT(i)
#endif
constexpr int kFinishMaxStagedEntries = 6144;   // lists*8 entries staged in shared memory (48 KB) when they fit

// Second screen pass for uncertified queries (one-call pipeline, single item): finish_kernel moves
// the query to slot `fb slot` of a compact query matrix, the tiled search kernel in collect mode
// appends every frame whose screened score reaches the query's cut, collect_rescore_kernel below
// rescores exactly those.  Queries whose buffer overflowed, or that did not get a slot, go on to the
// exhaustive scan through fb2_list.
struct CollectStage {
  const uint16_t* q_packed;   // [t, d] bf16 packed queries of this call
  uint16_t* qc_packed;        // [items * rows_c, d] compact copy, rows_c slots per item (NULL: stage disabled)
  float* c_cut;               // [items * rows_c]
  int* c_cnt;                 // [items * rows_c]
  int rows_c;
};

// monotone map score -> uint32 (larger score = larger key, -0 == +0, NaN above everything, 0 is
// below every score): lets the selections below run on redux.sync instead of shuffle trees
__device__ __forceinline__ unsigned score_key(float s) {
  if (s != s) return 0xFFFFFFFFu;
  const unsigned u = __float_as_uint(s + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(unsigned key) {
  if (key == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key);
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
finish_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int t, int lists, int k,
              const float* __restrict__ q_raw, const float* __restrict__ q_norm, const float* __restrict__ q_err,
              const float* __restrict__ lib_raw, const float* __restrict__ lib_norm,
              const unsigned int* __restrict__ lib_stats, long long n, int d, int r_max, long long idx_base,
              float a1, float a0, float* __restrict__ out, float* __restrict__ top_score,
              long long* __restrict__ top_idx, int* __restrict__ sel_n, int* __restrict__ fb_list,
              int* __restrict__ fb_count, int staged, int t_item, CollectStage cs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qh = reinterpret_cast<float*>(smem_raw);                         // [d]
  long long* cid = reinterpret_cast<long long*>(qh + d);                  // [r_max] (layout shared with rescore_kernel)
  float* csc = reinterpret_cast<float*>(cid + r_max);                     // [r_max]
  int* sel = reinterpret_cast<int*>(csc + r_max);                         // [r_max]
  float* st_sc = reinterpret_cast<float*>(sel + r_max);                   // [entries] (staged only)
  int* st_ix = reinterpret_cast<int*>(st_sc + (staged ? lists * kListLen : 0));
  constexpr int kWarps = kThreads / 32;
  __shared__ int s_total;
  __shared__ int s_fb;
  __shared__ float s_cut;
  __shared__ float s_ccut;
  __shared__ unsigned s_tau[kWarps];
  __shared__ unsigned s_wbest[kWarps * kListLen];
  __shared__ long long s_top[kMaxK];
  pdl_wait();                    // launched behind the search (launch_chained): its lists are complete now
  pdl_launch_dependents();
  ALIVE_FT(0);
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // everything the certificate needs from global memory is requested up front
  const float qn = q_norm[q];
  const float qe = q_err[q];
  const unsigned le_bits = lib_stats[0];
  const unsigned lib_bad = lib_stats[1];
  const size_t entries = static_cast<size_t>(lists) * kListLen;
  const float* g_sc = cand_score + q * entries;
  const int* g_ix = cand_idx + q * entries;
  // stage the screened lists (coalesced) and normalise the query frame (x / |x|, IEEE division)
  if (staged) {
    for (int e = threadIdx.x * 4; e < static_cast<int>(entries); e += kThreads * 4) {
      *reinterpret_cast<float4*>(st_sc + e) = *reinterpret_cast<const float4*>(g_sc + e);
      *reinterpret_cast<int4*>(st_ix + e) = *reinterpret_cast<const int4*>(g_ix + e);
    }
  }
  for (int j = threadIdx.x; j < d; j += kThreads)
    qh[j] = __fdiv_rn(q_raw[static_cast<size_t>(q) * d + j], qn);
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  ALIVE_FT(1);
  // ---- certificate + survivor compaction, spread over the whole CTA (prune_query's logic) ----
  const float* sc = staged ? st_sc : g_sc;
  const int* ix = staged ? st_ix : g_ix;
  const int n_entries = static_cast<int>(entries);
  {
    // tau = max of the list minima
    unsigned tau = 0u;
    for (int l = threadIdx.x; l < lists; l += kThreads) tau = max(tau, score_key(sc[l * kListLen + kListLen - 1]));
    tau = __reduce_max_sync(0xffffffffu, tau);
    if (lane == 0) s_tau[warp] = tau;
    // every warp: the k largest keys of its contiguous slice (duplicates count).  A lane keeps the
    // sorted top-k of its own entries in registers, then k rounds of redux-max + pop.
    const int e_lo = static_cast<int>(static_cast<long long>(n_entries) * warp / kWarps);
    const int e_hi = static_cast<int>(static_cast<long long>(n_entries) * (warp + 1) / kWarps);
    unsigned best[kListLen];
#pragma unroll
    for (int i = 0; i < kListLen; ++i) best[i] = 0u;
    for (int e = e_lo + lane; e < e_hi; e += 32) {
      unsigned v = score_key(sc[e]);
#pragma unroll
      for (int i = 0; i < kListLen; ++i) {       // sorted insert (descending); entries beyond k are never read
        const unsigned hi = max(best[i], v);
        v = min(best[i], v);
        best[i] = hi;
      }
    }
    for (int r = 0; r < k; ++r) {
      const unsigned m = __reduce_max_sync(0xffffffffu, best[0]);
      if (lane == 0) s_wbest[warp * kListLen + r] = m;      // 0 when the slice ran out
      const unsigned has = __ballot_sync(0xffffffffu, best[0] == m);
      if (lane == __ffs(static_cast<int>(has)) - 1) {
#pragma unroll
        for (int i = 0; i + 1 < kListLen; ++i) best[i] = best[i + 1];
        best[kListLen - 1] = 0u;
      }
    }
  }
  __syncthreads();
  ALIVE_FT(2);
  if (warp == 0) {
    // S_k = k-th largest of the kWarps*k slice winners (values only; duplicates count)
    constexpr int kPer = (kWarps * kListLen + 31) / 32;
    unsigned v[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = lane + 32 * i;
      v[i] = (e < kWarps * k) ? s_wbest[(e / k) * kListLen + (e % k)] : 0u;
    }
    unsigned sk_key = 0u;
    for (int r = 0; r < k; ++r) {
      unsigned m = v[0];
#pragma unroll
      for (int i = 1; i < kPer; ++i) m = max(m, v[i]);
      const unsigned mine = m;
      m = __reduce_max_sync(0xffffffffu, m);
      sk_key = m;
      // remove ONE occurrence of the maximum (lowest lane first, first slot first)
      const unsigned has = __ballot_sync(0xffffffffu, mine == m);
      if (lane == __ffs(static_cast<int>(has)) - 1) {
        bool gone = false;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          if (!gone && v[i] == m) {
            v[i] = 0u;
            gone = true;
          }
        }
      }
    }
    unsigned t2k = (lane < kWarps) ? s_tau[lane] : 0u;     // kWarps <= 32
    t2k = __reduce_max_sync(0xffffffffu, t2k);
    if (lane == 0) {
      const float sk = sk_key == 0u ? -INFINITY : key_score(sk_key);
      const float t2 = t2k == 0u ? -INFINITY : key_score(t2k);
      const float le = __uint_as_float(le_bits);
      const float eps = (le + qe + le * qe + accum_slack(d / 16)) * 1.00001f;
      const float cut = sk - 2.0f * eps - 1e-7f;
      const bool fb = lib_bad != 0u || !(qn > 0.f) || !isfinite(qn) || !(sk > -INFINITY) || !(cut > t2);
      s_cut = cut;
      s_fb = fb ? 1 : 0;
      // what the collect pass may use as this query's cut (+inf: nothing is known about the screen)
      s_ccut = (lib_bad == 0u && qn > 0.f && isfinite(qn) && sk > -INFINITY && cut > -INFINITY) ? cut : INFINITY;
    }
  }
  __syncthreads();
  ALIVE_FT(3);
  int n_sel = -1;
  if (s_fb == 0) {
    // compaction of the survivors: one shared-memory atomic per warp and pass (the order of the
    // survivors is immaterial - the final selection is a total order on (score, index))
    const float cut = s_cut;
    for (int e0 = 0; e0 < n_entries; e0 += kThreads) {
      const int e = e0 + threadIdx.x;
      const bool keep = e < n_entries && sc[e] >= cut && ix[e] >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (m != 0u) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_total, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < r_max) sel[pos] = ix[e];
      }
    }
    __syncthreads();
    const int total = s_total;
    n_sel = (total > r_max || total < k) ? -1 : total;
  }
  ALIVE_FT(4);
  if (threadIdx.x == 0) {
    sel_n[q] = n_sel;
    if (n_sel < 0) {
      const int item = q / t_item;                      // uncertified queries are listed per item
      const int slot = atomicAdd(&fb_count[item], 1);
      fb_list[static_cast<size_t>(item) * t_item + slot] = q;
      // collect pass: the query's packed row moves to its slot of the compact query matrix, next to
      // the cut.  One thread copies the 1.5 KB: anything heavier between here and the bare `return`
      // below (a block-wide copy behind a barrier was tried) costs the COMMON path 40% - measured.
      if (cs.qc_packed != nullptr && slot < cs.rows_c) {
        const size_t cslot = static_cast<size_t>(item) * cs.rows_c + slot;          // the item's own block of slots
        const uint4* src = reinterpret_cast<const uint4*>(cs.q_packed + static_cast<size_t>(q) * d);
        uint4* dst = reinterpret_cast<uint4*>(cs.qc_packed + cslot * d);
        for (int j = 0; j < d / 8; ++j) dst[j] = src[j];
        cs.c_cut[cslot] = s_ccut;
        cs.c_cnt[cslot] = 0;
      }
    }
  }
  if (n_sel < 0) return;   // the collect pass / the exact scan (and their gather) handle this query
  for (int c = warp; c < n_sel; c += kWarps) {
    const int idx = sel[c];
    const double acc = dot_norm_f64(qh, lib_raw + static_cast<size_t>(idx) * d, lib_norm[idx], d, lane);
    if (lane == 0) csc[c] = static_cast<float>(acc);
  }
  __syncthreads();
  ALIVE_FT(5);
  if (warp == 0) {
    // exact top-k of the survivors under (score desc, NaN first, index asc): k rounds of
    // redux-max on the score key, redux-min on the index among the lanes that hold that key
    unsigned taken = 0u;                                  // bit i: entry lane + 32*i is already out
    for (int r = 0; r < k; ++r) {
      unsigned bk = 0u;
      int bi = 0x7fffffff, bslot = -1;
      for (int i = 0, c = lane; c < n_sel; ++i, c += 32) {
        if (taken & (1u << i)) continue;
        const unsigned key = score_key(csc[c]);
        const int idx = sel[c];
        if (bslot < 0 || key > bk || (key == bk && idx < bi)) {
          bk = key;
          bi = idx;
          bslot = i;
        }
      }
      const unsigned m = __reduce_max_sync(0xffffffffu, bslot >= 0 ? bk : 0u);
      const int cand = (bslot >= 0 && bk == m) ? bi : 0x7fffffff;
      const int wi = __reduce_min_sync(0xffffffffu, cand);
      if (bslot >= 0 && bk == m && bi == wi) {              // exactly one lane: indices are distinct
        taken |= 1u << bslot;
        const float sv = csc[lane + 32 * bslot];
        s_top[r] = static_cast<long long>(wi) + idx_base;
        top_score[static_cast<size_t>(q) * k + r] = sv;
        top_idx[static_cast<size_t>(q) * k + r] = static_cast<long long>(wi) + idx_base;
      }
    }
  }
  if (out == nullptr) return;
  __syncthreads();
  ALIVE_FT(7);
  gather_mean_row(lib_raw, n, d, s_top, idx_base, k, q_raw + static_cast<size_t>(q) * d, a0 != 0.f || !isfinite(qn), a1, a0,
                  out + static_cast<size_t>(q) * d, threadIdx.x, kThreads);
#ifdef ALIVE_FINISH_TIMING
  __syncthreads();
  ALIVE_FT(8);
#endif
}

// Refinement prologue of the collect pass (libraries packed with their second bf16 plane), one CTA per uncertified
// query that got a slot.  finish_kernel left the query's first-pass cut S_k - 2 eps1 (eps1 ~ 4e-3: on a clustered
// library thousands of frames sit above it).  Here the R best SCREENED entries of the query are rescored exactly;
// the k-th best of those exact scores, L, is a lower bound of the exact k-th best similarity over the whole library
// (k distinct frames reach it).  The refined tensor-core pass computes s2 = hi.hi + hi.lo + lo.hi with
// |s2 - exact| <= eps2 (~4e-5), so every frame of the exact top-k has s2 >= L - eps2: that is the cut the collecting
// epilogue applies - a band of eps2 around the true k-th score instead of 2 eps1 below the screened one.
// Also moves the query's second plane into the compact matrix.
constexpr int kPrepThreads = 256;
constexpr int kPrepR = 64;                 // screened entries rescored for the bound (<= 2x with ties)
__device__ __forceinline__ void
refine_prep_slot(int item, int slot, const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int lists, int k,
                 const int* __restrict__ fb_list, int t_item, int rows_c,
                 const float* __restrict__ q_raw, const float* __restrict__ q_norm, const float* __restrict__ q_err,
                 const float* __restrict__ q_err2, const uint16_t* __restrict__ q_lo, uint16_t* __restrict__ qc_lo,
                 const float* __restrict__ lib_raw, const float* __restrict__ lib_norm,
                 const unsigned int* __restrict__ lib_stats, int d, float* __restrict__ c_cut) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qh = reinterpret_cast<float*>(smem_raw);            // [d] normalised query
  int* sel = reinterpret_cast<int*>(qh + d);                 // [2 * kPrepR] frame indices
  float* csc = reinterpret_cast<float*>(sel + 2 * kPrepR);   // [2 * kPrepR] exact scores
  __shared__ int s_cnt[kPrepThreads / 32];
  __shared__ int s_total;
  const int q = fb_list[static_cast<size_t>(item) * t_item + slot];
  const size_t cslot = static_cast<size_t>(item) * rows_c + slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // the second plane of the query moves next to the first one (finish_kernel parked that)
    const uint4* src = reinterpret_cast<const uint4*>(q_lo + static_cast<size_t>(q) * d);
    uint4* dst = reinterpret_cast<uint4*>(qc_lo + cslot * d);
    for (int j = threadIdx.x; j < d / 8; j += kPrepThreads) dst[j] = src[j];
  }
  if (!(c_cut[cslot] < INFINITY)) return;      // no usable screen (non-finite norms / rows): the exhaustive scan takes it
  const float qn = q_norm[q];
  for (int j = threadIdx.x; j < d; j += kPrepThreads) qh[j] = __fdiv_rn(q_raw[static_cast<size_t>(q) * d + j], qn);
  if (threadIdx.x == 0) s_total = 0;
  const int n_entries = lists * kListLen;
  const float* sc = cand_score + static_cast<size_t>(q) * n_entries;
  const int* ix = cand_idx + static_cast<size_t>(q) * n_entries;
  // largest key threshold that still leaves >= R entries: bisection on the monotone uint32 key of the score
  unsigned lo_key = 1u, hi_key = 0xFFFFFFFEu;                // (0 = padding / -inf lists never reach the count)
  const int want = min(kPrepR, n_entries);
  while (lo_key < hi_key) {
    const unsigned mid = lo_key + ((hi_key - lo_key + 1u) >> 1);
    int c = 0;
    for (int e = threadIdx.x; e < n_entries; e += kPrepThreads) c += (ix[e] >= 0 && score_key(sc[e]) >= mid) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    __syncthreads();
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kPrepThreads / 32; ++w) tot += s_cnt[w];
    if (tot >= want) lo_key = mid;
    else hi_key = mid - 1u;
  }
  __syncthreads();
  for (int e0 = 0; e0 < n_entries; e0 += kPrepThreads) {
    const int e = e0 + threadIdx.x;
    const bool keep = e < n_entries && ix[e] >= 0 && score_key(sc[e]) >= lo_key;
    if (keep) {
      const int pos = atomicAdd(&s_total, 1);
      if (pos < 2 * kPrepR) sel[pos] = ix[e];
    }
  }
  __syncthreads();
  const int n_sel = min(s_total, 2 * kPrepR);
  for (int c = warp; c < n_sel; c += kPrepThreads / 32) {
    const int idx = sel[c];
    const double acc = dot_norm_f64(qh, lib_raw + static_cast<size_t>(idx) * d, lib_norm[idx], d, lane);
    if (lane == 0) csc[c] = static_cast<float>(acc);
  }
  __syncthreads();
  if (warp == 0) {
    // k-th largest exact score among the rescored entries (distinct frames; NaN cannot occur: the library is finite)
    unsigned taken_lo = 0u, taken_hi = 0u, taken_2 = 0u, taken_3 = 0u;      // bit i: entry lane + 32*i used (n_sel <= 128)
    float kth = -INFINITY;
    for (int r = 0; r < k; ++r) {
      float best = -INFINITY;
      int bslot = -1;
      for (int i = 0, c = lane; c < n_sel; ++i, c += 32) {
        const unsigned used = i == 0 ? taken_lo : i == 1 ? taken_hi : i == 2 ? taken_2 : taken_3;
        if (used & 1u) continue;
        if (bslot < 0 || csc[c] > best) {
          best = csc[c];
          bslot = i;
        }
      }
      const unsigned key = bslot >= 0 ? score_key(best) : 0u;
      const unsigned m = __reduce_max_sync(0xffffffffu, key);
      const unsigned has = __ballot_sync(0xffffffffu, bslot >= 0 && key == m);
      if (m == 0u) {
        kth = -INFINITY;                    // fewer than k rescored entries: no bound
        break;
      }
      kth = key_score(m);
      if (lane == __ffs(static_cast<int>(has)) - 1) {                          // exactly one lane retires its entry
        if (bslot == 0) taken_lo = 1u;
        else if (bslot == 1) taken_hi = 1u;
        else if (bslot == 2) taken_2 = 1u;
        else taken_3 = 1u;
      }
    }
    if (lane == 0 && n_sel >= k && kth > -INFINITY) {
      const float le = __uint_as_float(lib_stats[0]), le2 = __uint_as_float(lib_stats[2]);
      const float qe = q_err[q], qe2 = q_err2[q];
      // |refined - exact| <= |ql||rl| + |q2| |r^| + |qh + ql| |r2| + accumulation over 3 d/16 instructions (+ roundings)
      const float eps2 = ((qe + qe2) * (le + le2) + qe2 * 1.001f + le2 * 1.003f + accum_slack(3 * (d / 16))) * 1.0001f + 2e-7f;
      const float cut2 = kth - eps2;
      // never looser than the first-pass cut (both are valid: a frame of the exact top-k clears each of them in ITS score)
      c_cut[cslot] = cut2;
    }
  }
}

// grid = (min(rows_c, one wave of CTAs), items): the CTAs stride over the item's live slots, so the usual case - no
// uncertified query at all - costs a launch of a few hundred CTAs that exit at once instead of rows_c = 8192 of them.
__global__ void __launch_bounds__(kPrepThreads)
refine_prep_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int lists, int k,
                   const int* __restrict__ fb_list, const int* __restrict__ fb_count, int t_item, int rows_c,
                   const float* __restrict__ q_raw, const float* __restrict__ q_norm, const float* __restrict__ q_err,
                   const float* __restrict__ q_err2, const uint16_t* __restrict__ q_lo, uint16_t* __restrict__ qc_lo,
                   const float* __restrict__ lib_raw, const float* __restrict__ lib_norm,
                   const unsigned int* __restrict__ lib_stats, int d, float* __restrict__ c_cut) {
  pdl_launch_dependents();
  pdl_wait();
  const int item = blockIdx.y;
  const int n_fb = min(fb_count[item], rows_c);
  for (int slot = blockIdx.x; slot < n_fb; slot += gridDim.x) {
    refine_prep_slot(item, slot, cand_score, cand_idx, lists, k, fb_list, t_item, rows_c, q_raw, q_norm, q_err, q_err2, q_lo,
                     qc_lo, lib_raw, lib_norm, lib_stats, d, c_cut);
    __syncthreads();                       // the next slot reuses the shared-memory buffers
  }
}

// One CTA per fallback slot: exact rescoring of the collected candidates, top-k, gather.  blockIdx.y = item.
constexpr int kCollectThreads = 512;
__device__ __forceinline__ void
collect_rescore_slot(int item, int slot, const int* __restrict__ fb_list, int rows_c,
                     const int* __restrict__ c_cnt, const int* __restrict__ c_idx, int c_cap, int k,
                     const float* __restrict__ q_raw, const float* __restrict__ q_norm,
                     const float* __restrict__ lib_raw, const float* __restrict__ lib_norm, long long n_total, int d,
                     float a1, float a0, float* __restrict__ out, float* __restrict__ top_score,
                     long long* __restrict__ top_idx, long long idx_base, int* __restrict__ fb2_list,
                     int* __restrict__ fb2_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qh = reinterpret_cast<float*>(smem_raw);          // [d]
  float* csc = qh + d;                                      // [c_cap]
  __shared__ long long s_top[kMaxK];
  constexpr int kWarps = kCollectThreads / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = fb_list[slot];
  const size_t cslot = static_cast<size_t>(item) * rows_c + slot;
  const int cnt = c_cnt[cslot];
  if (cnt > c_cap || cnt < k) {                             // overflow, or no usable cut: exhaustive scan
    if (threadIdx.x == 0) fb2_list[atomicAdd(&fb2_count[item], 1)] = q;
    return;
  }
  const float qn = q_norm[q];
  for (int j = threadIdx.x; j < d; j += kCollectThreads) qh[j] = __fdiv_rn(q_raw[static_cast<size_t>(q) * d + j], qn);
  __syncthreads();
  const int* cand = c_idx + cslot * c_cap;
  for (int c = warp; c < cnt; c += kWarps) {
    const int idx = cand[c];
    const double acc = dot_norm_f64(qh, lib_raw + static_cast<size_t>(idx) * d, lib_norm[idx], d, lane);
    if (lane == 0) csc[c] = static_cast<float>(acc);
  }
  __syncthreads();
  if (warp == 0) {
    // k rounds of (max score key, min index) over the not yet taken candidates; `prev` = last winner
    unsigned pk = 0xFFFFFFFFu;
    int pi = -1;
    for (int r = 0; r < k; ++r) {
      unsigned bk = 0u;
      int bi = 0x7fffffff;
      bool have = false;
      for (int c = lane; c < cnt; c += 32) {
        const unsigned key = score_key(csc[c]);
        const int idx = cand[c];
        const bool after = r == 0 || key < pk || (key == pk && idx > pi);   // strictly after the previous winner
        if (after && (!have || key > bk || (key == bk && idx < bi))) {
          bk = key;
          bi = idx;
          have = true;
        }
      }
      const unsigned m = __reduce_max_sync(0xffffffffu, have ? bk : 0u);
      const int wi = __reduce_min_sync(0xffffffffu, (have && bk == m) ? bi : 0x7fffffff);
      pk = m;
      pi = wi;
      if (lane == 0) {
        s_top[r] = wi;
        top_score[static_cast<size_t>(q) * k + r] = key_score(m);
        top_idx[static_cast<size_t>(q) * k + r] = wi + idx_base;
      }
    }
  }
  if (out == nullptr) return;
  __syncthreads();
  gather_mean_row(lib_raw, n_total, d, s_top, 0, k, q_raw + static_cast<size_t>(q) * d, a0 != 0.f || !isfinite(qn), a1, a0,
                  out + static_cast<size_t>(q) * d, threadIdx.x, kCollectThreads);
}

// grid = (min(rows_c, a few CTAs per SM), items); the CTAs stride over the item's live slots (see refine_prep_kernel)
__global__ void __launch_bounds__(kCollectThreads, 2)
collect_rescore_kernel(const int* __restrict__ fb_list, const int* __restrict__ fb_count, int t_item, int rows_c,
                       const int* __restrict__ c_cnt, const int* __restrict__ c_idx, int c_cap, int k,
                       const float* __restrict__ q_raw, const float* __restrict__ q_norm,
                       const float* __restrict__ lib_raw, const float* __restrict__ lib_norm, long long n_total, int d,
                       float a1, float a0, float* __restrict__ out, float* __restrict__ top_score,
                       long long* __restrict__ top_idx, long long idx_base, int* __restrict__ fb2_list,
                       int* __restrict__ fb2_count) {
  pdl_launch_dependents();
  pdl_wait();
  const int item = blockIdx.y;
  fb_list += static_cast<size_t>(item) * t_item;
  fb2_list += static_cast<size_t>(item) * t_item;
  const int n_fb = fb_count[item];
  // uncertified queries beyond the compact matrix go straight to the exhaustive scan
  if (threadIdx.x == 0)
    for (int sl = rows_c + blockIdx.x; sl < n_fb; sl += gridDim.x) fb2_list[atomicAdd(&fb2_count[item], 1)] = fb_list[sl];
  const int live = min(n_fb, rows_c);
  for (int slot = blockIdx.x; slot < live; slot += gridDim.x) {
    collect_rescore_slot(item, slot, fb_list, rows_c, c_cnt, c_idx, c_cap, k, q_raw, q_norm, lib_raw, lib_norm, n_total, d, a1,
                         a0, out, top_score, top_idx, idx_base, fb2_list, fb2_count);
    __syncthreads();                       // the next slot reuses the shared-memory buffers
  }
}

// ------------------------------------------------------------------------------------------
// exact scan: fp64-accumulated similarities of kEQ queries x all frames of one library split,
// tiled like a GEMM on the CUDA cores: a CTA holds a [kEK][kEQ] query chunk and a [kEK][kER]
// frame chunk in shared memory as doubles (frames normalised x/|x| with IEEE division while
// staging), every thread owns a 4 query x 8 frame block of accumulators, and after the K loop
// each warp folds the tile's scores into its queries' running top-k lists.
// grid = (query groups [grid-stride], library splits); partial lists go to part_score/part_idx.
// ------------------------------------------------------------------------------------------
constexpr int kFewQueries = 16; // at most this many queries (per item): warp-per-frame kernel instead
constexpr int kEQ = 64;         // queries per CTA
constexpr int kER = 128;        // frames per tile
constexpr int kEK = 32;         // channels per staged chunk
constexpr int kEThreads = 256;
constexpr int kEScPitch = kER + 1;

__global__ void __launch_bounds__(kEThreads)
exact_partial_kernel(const float* __restrict__ q_raw, const float* __restrict__ q_norm, int t,
                     const float* __restrict__ lib_raw, const float* __restrict__ lib_norm, long long n, int d,
                     int k, const int* __restrict__ q_list, const int* __restrict__ q_count, int splits,
                     float* __restrict__ part_score, long long* __restrict__ part_idx, int t_item, int few) {
  // blockIdx.y = item: its queries are q_list[item][..] (or [item*t_item, (item+1)*t_item) when no
  // list is given) and its frames are rows [item*n, (item+1)*n) of lib_raw; t = queries PER ITEM
  pdl_launch_dependents();
  pdl_wait();
  const int item = blockIdx.y;
  if (q_list) q_list += static_cast<size_t>(item) * t_item;
  if (q_count) q_count += item;
  lib_raw += static_cast<size_t>(item) * n * d;
  lib_norm += static_cast<size_t>(item) * n;
  const long long item_row0 = static_cast<long long>(item) * n;
  const int item_q0 = item * t_item;
  const size_t item_slot0 = static_cast<size_t>(item) * ((t_item + kEQ - 1) / kEQ) * kEQ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Qd = reinterpret_cast<double*>(smem_raw);                       // [kEK][kEQ]
  double* Rd = Qd + kEK * kEQ;                                            // [kEK][kER]
  float* sc = reinterpret_cast<float*>(Rd + kEK * kER);                   // [kEQ][kEScPitch]
  long long* li = reinterpret_cast<long long*>(sc + kEQ * kEScPitch + ((kEQ * kEScPitch) & 1));   // [kEQ][k]
  float* ls = reinterpret_cast<float*>(li + kEQ * k);                     // [kEQ][k]
  __shared__ int qids[kEQ];
  __shared__ float qnrm[kEQ];
  __shared__ int lcount[kEQ];
  __shared__ int lworst[kEQ];

  const int nq = q_count ? *q_count : t;
  if (nq <= few) return;                    // exact_rows_kernel takes the small lists
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a warp covers 8 queries x 128 frames (two 4-query blocks x sixteen 8-frame blocks): groups
  // with few valid queries (sparse fallbacks) let whole warps skip the FMA loop
  const int tq = threadIdx.x >> 4;          // queries 4*tq .. 4*tq+3
  const int tr = threadIdx.x & 15;          // frames  32*(b/2) + 2*tr + (b&1), b < 8 (conflict-free 16-byte loads)
  const long long per = (n + splits - 1) / splits;
  const long long n_items = static_cast<long long>((nq + kEQ - 1) / kEQ) * splits;

  // work items = (query group, library split), handed out grid-stride: a handful of uncertified
  // queries still spreads over every SM, and an empty list costs one early exit per CTA
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int g0 = static_cast<int>(item / splits) * kEQ;
    const int split = static_cast<int>(item % splits);
    const long long r0 = per * split;
    const long long r1 = min(n, r0 + per);
    __syncthreads();
    if (threadIdx.x < kEQ) {
      const int slot = g0 + threadIdx.x;
      const int q = slot < nq ? (q_list ? q_list[slot] : item_q0 + slot) : -1;
      qids[threadIdx.x] = q;
      qnrm[threadIdx.x] = q >= 0 ? q_norm[q] : 1.f;
      lcount[threadIdx.x] = 0;
      lworst[threadIdx.x] = 0;
    }
    __syncthreads();

    for (long long rb = r0; rb < r1; rb += kER) {
      double acc[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

      for (int kc = 0; kc < d; kc += kEK) {
        // ---- stage: queries (8 values per thread) and frames (16 values per thread), as doubles ----
        {
          const int qi = threadIdx.x >> 2, part = threadIdx.x & 3;        // 64 queries x 4 parts of 8
          const int q = qids[qi];
          const float qn = qnrm[qi];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = part * 8 + h * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q >= 0 && kc + c < d) v = *reinterpret_cast<const float4*>(q_raw + static_cast<size_t>(q) * d + kc + c);
            Qd[(c + 0) * kEQ + qi] = static_cast<double>(__fdiv_rn(v.x, qn));
            Qd[(c + 1) * kEQ + qi] = static_cast<double>(__fdiv_rn(v.y, qn));
            Qd[(c + 2) * kEQ + qi] = static_cast<double>(__fdiv_rn(v.z, qn));
            Qd[(c + 3) * kEQ + qi] = static_cast<double>(__fdiv_rn(v.w, qn));
          }
          const int ri = threadIdx.x >> 1, rpart = threadIdx.x & 1;       // 128 frames x 2 parts of 16
          const long long row = rb + ri;
          const bool ok = row < r1;
          const float rn = ok ? lib_norm[row] : 1.f;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int c = rpart * 16 + h * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && kc + c < d) v = *reinterpret_cast<const float4*>(lib_raw + static_cast<size_t>(row) * d + kc + c);
            Rd[(c + 0) * kER + ri] = static_cast<double>(__fdiv_rn(v.x, rn));
            Rd[(c + 1) * kER + ri] = static_cast<double>(__fdiv_rn(v.y, rn));
            Rd[(c + 2) * kER + ri] = static_cast<double>(__fdiv_rn(v.z, rn));
            Rd[(c + 3) * kER + ri] = static_cast<double>(__fdiv_rn(v.w, rn));
          }
        }
        __syncthreads();
        // ---- 4 x 8 outer products per channel (valid queries are a prefix of the group) ----
        if (__any_sync(0xffffffffu, qids[4 * tq] >= 0)) {
#pragma unroll 4
        for (int j = 0; j < kEK; ++j) {
          const double2 q01 = *reinterpret_cast<const double2*>(Qd + j * kEQ + 4 * tq);
          const double2 q23 = *reinterpret_cast<const double2*>(Qd + j * kEQ + 4 * tq + 2);
          const double qv[4] = {q01.x, q01.y, q23.x, q23.y};
          double rv[8];
#pragma unroll
          for (int b = 0; b < 8; b += 2) {
            const double2 r2 = *reinterpret_cast<const double2*>(Rd + j * kER + 16 * b + 2 * tr);
            rv[b] = r2.x;
            rv[b + 1] = r2.y;
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = fma(qv[a], rv[b], acc[a][b]);
        }
        }
        __syncthreads();
      }
      // ---- scores of this tile ----
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b)
          sc[(4 * tq + a) * kEScPitch + 32 * (b >> 1) + 2 * tr + (b & 1)] = static_cast<float>(acc[a][b]);
      __syncthreads();
      // ---- fold into the running top-k lists: warp w owns queries 8w .. 8w+7 ----
      for (int qq = 0; qq < kEQ / 8; ++qq) {
        const int ql = warp * (kEQ / 8) + qq;
        if (qids[ql] < 0) continue;                       // warp-uniform
        float* my_s = ls + ql * k;
        long long* my_i = li + ql * k;
        for (int u = 0; u < kER / 32; ++u) {
          const int b = lane + 32 * u;
          const long long row = rb + b;
          bool pending = row < r1;
          const float s = sc[ql * kEScPitch + b];
          while (true) {
            const int cnt = lcount[ql];
            const int wpos = lworst[ql];
            const bool want = pending && (cnt < k || score_better(s, row + item_row0, my_s[wpos], my_i[wpos]));
            const unsigned m = __ballot_sync(0xffffffffu, want);
            if (m == 0) break;
            const int src = __ffs(static_cast<int>(m)) - 1;
            if (lane == src) {
              const int pos = cnt < k ? cnt : wpos;
              my_s[pos] = s;
              my_i[pos] = row + item_row0;              // global frame index
              if (cnt < k) lcount[ql] = cnt + 1;
              pending = false;
            }
            __syncwarp();
            if (lcount[ql] == k) {                        // list full: locate its worst entry
              float ws = 0.f;
              long long wi = -1;
              int wp = -1;
              for (int e = lane; e < k; e += 32) {
                if (wp < 0 || score_better(ws, wi, my_s[e], my_i[e])) {
                  ws = my_s[e];
                  wi = my_i[e];
                  wp = e;
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, ws, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, wi, o);
                const int op = __shfl_xor_sync(0xffffffffu, wp, o);
                if (op >= 0 && (wp < 0 || score_better(ws, wi, os, oi))) {
                  ws = os;
                  wi = oi;
                  wp = op;
                }
              }
              if (lane == 0) lworst[ql] = wp;
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
    }
    // ---- this split's top-k of every query of the group (sorted) ----
    for (int qq = 0; qq < kEQ / 8; ++qq) {
      const int ql = warp * (kEQ / 8) + qq;
      if (qids[ql] < 0) continue;
      // pad unused entries so the selector skips them
      for (int e = lcount[ql] + lane; e < k; e += 32) li[ql * k + e] = -1;
      __syncwarp();
      const size_t o = ((item_slot0 + g0 + ql) * splits + split) * k;
      warp_select_topk(ls + ql * k, li + ql * k, k, k, part_score + o, part_idx + o, 0, lane);
    }
  }
}

// Few queries (a handful of uncertified queries in an otherwise clean batch; tiny realtime
// batches in exact mode): the 64-query tile above would be mostly empty.  Here a CTA takes 8
// queries (normalised, as doubles, in shared memory) and one library split; every WARP owns whole
// frames: load + normalise the frame once, 8 fp64 dot products, warp-reduce, lane q keeps query q's
// running top-k.  Same arithmetic, same partial-list layout as exact_partial_kernel.
constexpr int kRQ = 8;   // queries per CTA
constexpr int kRR = 4;   // frames per warp bundle

// bytes of the query area (it is reused for the final merge of the per-lane lists)
__host__ __device__ inline size_t rows_query_bytes(int d, int k) {
  const size_t q = static_cast<size_t>(kRQ) * d * 8;
  const size_t tmp = (static_cast<size_t>(kRQ) * 8 * kRR * k * 12 + 16 + 15) / 16 * 16;
  return q > tmp ? q : tmp;
}

__global__ void __launch_bounds__(256)
exact_rows_kernel(const float* __restrict__ q_raw, const float* __restrict__ q_norm, int t,
                  const float* __restrict__ lib_raw, const float* __restrict__ lib_norm, long long n, int d,
                  int k, const int* __restrict__ q_list, const int* __restrict__ q_count, int splits,
                  float* __restrict__ part_score, long long* __restrict__ part_idx, int t_item, int few) {
  pdl_launch_dependents();
  pdl_wait();
  const int item = blockIdx.y;
  if (q_list) q_list += static_cast<size_t>(item) * t_item;
  if (q_count) q_count += item;
  const int nq = q_count ? *q_count : t;
  if (nq > few || nq <= 0) return;
  lib_raw += static_cast<size_t>(item) * n * d;
  lib_norm += static_cast<size_t>(item) * n;
  const long long item_row0 = static_cast<long long>(item) * n;
  const int item_q0 = item * t_item;
  const size_t item_slot0 = static_cast<size_t>(item) * ((t_item + kEQ - 1) / kEQ) * kEQ;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qd = reinterpret_cast<double*>(smem_raw);                        // [kRQ][d]
  const size_t q_bytes = rows_query_bytes(d, k);
  long long* lid = reinterpret_cast<long long*>(smem_raw + q_bytes);       // [8 warps][32 lanes][k]
  float* lsc = reinterpret_cast<float*>(lid + 8 * 32 * k);                 // [8 warps][32 lanes][k]
  __shared__ int qids[kRQ];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (n + splits - 1) / splits;
  const int n_groups = (nq + kRQ - 1) / kRQ;
  const long long n_items = static_cast<long long>(n_groups) * splits;

  for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int g0 = static_cast<int>(it / splits) * kRQ;
    const int split = static_cast<int>(it % splits);
    const long long r0 = per * split;
    const long long r1 = min(n, r0 + per);
    __syncthreads();
    if (threadIdx.x < kRQ) {
      const int slot = g0 + threadIdx.x;
      qids[threadIdx.x] = slot < nq ? (q_list ? q_list[slot] : item_q0 + slot) : -1;
    }
    __syncthreads();
    for (int qi = 0; qi < kRQ; ++qi) {
      const int q = qids[qi];
      const float qn = q >= 0 ? q_norm[q] : 1.f;
      for (int j = threadIdx.x; j < d; j += blockDim.x)
        qd[qi * d + j] = q >= 0 ? static_cast<double>(__fdiv_rn(q_raw[static_cast<size_t>(q) * d + j], qn)) : 0.0;
    }
    for (int e = threadIdx.x; e < 8 * 32 * k; e += blockDim.x) lid[e] = -1;
    __syncthreads();

    // lane l keeps the list of query (l & 7) over the rows whose position in a 4-row bundle is (l >> 3)
    float* my_s = lsc + (warp * 32 + lane) * k;
    long long* my_i = lid + (warp * 32 + lane) * k;
    const bool q_ok = qids[lane & (kRQ - 1)] >= 0;
    const int my_rr = lane >> 3;
    int filled = 0, worst = 0;
    constexpr int kMaxVec = 12;                         // d <= 1536 -> at most 12 float4 per lane
    for (long long rb = r0 + warp * kRR; rb < r1; rb += 8 * kRR) {
      const float* row[kRR];
      float nrm[kRR];
#pragma unroll
      for (int rr = 0; rr < kRR; ++rr) {
        const long long r = min(rb + rr, r1 - 1);       // bundle tail: re-read the last frame, never inserted
        row[rr] = lib_raw + static_cast<size_t>(r) * d;
        nrm[rr] = lib_norm[r];
      }
      // acc[rr * 8 + qi]: 32 independent fp64 chains per lane; every query double2 read from shared
      // memory feeds 4 frames (shared-memory bandwidth is what bounds this kernel)
      double acc[kRR * kRQ];
#pragma unroll
      for (int e = 0; e < kRR * kRQ; ++e) acc[e] = 0.0;
      float4 nxt[kRR];
#pragma unroll
      for (int rr = 0; rr < kRR; ++rr)
        nxt[rr] = lane * 4 < d ? *reinterpret_cast<const float4*>(row[rr] + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i) {
        const int j = lane * 4 + i * 128;
        if (i * 128 < d) {                               // warp-uniform
          double rn[kRR][4];
#pragma unroll
          for (int rr = 0; rr < kRR; ++rr) {
            rn[rr][0] = static_cast<double>(__fdiv_rn(nxt[rr].x, nrm[rr]));
            rn[rr][1] = static_cast<double>(__fdiv_rn(nxt[rr].y, nrm[rr]));
            rn[rr][2] = static_cast<double>(__fdiv_rn(nxt[rr].z, nrm[rr]));
            rn[rr][3] = static_cast<double>(__fdiv_rn(nxt[rr].w, nrm[rr]));
          }
          if (i + 1 < kMaxVec) {
            const int jn = j + 128;
#pragma unroll
            for (int rr = 0; rr < kRR; ++rr)
              nxt[rr] = jn < d ? *reinterpret_cast<const float4*>(row[rr] + jn) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (j < d) {
#pragma unroll
            for (int qi = 0; qi < kRQ; ++qi) {
              const double2 a01 = *reinterpret_cast<const double2*>(qd + qi * d + j);
              const double2 a23 = *reinterpret_cast<const double2*>(qd + qi * d + j + 2);
#pragma unroll
              for (int rr = 0; rr < kRR; ++rr) {
                double a = acc[rr * kRQ + qi];
                a = fma(a01.x, rn[rr][0], a);
                a = fma(a01.y, rn[rr][1], a);
                a = fma(a23.x, rn[rr][2], a);
                a = fma(a23.y, rn[rr][3], a);
                acc[rr * kRQ + qi] = a;
              }
            }
          }
        }
      }
      // transposing butterfly: 31 exchanges leave lane l with the full sum of entry l (= rr * 8 + qi)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int e = 0; e < o; ++e) {
          const double keep = hi ? acc[e + o] : acc[e];
          const double send = hi ? acc[e] : acc[e + o];
          acc[e] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      const float mine = static_cast<float>(acc[0]);
      const long long r = rb + my_rr;
      if (q_ok && r < r1) {
        const long long gi = r + item_row0;             // global frame index
        if (filled < k) {
          my_s[filled] = mine;
          my_i[filled] = gi;
          ++filled;
          if (filled == k) {
            worst = 0;
            for (int e = 1; e < k; ++e)
              if (score_better(my_s[worst], my_i[worst], my_s[e], my_i[e])) worst = e;
          }
        } else if (score_better(mine, gi, my_s[worst], my_i[worst])) {
          my_s[worst] = mine;
          my_i[worst] = gi;
          worst = 0;
          for (int e = 1; e < k; ++e)
            if (score_better(my_s[worst], my_i[worst], my_s[e], my_i[e])) worst = e;
        }
      }
    }
    __syncthreads();
    // merge the 8 x kRR per-lane lists of each query: warp w finishes query w.  The lists are copied
    // next to each other into the (now free) query area and go through the selector.
    {
      const int qi = warp;
      constexpr int kLists = 8 * kRR;
      if (qi < kRQ && qids[qi] >= 0) {
        float* tmp_s = reinterpret_cast<float*>(qd) + qi * kLists * k;
        long long* tmp_i =
            reinterpret_cast<long long*>(reinterpret_cast<float*>(qd) + kRQ * kLists * k + (kRQ * kLists * k & 1)) +
            qi * kLists * k;
        for (int e = lane; e < kLists * k; e += 32) {
          const int l = e / k, j = e % k;                // l = warp' * kRR + rr'
          const int src = ((l / kRR) * 32 + (l % kRR) * kRQ + qi) * k + j;
          tmp_s[e] = lsc[src];
          tmp_i[e] = lid[src];
        }
        __syncwarp();
        const size_t o = ((item_slot0 + g0 + qi) * splits + split) * k;
        warp_select_topk(tmp_s, tmp_i, kLists * k, k, part_score + o, part_idx + o, 0, lane);
      }
    }
  }
}

__global__ void __launch_bounds__(128)
exact_final_kernel(int t, int k, const int* __restrict__ q_list, const int* __restrict__ q_count, int splits,
                   const float* __restrict__ part_score, const long long* __restrict__ part_idx,
                   long long idx_base, float* __restrict__ top_score, long long* __restrict__ top_idx,
                   const float* __restrict__ lib_raw, long long n, int d, const float* __restrict__ q_raw,
                   float a1, float a0, float* __restrict__ out, int t_item) {
  // blockIdx.y = item; t = queries PER ITEM; n = frames of ALL items (gather reads global indices)
  pdl_wait();
  const int item = blockIdx.y;
  const int nq = q_count ? q_count[item] : t;
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= nq) return;
  const int lane = threadIdx.x & 31;
  const int q = q_list ? q_list[static_cast<size_t>(item) * t_item + slot] : item * t_item + slot;
  const size_t item_slot0 = static_cast<size_t>(item) * ((t_item + kEQ - 1) / kEQ) * kEQ;
  const size_t o = (item_slot0 + slot) * splits * k;
  warp_select_topk(part_score + o, part_idx + o, splits * k, k, top_score + static_cast<size_t>(q) * k,
                   top_idx + static_cast<size_t>(q) * k, idx_base, lane);
  if (out == nullptr) return;
  __syncwarp();
  __threadfence_block();
  // the warp that selected the top-k of this query also gathers it (top_idx written by lane 0)
  gather_mean_row(lib_raw, n, d, top_idx + static_cast<size_t>(q) * k, idx_base, k, q_raw + static_cast<size_t>(q) * d,
                  true, a1, a0, out + static_cast<size_t>(q) * d, lane, 32);
}

__global__ void __launch_bounds__(128)
merge_kernel(const float* __restrict__ scores, const long long* __restrict__ idx, int ranks, int t, int k,
             float* __restrict__ top_score, long long* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + warp;
  long long* ci = reinterpret_cast<long long*>(smem_raw) + warp * ranks * k;
  float* cs = reinterpret_cast<float*>(reinterpret_cast<long long*>(smem_raw) + 4 * ranks * k) + warp * ranks * k;
  if (q >= t) return;
  for (int e = lane; e < ranks * k; e += 32) {
    const int r = e / k, j = e % k;
    cs[e] = scores[(static_cast<size_t>(r) * t + q) * k + j];
    ci[e] = idx[(static_cast<size_t>(r) * t + q) * k + j];
  }
  __syncwarp();
  warp_select_topk(cs, ci, ranks * k, k, top_score + static_cast<size_t>(q) * k, top_idx + static_cast<size_t>(q) * k, 0, lane);
}

int exact_splits(int t, long long n, int k) {
  // library splits of ~4096 frames (at most 256) so that even a handful of uncertified queries is
  // spread over the whole GPU; the partial-list workspace (slots * splits * k * 12 B) is capped
  long long s = (n + 4095) / 4096;
  if (s > 256) s = 256;
  const long long groups = (static_cast<long long>(t) + kEQ - 1) / kEQ;
  const long long s_fill = (2 * 148 + groups - 1) / groups;    // few query groups: split finer to fill the GPU
  if (s < s_fill) s = s_fill;
  const long long s_max = (n + kER - 1) / kER;                 // at least one 128-frame tile per split
  if (s > s_max) s = s_max;
  if (s < 1) s = 1;
  const long long slots = groups * kEQ;
  while (s > 1 && slots * s * k * 12 > (384ll << 20)) s /= 2;
  return static_cast<int>(s);
}

}  // namespace
}  // namespace alive

extern "C" int alive_knn_prune(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists,
                               int32_t k, int32_t d, const float* q_err, const float* q_norm, const uint32_t* lib_stats,
                               int32_t r_max, int32_t* sel_idx, int32_t* sel_n, int32_t* fb_list,
                               int32_t* fb_count, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(cand_score && cand_idx && q_err && q_norm && lib_stats && sel_idx && sel_n && fb_list && fb_count,
                "alive_knn_prune: NULL argument");
  ALIVE_REQUIRE(t >= 1 && lists >= 1 && d >= 16 && d <= 1536, "alive_knn_prune: bad sizes");
  ALIVE_REQUIRE(k >= 1 && k <= kListLen, "alive_knn_prune: k must be in [1,%d] for the screened path (got %d)", kListLen, k);
  ALIVE_REQUIRE(r_max >= k && r_max <= kMaxRMax, "alive_knn_prune: r_max must be in [k,%d]", kMaxRMax);
  ALIVE_CHECK_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(int32_t), as_stream(stream)));
  const int wpb = 8;
  prune_kernel<<<(t + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(
      cand_score, cand_idx, t, lists, k, q_err, q_norm, lib_stats, r_max, sel_idx, sel_n, fb_list, fb_count, d);
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int alive_knn_rescore(const float* q_raw, const float* q_norm, int32_t t, const float* lib_raw,
                                 const float* lib_norm, int32_t d, const int32_t* sel_idx, const int32_t* sel_n,
                                 int32_t r_max, int32_t k, int64_t idx_base, float* top_score, int64_t* top_idx,
                                 alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(q_raw && q_norm && lib_raw && lib_norm && sel_idx && sel_n && top_score && top_idx,
                "alive_knn_rescore: NULL argument");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 1536, "alive_knn_rescore: d must be a multiple of 4, <= 1536 (got %d)", d);
  ALIVE_REQUIRE(k >= 1 && k <= r_max && r_max <= kMaxRMax, "alive_knn_rescore: need 1 <= k <= r_max <= %d", kMaxRMax);
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(lib_raw) | reinterpret_cast<uintptr_t>(q_raw)) & 15) == 0,
                "alive_knn_rescore: raw buffers must be 16-byte aligned");
  if (t <= 0) return 0;
  const size_t smem = static_cast<size_t>(d) * 4 + static_cast<size_t>(r_max) * 12 + 16;
  rescore_kernel<<<t, 128, smem, as_stream(stream)>>>(q_raw, q_norm, t, lib_raw, lib_norm, d, sel_idx, sel_n, r_max, k,
                                                      idx_base, top_score, reinterpret_cast<long long*>(top_idx));
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t alive_knn_exact_workspace_bytes(int32_t t, int64_t n, int32_t k, int32_t items) {
  using namespace alive;
  if (t < 1 || k < 1 || items < 1) return 0;
  const int s = exact_splits(t, n, k);
  const size_t groups = (static_cast<size_t>(t / items) + kEQ - 1) / kEQ;      // per item
  return static_cast<size_t>(items) * groups * kEQ * s * k * 12 + 256;
}

extern "C" int alive_knn_exact(const float* q_raw, const float* q_norm, int32_t t, const float* lib_raw,
                               const float* lib_norm, int64_t n, int32_t d, int32_t k, const int32_t* q_list,
                               const int32_t* q_count, int64_t idx_base, void* workspace, float* top_score,
                               int64_t* top_idx, float alpha, float* out, int32_t items, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(q_raw && q_norm && lib_raw && lib_norm && workspace && top_score && top_idx,
                "alive_knn_exact: NULL argument");
  ALIVE_REQUIRE((q_list == nullptr) == (q_count == nullptr), "alive_knn_exact: q_list and q_count go together");
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 1536, "alive_knn_exact: d must be a multiple of 4, <= 1536");
  ALIVE_REQUIRE(k >= 1 && k <= kMaxK, "alive_knn_exact: k must be in [1,%d]", kMaxK);
  ALIVE_REQUIRE(n >= k, "selected index k out of range");
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(lib_raw) | reinterpret_cast<uintptr_t>(q_raw)) & 15) == 0,
                "alive_knn_exact: raw buffers must be 16-byte aligned");
  if (t <= 0) return 0;
  ALIVE_REQUIRE(items >= 1 && t % items == 0, "alive_knn_exact: t must be a multiple of items");
  ALIVE_REQUIRE(items == 1 || idx_base == 0, "alive_knn_exact: batched items cannot be row-sharded");
  const int t_item = t / items;
  const int splits = exact_splits(t, n, k);
  const size_t groups = (static_cast<size_t>(t_item) + kEQ - 1) / kEQ;          // per item
  // workspace layout: part_idx [items*groups*kEQ*splits*k] int64, then part_score float
  long long* part_idx = reinterpret_cast<long long*>(workspace);
  float* part_score = reinterpret_cast<float*>(part_idx + static_cast<size_t>(items) * groups * kEQ * splits * k);
  const size_t smem = static_cast<size_t>(kEK) * (kEQ + kER) * 8 + (static_cast<size_t>(kEQ) * kEScPitch + 1) * 4 +
                      static_cast<size_t>(kEQ) * k * 12 + 16;
  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA(cudaFuncSetAttribute(exact_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      ALIVE_CHECK_CUDA(cudaFuncSetAttribute(exact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  ALIVE_REQUIRE(smem <= 160 * 1024, "alive_knn_exact: shared memory budget exceeded");
  // The (group, split) items are reached grid-stride.  Scanning EVERY query (no list: mode = exact, k > 8): ~4 waves of
  // CTAs - one wave measured 22 % slower (13.2 instead of 10.8 ms at T = 1000 x N = 100k).  As the tail of the screened
  // pipeline (a list of uncertified queries, normally empty): ONE wave, so that an idle launch releases its dependents
  // before its own wait returns and the idle chain drains without launch gaps (cfg2 p50 76.4 -> 74.7 us).
  const size_t work = groups * static_cast<size_t>(splits);
  const int waves = q_list != nullptr ? 1 : 4;
  size_t gx = (static_cast<size_t>(waves) * 148 + items - 1) / items;
  if (gx > work) gx = work;
  if (gx < 1) gx = 1;
  dim3 pgrid(static_cast<unsigned>(gx), static_cast<unsigned>(items));
  // lists of at most `few` queries per item go to the warp-per-frame kernel (chosen on the device from
  // q_count); its per-lane lists only fit in shared memory for moderate k
  const size_t rsmem = rows_query_bytes(d, k) + static_cast<size_t>(8) * 32 * k * 12;
  const int few = rsmem <= 200 * 1024 ? kFewQueries : 0;
  ALIVE_CHECK_CUDA(launch_chained(exact_partial_kernel, pgrid, dim3(kEThreads), smem, as_stream(stream), q_raw, q_norm, t_item,
                                  lib_raw, lib_norm, static_cast<long long>(n), d, k, q_list, q_count, splits, part_score,
                                  part_idx, t_item, few));
  if (few > 0) {
    const size_t rwork = static_cast<size_t>((few + kRQ - 1) / kRQ) * splits;
    size_t rgx = (static_cast<size_t>(waves) * 148 + items - 1) / items;
    if (rgx > rwork) rgx = rwork;
    if (rgx < 1) rgx = 1;
    dim3 rgrid(static_cast<unsigned>(rgx), static_cast<unsigned>(items));
    ALIVE_CHECK_CUDA(launch_chained(exact_rows_kernel, rgrid, dim3(256), rsmem, as_stream(stream), q_raw, q_norm, t_item,
                                    lib_raw, lib_norm, static_cast<long long>(n), d, k, q_list, q_count, splits, part_score,
                                    part_idx, t_item, few));
  }
  ALIVE_REQUIRE(out == nullptr || ((reinterpret_cast<uintptr_t>(out) & 15) == 0 && idx_base == 0),
                "alive_knn_exact: gather needs a 16-byte aligned `out` and an unsharded library");
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  dim3 fgrid(static_cast<unsigned>((t_item + 3) / 4), static_cast<unsigned>(items));
  ALIVE_CHECK_CUDA(launch_chained(exact_final_kernel, fgrid, dim3(128), 0, as_stream(stream), t_item, k, q_list, q_count,
                                  splits, part_score, part_idx, static_cast<long long>(idx_base), top_score,
                                  reinterpret_cast<long long*>(top_idx), lib_raw, static_cast<long long>(n * items), d,
                                  q_raw, a1, alpha, out, t_item));
  return 0;
}

extern "C" int alive_knn_merge(const float* scores, const int64_t* idx, int32_t ranks, int32_t t, int32_t k,
                               float* top_score, int64_t* top_idx, alive_stream_t stream) {
  using namespace alive;
  ALIVE_REQUIRE(scores && idx && top_score && top_idx, "alive_knn_merge: NULL argument");
  ALIVE_REQUIRE(ranks >= 1 && ranks <= 64 && k >= 1 && k <= kMaxK, "alive_knn_merge: bad sizes");
  if (t <= 0) return 0;
  const size_t smem = static_cast<size_t>(4) * ranks * k * 12;
  ALIVE_REQUIRE(smem <= 48 * 1024, "alive_knn_merge: ranks*k too large");
  merge_kernel<<<(t + 3) / 4, 128, smem, as_stream(stream)>>>(scores, reinterpret_cast<const long long*>(idx), ranks, t, k,
                                                              top_score, reinterpret_cast<long long*>(top_idx));
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace alive {
int finish_impl(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists, int32_t k,
                const float* q_raw, const float* q_norm, const float* q_err, const float* lib_raw,
                const float* lib_norm, const uint32_t* lib_stats, int64_t n, int32_t d, int32_t r_max,
                int64_t idx_base, float alpha, float* out, float* top_score, int64_t* top_idx, int32_t* sel_n,
                int32_t* fb_list, int32_t* fb_count, int32_t items, int zero_counts, const uint16_t* q_packed,
                uint16_t* qc_packed, float* c_cut, int32_t* c_cnt, int32_t rows_c, alive_stream_t stream) {
  ALIVE_REQUIRE(cand_score && cand_idx && q_raw && q_norm && q_err && lib_raw && lib_norm && lib_stats && top_score &&
                    top_idx && sel_n && fb_list && fb_count,
                "alive_knn_finish: NULL argument");
  ALIVE_REQUIRE(t >= 1 && lists >= 1, "alive_knn_finish: bad sizes");
  ALIVE_REQUIRE(items >= 1 && t % items == 0, "alive_knn_finish: t must be a multiple of items");
  ALIVE_REQUIRE(k >= 1 && k <= kListLen, "alive_knn_finish: k must be in [1,%d] for the screened path (got %d)", kListLen, k);
  ALIVE_REQUIRE(r_max >= k && r_max <= kMaxRMax, "alive_knn_finish: r_max must be in [k,%d]", kMaxRMax);
  ALIVE_REQUIRE(d % 4 == 0 && d >= 4 && d <= 1536, "alive_knn_finish: d must be a multiple of 4, <= 1536 (got %d)", d);
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(lib_raw) | reinterpret_cast<uintptr_t>(q_raw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "alive_knn_finish: raw/out buffers must be 16-byte aligned");
  ALIVE_REQUIRE(out == nullptr || idx_base == 0, "alive_knn_finish: gather needs an unsharded library");
  if (zero_counts) ALIVE_CHECK_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(int32_t) * items, as_stream(stream)));
  const int entries = lists * kListLen;
  const int staged = entries <= kFinishMaxStagedEntries ? 1 : 0;
  const size_t smem = static_cast<size_t>(d) * 4 + static_cast<size_t>(r_max) * 16 + 16 +
                      (staged ? static_cast<size_t>(entries) * 8 : 0);
  // 0 = by batch size: 128 threads x 8 CTAs/SM keeps more queries in flight once the batch spans many
  // waves (measured at T = 10k: 360 -> 286 us; 64 x 16 from 32k queries: another 0.3 ms of 33.5 at T = 64k, nothing at
  // 10k, slower at 1k), 256 x 4 in between, 512 / 1024 threads when there are
  // fewer queries than SMs can hold; ALIVE_KNN_FINISH_THREADS=128|256|512|1024 forces one
  static const int variant = getenv("ALIVE_KNN_FINISH_THREADS") ? atoi(getenv("ALIVE_KNN_FINISH_THREADS")) : 0;
  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(finish_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(finish_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(finish_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(finish_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(finish_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  ALIVE_REQUIRE(smem <= 100 * 1024, "alive_knn_finish: shared memory budget exceeded");
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  ALIVE_REQUIRE(qc_packed == nullptr || (q_packed && c_cut && c_cnt && rows_c >= 1 && d % 8 == 0),
                "alive_knn_finish: bad collect stage");
  CollectStage cs{q_packed, qc_packed, c_cut, c_cnt, rows_c};
#define ALIVE_LAUNCH_FINISH(TH, B)                                                                                   \
  ALIVE_CHECK_CUDA(launch_chained(finish_kernel<TH, B>, dim3(t), dim3(TH), smem, as_stream(stream), cand_score, cand_idx, t,   \
                                  lists, k, q_raw, q_norm, q_err, lib_raw, lib_norm, lib_stats, static_cast<long long>(n), d,   \
                                  r_max, static_cast<long long>(idx_base), a1, alpha, out, top_score,                           \
                                  reinterpret_cast<long long*>(top_idx), sel_n, fb_list, fb_count, staged, t / items, cs))
  // a batch that does not fill the GPU is pure latency: give every query a whole SM's worth of warps
  // (one survivor frame per warp in flight -> the rescoring is a single DRAM round trip)
  const int threads = variant == 64 || variant == 128 || variant == 256 || variant == 512 || variant == 1024
                          ? variant
                          : (t >= 32768 ? 64 : t >= 4096 ? 128 : t > 296 ? 256 : t > 148 ? 512 : 1024);
  if (threads == 64) ALIVE_LAUNCH_FINISH(64, 16);
  else if (threads == 128) ALIVE_LAUNCH_FINISH(128, 8);
  else if (threads == 256) ALIVE_LAUNCH_FINISH(256, 4);
  else if (threads == 512) ALIVE_LAUNCH_FINISH(512, 2);
  else ALIVE_LAUNCH_FINISH(1024, 1);
#undef ALIVE_LAUNCH_FINISH
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace alive

extern "C" int alive_knn_finish(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists, int32_t k,
                                const float* q_raw, const float* q_norm, const float* q_err, const float* lib_raw,
                                const float* lib_norm, const uint32_t* lib_stats, int64_t n, int32_t d, int32_t r_max,
                                int64_t idx_base, float alpha, float* out, float* top_score, int64_t* top_idx,
                                int32_t* sel_n, int32_t* fb_list, int32_t* fb_count, int32_t items,
                                alive_stream_t stream) {
  return alive::finish_impl(cand_score, cand_idx, t, lists, k, q_raw, q_norm, q_err, lib_raw, lib_norm, lib_stats, n, d, r_max,
                            idx_base, alpha, out, top_score, top_idx, sel_n, fb_list, fb_count, items, 1, nullptr, nullptr,
                            nullptr, nullptr, 0, stream);
}

namespace alive {
int collect_rescore_impl(const int32_t* fb_list, const int32_t* fb_count, int32_t t_item, int32_t items, int32_t rows_c,
                         const int32_t* c_cnt, const int32_t* c_idx, int32_t c_cap, int32_t k, const float* q_raw,
                         const float* q_norm, const float* lib_raw, const float* lib_norm, int64_t n_total, int32_t d,
                         float alpha, float* out, float* top_score, int64_t* top_idx, int64_t idx_base, int32_t* fb2_list,
                         int32_t* fb2_count, alive_stream_t stream) {
  const size_t smem = static_cast<size_t>(d) * 4 + static_cast<size_t>(c_cap) * 4;
  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA(cudaFuncSetAttribute(collect_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  ALIVE_REQUIRE(smem <= 64 * 1024, "collect pass: candidate buffer too large for shared memory");
  const float a1 = static_cast<float>(1.0 - static_cast<double>(alpha));
  ALIVE_CHECK_CUDA(launch_chained(collect_rescore_kernel, dim3(rows_c < 296 ? rows_c : 296, items), dim3(kCollectThreads), smem, as_stream(stream),
                                  fb_list, fb_count, t_item, rows_c, c_cnt, c_idx, c_cap, k, q_raw, q_norm, lib_raw, lib_norm,
                                  static_cast<long long>(n_total), d, a1, alpha, out, top_score,
                                  reinterpret_cast<long long*>(top_idx), static_cast<long long>(idx_base), fb2_list,
                                  fb2_count));
  return 0;
}

int refine_prep_impl(const float* cand_score, const int32_t* cand_idx, int32_t lists, int32_t k, const int32_t* fb_list,
                     const int32_t* fb_count, int32_t t_item, int32_t items, int32_t rows_c, const float* q_raw,
                     const float* q_norm, const float* q_err, const float* q_err2, const uint16_t* q_lo, uint16_t* qc_lo,
                     const float* lib_raw, const float* lib_norm, const uint32_t* lib_stats, int32_t d, float* c_cut,
                     alive_stream_t stream) {
  ALIVE_REQUIRE(cand_score && cand_idx && fb_list && fb_count && q_raw && q_norm && q_err && q_err2 && q_lo && qc_lo &&
                    lib_raw && lib_norm && lib_stats && c_cut,
                "collect pass (refine): NULL argument");
  ALIVE_REQUIRE(d % 8 == 0 && k >= 1 && k <= kListLen, "collect pass (refine): bad sizes");
  const size_t smem = static_cast<size_t>(d) * 4 + static_cast<size_t>(2 * kPrepR) * 8;
  ALIVE_CHECK_CUDA(launch_chained(refine_prep_kernel, dim3(rows_c < 592 ? rows_c : 592, items), dim3(kPrepThreads), smem, as_stream(stream),
                                  cand_score, cand_idx, lists, k, fb_list, fb_count, t_item, rows_c, q_raw, q_norm, q_err,
                                  q_err2, q_lo, qc_lo, lib_raw, lib_norm, lib_stats, d, c_cut));
  return 0;
}
}  // namespace alive

#ifdef ALIVE_FINISH_TIMING
extern "C" int alive_knn_debug_finish_times(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, alive::g_finish_t, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}
#endif
