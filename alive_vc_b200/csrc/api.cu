// Error plumbing and version of the alive_knn C ABI.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>

#include "common.cuh"

namespace alive {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace alive

extern "C" const char* alive_knn_last_error(void) { return alive::g_error; }
extern "C" int alive_knn_abi_version(void) { return ALIVE_KNN_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// One-call pipeline: pack queries -> search -> prune -> rescore -> exact (uncertified) -> gather.
// All launches go to `stream`, all scratch lives in one caller-provided workspace; nothing
// synchronises the host, so the whole call is CUDA-graph capturable.
// ---------------------------------------------------------------------------------------------
namespace alive {
namespace {

constexpr int kOffQRaw = 0, kOffQNorm = 1, kOffQPacked = 2, kOffQErr = 3, kOffCandScore = 4, kOffCandIdx = 5,
              kOffCollect = 6, kOffSelN = 7, kOffFbList = 8, kOffFbCount = 9, kOffExact = 10, kOffTotal = 11,
              kOffQLo = 12, kOffQErr2 = 13, kOffSlots = 14;

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// Early result notification (alive_knn_arm_notify): a one-thread kernel right behind finish_kernel tells the HOST that
// the certified results are complete - and whether any query was left to the fallback chain - so a latency-critical
// caller whose results land in mapped host memory need not wait for the (normally idle) fallback launches behind it.
struct NotifyArm {
  unsigned long long* host_flag = nullptr;
  unsigned long long* dev_counter = nullptr;
};
thread_local NotifyArm g_notify;

__global__ void notify_kernel(const int32_t* __restrict__ fb_count, int items, unsigned long long* dev_counter,
                              volatile unsigned long long* host_flag) {
  pdl_wait();                              // finish_kernel (or the exact scan) has completed, its writes are performed
  pdl_launch_dependents();
  int any = 0;
  if (fb_count != nullptr)
    for (int i = 0; i < items; ++i) any |= fb_count[i];
  const unsigned long long c = *dev_counter + 1ull;
  *dev_counter = c;
  __threadfence_system();                  // results (mapped host memory) before the flag, system-wide
  *host_flag = (c << 1) | (any != 0 ? 1ull : 0ull);
}

int launch_notify(const NotifyArm& arm, const int32_t* fb_count, int items, alive_stream_t stream) {
  if (arm.host_flag == nullptr) return 0;
  ALIVE_CHECK_CUDA(launch_chained(notify_kernel, dim3(1), dim3(1), 0, as_stream(stream), fb_count, items, arm.dev_counter,
                                  static_cast<volatile unsigned long long*>(arm.host_flag)));
  return 0;
}

int resolve_mode(int mode, int64_t n, int d, int k) {
  (void)n;   // the screen handles any library size; the exact scan is for k > 8 or odd feature dims
  if (mode == 0) return (k > ALIVE_KNN_LIST_LEN || d % 64 != 0) ? 2 : 1;
  return mode;
}

// Second screen pass for uncertified queries (the first screen may have been the tiled or the skinny kernel - the
// collect pass always runs the tiled one on its own compact query matrix; enough work for the exhaustive scan to
// hurt): per item at most rows_c queries get a slot and a buffer of kCollectCap candidate frames each.
// Area layout: fb2_list [rows] | qc_packed [items*rows_c, d] bf16 | qc_lo [items*rows_c, d] bf16 (the second plane, used
// when the library carries one) | cut [items*rows_c] | cnt [items*rows_c] | idx [items*rows_c, cap];
// the fb2 counters are the `items` words after the per-item fallback counters.
constexpr int kCollectRows = 8192;
constexpr int kCollectCap = 2048;
struct CollectLayout {
  bool on;
  int rows_c;
  alive_knn_plan_t plan;
  size_t fb2_list, qc, qc_lo, cut, cnt, idx, bytes;
};

int collect_layout(int32_t rows, int64_t n, int32_t d, int32_t num_sms, int mode, int32_t items,
                   const alive_knn_plan_t& screen_plan, CollectLayout* cl) {
  static const bool enabled = !(getenv("ALIVE_KNN_COLLECT") && atoi(getenv("ALIVE_KNN_COLLECT")) == 0);
  cl->on = enabled && mode == 1 && (screen_plan.kernel == 0 || screen_plan.kernel == 1) && d % 64 == 0 &&
           static_cast<double>(rows) * static_cast<double>(n) >= 16777216.0;
  cl->bytes = 0;
  if (!cl->on) return 0;
  const int t_item = rows / items;
  long long rc_rows = t_item < kCollectRows ? t_item : kCollectRows;
  const long long budget = (1ll << 30) / (static_cast<long long>(items) * kCollectCap * 4);     // candidate buffers <= 1 GiB
  if (rc_rows > budget) rc_rows = budget;
  if (rc_rows < 1) {
    cl->on = false;
    return 0;
  }
  cl->rows_c = static_cast<int>(rc_rows);
  int rc = alive_knn_plan_batched(items, cl->rows_c, n, d, num_sms, cl->rows_c > ALIVE_KNN_TILE_M ? 2 : 1, &cl->plan);
  if (rc) return rc;
  size_t cur = 0;
  auto take = [&](size_t bytes) {
    const size_t at = cur;
    cur += align256(bytes);
    return at;
  };
  const size_t slots = static_cast<size_t>(items) * cl->rows_c;
  cl->fb2_list = take(static_cast<size_t>(rows) * 4);
  cl->qc = take(slots * d * 2);
  cl->qc_lo = take(slots * d * 2);
  cl->cut = take(slots * 4);
  cl->cnt = take(slots * 4);
  cl->idx = take(slots * kCollectCap * 4);
  cl->bytes = cur;
  return 0;
}

int layout(int32_t rows, int64_t n, int32_t d, int32_t k, int32_t r_max, int32_t num_sms, int32_t variant, int mode,
           int32_t items, alive_knn_plan_t* plan, int64_t* off, CollectLayout* cl) {
  size_t cur = 0;
  auto take = [&](int slot, size_t bytes) {
    off[slot] = static_cast<int64_t>(cur);
    cur += align256(bytes);
  };
  take(kOffQRaw, static_cast<size_t>(rows) * d * 4);
  take(kOffQNorm, static_cast<size_t>(rows) * 4);
  take(kOffQPacked, static_cast<size_t>(rows) * d * 2);
  take(kOffQErr, static_cast<size_t>(rows) * 4);
  take(kOffQLo, static_cast<size_t>(rows) * d * 2);
  take(kOffQErr2, static_cast<size_t>(rows) * 4);
  size_t lists = 0;
  if (mode == 1) {
    int rc = alive_knn_plan_batched(items, rows / items, n, d, num_sms, variant, plan);
    if (rc) return rc;
    lists = static_cast<size_t>(plan->lists);
  }
  take(kOffCandScore, static_cast<size_t>(rows) * lists * ALIVE_KNN_LIST_LEN * 4);
  take(kOffCandIdx, static_cast<size_t>(rows) * lists * ALIVE_KNN_LIST_LEN * 4);
  {
    CollectLayout local;
    CollectLayout* c = cl ? cl : &local;
    if (mode == 1) {
      int rc = collect_layout(rows, n, d, num_sms, mode, items, *plan, c);
      if (rc) return rc;
    } else {
      c->on = false;
      c->bytes = 0;
    }
    take(kOffCollect, c->bytes);
  }
  take(kOffSelN, static_cast<size_t>(rows) * 4);
  take(kOffFbList, static_cast<size_t>(rows) * 4);
  take(kOffFbCount, static_cast<size_t>(2 * items) * 4);   // per-item counters: uncertified after the screen, after the collect pass
  take(kOffExact, alive_knn_exact_workspace_bytes(rows, n, k, items));
  off[kOffTotal] = static_cast<int64_t>(cur);
  return 0;
}

}  // namespace
}  // namespace alive

extern "C" int alive_knn_match_layout(int32_t rows, int64_t n, int32_t d, int32_t k, int32_t r_max, int32_t mode,
                                      int32_t num_sms, int32_t variant, int32_t items, int64_t* offsets14) {
  using namespace alive;
  ALIVE_REQUIRE(offsets14 != nullptr, "alive_knn_match_layout: offsets is NULL");
  ALIVE_REQUIRE(rows >= 1 && n >= 1 && k >= 1 && k <= ALIVE_KNN_MAX_K, "alive_knn_match_layout: bad sizes");
  ALIVE_REQUIRE(items >= 1 && rows % items == 0, "alive_knn_match_layout: rows must be a multiple of items");
  ALIVE_REQUIRE(d >= 4 && d % 4 == 0 && d <= 1536, "alive_knn_match: d must be a multiple of 4, <= 1536 (got %d)", d);
  alive_knn_plan_t plan;
  return layout(rows, n, d, k, r_max, num_sms, variant, resolve_mode(mode, n, d, k), items, &plan, offsets14, nullptr);
}

namespace alive {
namespace {
// queries that arrive already packed (alive_knn_match_packed): the buffers K1 would have written
struct PackedQueries {
  const float* raw;
  const float* norm;
  const uint16_t* packed;
  const float* err;
  const uint16_t* lo;      // second plane + its error norms (both or neither; without them the collect pass is not refined)
  const float* err2;
};
int match_impl(const float* source, const PackedQueries* pq, int32_t batch, int32_t t, int64_t stride_b, int64_t stride_t,
               int64_t stride_d, const alive_knn_library_t* lib, int32_t k, float alpha, int32_t r_max, int32_t mode,
               int32_t num_sms, int32_t variant, void* workspace, size_t workspace_bytes, float* out, int64_t* top_idx,
               float* top_score, void* ev_search_start, void* ev_search_stop, alive_stream_t stream, int phase = 0);
}  // namespace
}  // namespace alive

extern "C" int alive_knn_match(const float* source, int32_t batch, int32_t t, int64_t stride_b, int64_t stride_t,
                               int64_t stride_d, const alive_knn_library_t* lib, int32_t k, float alpha,
                               int32_t r_max, int32_t mode, int32_t num_sms, int32_t variant, void* workspace,
                               size_t workspace_bytes, float* out, int64_t* top_idx, float* top_score,
                               void* ev_search_start, void* ev_search_stop, alive_stream_t stream) {
  ALIVE_REQUIRE(source != nullptr, "alive_knn_match: NULL argument");
  return alive::match_impl(source, nullptr, batch, t, stride_b, stride_t, stride_d, lib, k, alpha, r_max, mode, num_sms,
                           variant, workspace, workspace_bytes, out, top_idx, top_score, ev_search_start, ev_search_stop,
                           stream);
}

extern "C" int alive_knn_match_packed(const float* q_raw, const float* q_norm, const uint16_t* q_packed,
                                      const float* q_err, const uint16_t* q_lo, const float* q_err2, int32_t batch,
                                      int32_t t, const alive_knn_library_t* lib, int32_t k, float alpha, int32_t r_max,
                                      int32_t mode, int32_t num_sms, int32_t variant, void* workspace,
                                      size_t workspace_bytes, float* out, int64_t* top_idx, float* top_score,
                                      alive_stream_t stream) {
  ALIVE_REQUIRE(q_raw && q_norm && q_packed && q_err, "alive_knn_match_packed: NULL argument");
  ALIVE_REQUIRE((q_lo == nullptr) == (q_err2 == nullptr), "alive_knn_match_packed: q_lo and q_err2 go together");
  ALIVE_REQUIRE(((reinterpret_cast<uintptr_t>(q_raw) | reinterpret_cast<uintptr_t>(q_packed) | reinterpret_cast<uintptr_t>(q_lo)) & 15) == 0,
                "alive_knn_match_packed: q_raw, q_packed and q_lo must be 16-byte aligned");
  const alive::PackedQueries pq{q_raw, q_norm, q_packed, q_err, q_lo, q_err2};
  return alive::match_impl(nullptr, &pq, batch, t, 0, 0, 0, lib, k, alpha, r_max, mode, num_sms, variant, workspace,
                           workspace_bytes, out, top_idx, top_score, nullptr, nullptr, stream);
}

extern "C" int alive_knn_match_fallback(const float* q_raw, const float* q_norm, const uint16_t* q_packed,
                                        const float* q_err, const uint16_t* q_lo, const float* q_err2, int32_t batch,
                                        int32_t t, const alive_knn_library_t* lib, int32_t k, float alpha,
                                        int32_t r_max, int32_t mode, int32_t num_sms, int32_t variant, void* workspace,
                                        size_t workspace_bytes, float* out, int64_t* top_idx, float* top_score,
                                        alive_stream_t stream) {
  if (q_raw == nullptr)   // the front half was alive_knn_match: its packed queries are in the workspace
    return alive::match_impl(nullptr, nullptr, batch, t, 0, 0, 0, lib, k, alpha, r_max, mode, num_sms, variant, workspace,
                             workspace_bytes, out, top_idx, top_score, nullptr, nullptr, stream, 2);
  ALIVE_REQUIRE(q_norm && q_packed && q_err, "alive_knn_match_fallback: NULL argument");
  ALIVE_REQUIRE((q_lo == nullptr) == (q_err2 == nullptr), "alive_knn_match_fallback: q_lo and q_err2 go together");
  alive::PackedQueries pq{q_raw, q_norm, q_packed, q_err, q_lo, q_err2};
  return alive::match_impl(nullptr, &pq, batch, t, 0, 0, 0, lib, k, alpha, r_max, mode, num_sms, variant, workspace,
                           workspace_bytes, out, top_idx, top_score, nullptr, nullptr, stream, 2);
}

namespace alive {
namespace {
int match_impl(const float* source, const PackedQueries* pq, int32_t batch, int32_t t, int64_t stride_b, int64_t stride_t,
               int64_t stride_d, const alive_knn_library_t* lib, int32_t k, float alpha, int32_t r_max, int32_t mode,
               int32_t num_sms, int32_t variant, void* workspace, size_t workspace_bytes, float* out, int64_t* top_idx,
               float* top_score, void* ev_search_start, void* ev_search_stop, alive_stream_t stream, int phase) {
  // the notification request belongs to THIS call, whether it gets as far as the launch or fails on the way
  const NotifyArm notify = g_notify;
  g_notify = NotifyArm{};
  // phase 0: the whole chain; 1 (mode | ALIVE_KNN_MODE_DEFER_FALLBACK): pack, search, finish [, notify] only - the
  // fallback chain for uncertified queries is left to a later alive_knn_match_fallback call on the same workspace;
  // 2: that call (queries: the ones the front half packed into the workspace, or `pq`)
  if (phase == 0 && (mode & ALIVE_KNN_MODE_DEFER_FALLBACK)) phase = 1;
  mode &= ~ALIVE_KNN_MODE_DEFER_FALLBACK;
  const bool run_front = phase != 2, run_back = phase != 1;
  ALIVE_REQUIRE((source || pq || phase == 2) && lib && workspace && top_idx && top_score, "alive_knn_match: NULL argument");
  ALIVE_REQUIRE(batch >= 1 && t >= 1, "alive_knn_match: empty query batch");
  ALIVE_REQUIRE(static_cast<int64_t>(batch) * t < (1ll << 31), "alive_knn_match: too many query frames");
  ALIVE_REQUIRE(k >= 1 && k <= lib->n, "selected index k out of range");
  ALIVE_REQUIRE(k <= ALIVE_KNN_MAX_K, "alive_knn_match: k must be <= %d", ALIVE_KNN_MAX_K);
  const int32_t rows = batch * t;
  const int32_t d = lib->d;
  const int32_t items = lib->items < 1 ? 1 : lib->items;
  // checked before anything is enqueued (the exact scan at the end of the chain has the same limit)
  ALIVE_REQUIRE(d >= 4 && d % 4 == 0 && d <= 1536, "alive_knn_match: d must be a multiple of 4, <= 1536 (got %d)", d);
  ALIVE_REQUIRE(items == 1 || items == batch,
                "alive_knn_match: a library of %d items needs a query batch of the same size (got %d)", items, batch);
  ALIVE_REQUIRE(items == 1 || lib->row_base == 0, "alive_knn_match: batched items cannot be row-sharded");
  mode = resolve_mode(mode, lib->n, d, k);
  ALIVE_REQUIRE(mode == 1 || mode == 2, "alive_knn_match: mode must be 0 (auto), 1 (screen) or 2 (exact)");
  ALIVE_REQUIRE(mode == 2 || k <= ALIVE_KNN_LIST_LEN, "alive_knn_match: the screened path needs k <= %d", ALIVE_KNN_LIST_LEN);
  alive_knn_plan_t plan;
  int64_t off[kOffSlots];
  CollectLayout cl;
  int rc = layout(rows, lib->n, d, k, r_max, num_sms, variant, mode, items, &plan, off, &cl);
  if (rc) return rc;
  ALIVE_REQUIRE(lib->format == ALIVE_KNN_FORMAT_BF16 || lib->format == ALIVE_KNN_FORMAT_FP16, "alive_knn_match: unknown plane format %d",
                lib->format);
  plan.format = lib->format;
  cl.plan.format = lib->format;
  ALIVE_REQUIRE(static_cast<size_t>(off[kOffTotal]) <= workspace_bytes,
                "alive_knn_match: workspace too small (%zu < %lld)", workspace_bytes, static_cast<long long>(off[kOffTotal]));
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "alive_knn_match: workspace must be 256-byte aligned");
  char* ws = static_cast<char*>(workspace);
  const float* q_raw = pq ? pq->raw : reinterpret_cast<float*>(ws + off[kOffQRaw]);
  const float* q_norm = pq ? pq->norm : reinterpret_cast<float*>(ws + off[kOffQNorm]);
  const uint16_t* q_packed = pq ? pq->packed : reinterpret_cast<uint16_t*>(ws + off[kOffQPacked]);
  const float* q_err = pq ? pq->err : reinterpret_cast<float*>(ws + off[kOffQErr]);
  const uint16_t* q_lo = pq ? pq->lo : reinterpret_cast<uint16_t*>(ws + off[kOffQLo]);
  const float* q_err2 = pq ? pq->err2 : reinterpret_cast<float*>(ws + off[kOffQErr2]);
  float* cand_score = reinterpret_cast<float*>(ws + off[kOffCandScore]);
  int32_t* cand_idx = reinterpret_cast<int32_t*>(ws + off[kOffCandIdx]);
  int32_t* sel_n = reinterpret_cast<int32_t*>(ws + off[kOffSelN]);
  int32_t* fb_list = reinterpret_cast<int32_t*>(ws + off[kOffFbList]);
  int32_t* fb_count = reinterpret_cast<int32_t*>(ws + off[kOffFbCount]);
  void* exact_ws = ws + off[kOffExact];

  if (!run_front) {
    // (the front half of an earlier call left the packed queries, the lists and the counters in the workspace)
  } else if (pq == nullptr) {
    // all batch items in ONE pack launch, which also zeroes the fallback counters (no separate memset node)
    rc = pack_impl(source, rows, d, stride_t, stride_d, reinterpret_cast<float*>(ws + off[kOffQRaw]),
                   reinterpret_cast<float*>(ws + off[kOffQNorm]), reinterpret_cast<uint16_t*>(ws + off[kOffQPacked]),
                   reinterpret_cast<float*>(ws + off[kOffQErr]), nullptr, fb_count, 2 * items, stream, t, stride_b,
                   reinterpret_cast<uint16_t*>(ws + off[kOffQLo]), reinterpret_cast<float*>(ws + off[kOffQErr2]), lib->format);
    if (rc) return rc;
  } else {
    // the producer packed the queries (K1 ran as ITS epilogue): only the fallback counters are left to reset
    ALIVE_CHECK_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(int32_t) * 2 * items, as_stream(stream)));
  }
  if (mode == 1) {
    ALIVE_REQUIRE(out == nullptr || lib->row_base == 0, "alive_knn_match: gather needs an unsharded library (row_base == 0)");
    char* ca = ws + off[kOffCollect];
    uint16_t* qc = cl.on ? reinterpret_cast<uint16_t*>(ca + cl.qc) : nullptr;
    float* c_cut = cl.on ? reinterpret_cast<float*>(ca + cl.cut) : nullptr;
    int32_t* c_cnt = cl.on ? reinterpret_cast<int32_t*>(ca + cl.cnt) : nullptr;
    if (run_front) {
      if (ev_search_start) ALIVE_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_search_start), as_stream(stream)));
      // the search may start behind the (still running) query pack: see search_impl
      static const bool pdl = !(getenv("ALIVE_KNN_PDL") && atoi(getenv("ALIVE_KNN_PDL")) == 0);
      rc = search_impl(q_packed, lib->packed, &plan, cand_score, cand_idx, (pdl && !ev_search_start && pq == nullptr) ? 1 : 0, stream);
      if (rc) return rc;
      if (ev_search_stop) ALIVE_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_search_stop), as_stream(stream)));
      rc = finish_impl(cand_score, cand_idx, rows, plan.lists, k, q_raw, q_norm, q_err, lib->raw, lib->norms,
                       lib->stats, lib->n * items, d, r_max, lib->row_base, alpha, out, top_score, top_idx, sel_n, fb_list,
                       fb_count, items, 0, q_packed, qc, c_cut, c_cnt, cl.on ? cl.rows_c : 0, stream);
      if (rc) return rc;
      rc = launch_notify(notify, fb_count, items, stream);
      if (rc) return rc;
    }
    if (!run_back) return 0;
    const int32_t* x_list = fb_list;
    const int32_t* x_count = fb_count;
    if (cl.on) {
      // uncertified queries: a second tensor-core pass collects every frame that reaches the query's cut, those are
      // rescored exactly; only buffer overflows go on to the exhaustive scan.  With the second bf16 plane of library
      // and queries the pass is REFINED (hi.hi + hi.lo + lo.hi, error ~4e-5) against a cut derived from a handful of
      // exact rescorings (refine_prep_kernel): tens of candidates per query on clustered libraries, not thousands.
      static const bool refine_on = !(getenv("ALIVE_KNN_REFINE") && atoi(getenv("ALIVE_KNN_REFINE")) == 0);
      const bool refine = refine_on && lib->lo != nullptr && q_lo != nullptr && q_err2 != nullptr;
      int32_t* c_idx = reinterpret_cast<int32_t*>(ca + cl.idx);
      int32_t* fb2_list = reinterpret_cast<int32_t*>(ca + cl.fb2_list);
      int32_t* fb2_count = fb_count + items;
      uint16_t* qc_lo = reinterpret_cast<uint16_t*>(ca + cl.qc_lo);
      if (refine) {
        rc = refine_prep_impl(cand_score, cand_idx, plan.lists, k, fb_list, fb_count, rows / items, items, cl.rows_c, q_raw,
                              q_norm, q_err, q_err2, q_lo, qc_lo, lib->raw, lib->norms, lib->stats, d, c_cut, stream);
        if (rc) return rc;
      }
      rc = collect_impl(qc, lib->packed, &cl.plan, fb_count, c_cut, c_cnt, c_idx, kCollectCap, refine ? qc_lo : nullptr,
                        refine ? lib->lo : nullptr, stream);
      if (rc) return rc;
      rc = collect_rescore_impl(fb_list, fb_count, rows / items, items, cl.rows_c, c_cnt, c_idx, kCollectCap, k, q_raw, q_norm,
                                lib->raw, lib->norms, lib->n * items, d, alpha, out, top_score, top_idx, lib->row_base,
                                fb2_list, fb2_count, stream);
      if (rc) return rc;
      x_list = fb2_list;
      x_count = fb2_count;
    }
    rc = alive_knn_exact(q_raw, q_norm, rows, lib->raw, lib->norms, lib->n, d, k, x_list, x_count, lib->row_base,
                         exact_ws, top_score, top_idx, alpha, out, items, stream);
    if (rc) return rc;
  } else if (run_front) {       // exact mode has no fallback half: the front half is everything
    ALIVE_REQUIRE(out == nullptr || lib->row_base == 0, "alive_knn_match: gather needs an unsharded library (row_base == 0)");
    rc = alive_knn_exact(q_raw, q_norm, rows, lib->raw, lib->norms, lib->n, d, k, nullptr, nullptr, lib->row_base,
                         exact_ws, top_score, top_idx, alpha, out, items, stream);
    if (rc) return rc;
    rc = launch_notify(notify, nullptr, items, stream);  // exact mode: everything is final here
    if (rc) return rc;
  }
  return 0;
}
}  // namespace
}  // namespace alive

// ---------------------------------------------------------------------------------------------
// CUDA IPC helpers for the peer-memory gather (one process per GPU on one NVLink box): export
// the allocation that contains a device pointer, open another process's export.  The mapping is
// made in the CURRENT device's context with lazy peer access - no context is created on the
// peer GPU in this process.
// ---------------------------------------------------------------------------------------------
extern "C" int alive_knn_ipc_export(const void* dev_ptr, uint8_t* handle64_host, int64_t* offset_host) {
  using namespace alive;
  ALIVE_REQUIRE(dev_ptr && handle64_host && offset_host, "alive_knn_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
  static GetRangeFn get_range = nullptr;
  if (!get_range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    ALIVE_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q));
    ALIVE_REQUIRE(q == cudaDriverEntryPointSuccess && fp, "cuMemGetAddressRange entry point not available");
    get_range = reinterpret_cast<GetRangeFn>(fp);
  }
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = get_range(&base, &size, reinterpret_cast<unsigned long long>(dev_ptr));
  ALIVE_REQUIRE(rc == 0, "cuMemGetAddressRange failed with CUresult %d", rc);
  cudaIpcMemHandle_t h;
  ALIVE_CHECK_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle64_host, &h, 64);
  *offset_host = static_cast<int64_t>(reinterpret_cast<unsigned long long>(dev_ptr) - base);
  return 0;
}

extern "C" int alive_knn_ipc_open(const uint8_t* handle64_host, void** base_out_host) {
  using namespace alive;
  ALIVE_REQUIRE(handle64_host && base_out_host, "alive_knn_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  ALIVE_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out_host = p;
  return 0;
}

extern "C" int alive_knn_ipc_close(void* base) {
  using namespace alive;
  if (base) ALIVE_CHECK_CUDA(cudaIpcCloseMemHandle(base));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Chunk-loop runtime (realtime_inference.py:130-191): the per-chunk host work of a captured pipeline is one graph
// launch and one wait.  Both as single C calls: the interpreter overhead of three framework calls per chunk is a
// tenth of a 100 us chunk.
// ---------------------------------------------------------------------------------------------
extern "C" int alive_knn_graph_launch(void* graph_exec, alive_stream_t stream, void* event) {
  using namespace alive;
  ALIVE_REQUIRE(graph_exec != nullptr, "alive_knn_graph_launch: NULL graph");
  ALIVE_CHECK_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
  if (event) ALIVE_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), as_stream(stream)));
  return 0;
}

extern "C" int alive_knn_arm_notify(void* host_flag, void* dev_counter) {
  using namespace alive;
  ALIVE_REQUIRE((host_flag == nullptr) == (dev_counter == nullptr), "alive_knn_arm_notify: both pointers or none");
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(host_flag) & 7) == 0 && (reinterpret_cast<uintptr_t>(dev_counter) & 7) == 0,
                "alive_knn_arm_notify: 8-byte aligned words");
  g_notify.host_flag = static_cast<unsigned long long*>(host_flag);
  g_notify.dev_counter = static_cast<unsigned long long*>(dev_counter);
  return 0;
}

extern "C" int alive_knn_flag_wait(const void* host_flag, uint64_t last, uint64_t* value_out, int64_t timeout_us) {
  using namespace alive;
  ALIVE_REQUIRE(host_flag != nullptr && value_out != nullptr, "alive_knn_flag_wait: NULL argument");
  const volatile uint64_t* f = static_cast<const volatile uint64_t*>(host_flag);
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spin = 0;; ++spin) {
    const uint64_t v = *f;
    if (v != last) {
      std::atomic_thread_fence(std::memory_order_acquire);
      *value_out = v;
      return 0;
    }
    if ((spin & 0x3ff) == 0x3ff && timeout_us > 0 &&
        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_us) {
      set_error("alive_knn_flag_wait: no notification within %lld us", static_cast<long long>(timeout_us));
      return -3;
    }
  }
}

extern "C" int alive_knn_event_wait(void* event) {
  using namespace alive;
  ALIVE_REQUIRE(event != nullptr, "alive_knn_event_wait: NULL event");
  // poll: the wake-up of a blocking wait costs more than the few microseconds this spins
  for (;;) {
    const cudaError_t e = cudaEventQuery(static_cast<cudaEvent_t>(event));
    if (e == cudaSuccess) return 0;
    if (e != cudaErrorNotReady) {
      set_error("cudaEventQuery failed: %s", cudaGetErrorString(e));
      return -2;
    }
  }
}

