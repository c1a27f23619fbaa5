/*
 * alive_knn.h - C ABI of the B200-native kNN voice-library matching path.
 *
 * This is the drop-in boundary for ONE hot path of uthree/ALiVE-VC:
 *   module/common.py:96-109         match_features(source, reference, k, alpha)
 *   module/voice_library.py:15-33   VoiceLibrary.match(source, k, alpha)
 * The reference has no FFI/plugin layer (it is pure Python calling torch ops),
 * so each entry point below cites the reference LINES whose torch ops it
 * replaces.  The Python host side (alive_vc_b200/) binds these with ctypes and
 * keeps the reference's Python signatures; INTEGRATION.md shows the stub a
 * maintainer adds to module/common.py / module/voice_library.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises
 *     the host unless stated;
 *   - return value: 0 on success, negative on error; the message is available
 *     from alive_knn_last_error() (thread-local);
 *   - "rows" are frames: a library of N frames is N rows of D floats.
 *   - there is NO CPU fallback anywhere behind this ABI.
 *
 * Built for sm_100a only (tcgen05 / TMEM / TMA).
 */
#ifndef ALIVE_KNN_H_
#define ALIVE_KNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* alive_stream_t; /* cudaStream_t */

#define ALIVE_KNN_ABI_VERSION 7
#define ALIVE_KNN_LIST_LEN 8      /* entries kept per running top list in the fused kernel */
#define ALIVE_KNN_TILE_M 128      /* query frames per tensor-core tile   */
#define ALIVE_KNN_TILE_N 256      /* library frames per tensor-core tile */
#define ALIVE_KNN_MAX_K 64        /* largest k supported by the exact selectors */
/* 16-bit format of the packed planes (tensor-core operands).  bf16 is what the project brief names; IEEE fp16 runs
 * at the same tensor-core rate (tcgen05.mma kind::f16 takes either) and, normalised frames living in [-1, 1], rounds
 * 8x finer: the screening error bound - and with it the band the certificate has to clear - shrinks 8x.  Queries
 * must be packed in the format of the library they are matched against. */
#define ALIVE_KNN_FORMAT_BF16 0
#define ALIVE_KNN_FORMAT_FP16 1

/* Work decomposition of one alive_knn_search launch (filled by alive_knn_plan). */
typedef struct alive_knn_plan {
  int32_t t;                 /* query frames */
  int64_t n;                 /* library frames */
  int32_t d;                 /* feature dim (multiple of 64) */
  int32_t ctas_per_unit;     /* 1: cta_group::1 (128-query units), 2: CTA pair (256-query units) */
  int32_t m_units;           /* ceil(t / (128*ctas_per_unit)) */
  int32_t n_tiles;           /* ceil(n / 256) */
  int32_t segments;          /* library is cut into this many runs of whole tiles */
  int32_t tiles_per_segment;
  int32_t lists;             /* running lists per query = 2 * segments */
  int32_t grid;              /* CTAs launched (multiple of ctas_per_unit) */
  int32_t items;             /* independent (query batch, library) pairs laid out back to back; 1 = plain */
  int32_t format;            /* ALIVE_KNN_FORMAT_* of BOTH packed operands (set by the caller; the planners write BF16) */
  int32_t kernel;            /* 0: tiled kernel (fields as described above)
                              * 1: "skinny" kernel for t <= 32 (one realtime chunk): 128-frame library tiles on
                              *    the M side of the MMA, the query chunk resident in shared memory; n_tiles =
                              *    ceil(n / 128), CTA c takes tiles c, c + grid, ... (segments = grid,
                              *    tiles_per_segment = ceil(n_tiles / grid)) and leaves ONE list per query:
                              *    lists = grid */
} alive_knn_plan_t;

/* Library statistics produced by alive_knn_pack (4 x uint32 on the device, zeroed by the caller):
 *   [0] bit pattern of max over rows of || bf16(x/|x|) - x/|x| ||_2   (float >= 0)
 *   [1] number of rows whose normalised form is not finite (zero or inf/nan rows)
 *   [2] bit pattern of max over rows of || x/|x| - hi - lo ||_2, the residual of the TWO-plane split
 *       (hi = bf16(x/|x|), lo = bf16(x/|x| - hi)); [3] reserved */
#define ALIVE_KNN_STATS_WORDS 4

const char* alive_knn_last_error(void);
int alive_knn_abi_version(void);

/* K1 - normalise and pack.  Replaces, ONCE per library instead of once per
 * call, common.py:101,103 and the `reference / reference_norm` of :104
 * (voice_library.py:25,27,28): reads frames x[i*stride_n + j*stride_d]
 * (i < n, j < d; the reference's [1,D,N] layout is stride_n=1, stride_d=N),
 * writes
 *   raw    [n,d] float32 row-major copy of the UN-normalised frames (gather + rescoring)
 *   norms  [n]   float32 L2 norms (fp64 accumulation, rounded once)
 *   packed [n,d] 16-bit row-major (bf16 or fp16, `format`), frames divided by their norm (tensor-core operand,
 *                TMA friendly)
 *   err    [n]   float32 || bf16(x/|x|) - x/|x| ||_2 per row (may be NULL)
 *   stats  [4]   see above; must be zeroed by the caller before the FIRST pack
 *                of a library (several packs may accumulate into one stats).
 *   lo     [n,d] bf16 row-major, the SECOND plane bf16(x/|x| - packed) (may be NULL: 2 B per element that buy
 *                the refined collect pass - hi.hi + hi.lo + lo.hi on the tensor cores, error ~4e-5 instead of
 *                ~4e-3 - which keeps clustered libraries, e.g. thousands of near-identical silence frames, off the
 *                exhaustive scan);  err2 [n] float32 || x/|x| - packed - lo ||_2 per row (may be NULL).
 * Also used per call for the query frames (common.py:100,102,104 lhs).
 * Any strides are accepted; the two layouts that matter have kernels of their own: the
 * reference's channel-major [D,N] (stride_n == 1) and a producer's row-major [N,D]
 * (stride_d == 1), both with 16-byte aligned rows.  Every variant writes the same bits. */
int alive_knn_pack(const float* x, int64_t n, int32_t d, int64_t stride_n, int64_t stride_d,
                   float* raw, float* norms, uint16_t* packed, float* err, uint32_t* stats,
                   uint16_t* lo, float* err2, int32_t format, alive_stream_t stream);

/* Fill `plan` for t queries against n library frames on a device with
 * `num_sms` SMs.  variant: 1 or 2 CTAs per unit, 3 = the skinny kernel (t <= 32, single item);
 * 0 = library default (skinny when it applies, else CTA pairs once t > 128). Host only. */
int alive_knn_plan(int32_t t, int64_t n, int32_t d, int32_t num_sms, int32_t variant,
                   alive_knn_plan_t* plan_host);
/* Batched form (BASELINE cfg5, train_decoder.py:134-135): `items` independent problems of t
 * queries x n frames each, queries stored as [items*t, d] and libraries as [items*n, d]; item i's
 * queries only see item i's frames.  One launch covers all items. */
int alive_knn_plan_batched(int32_t items, int32_t t, int64_t n, int32_t d, int32_t num_sms,
                           int32_t variant, alive_knn_plan_t* plan_host);

/* K2 - fused similarity + running top list.  Replaces the bmm of
 * common.py:104 and the first pass of torch.topk :105 WITHOUT materialising
 * the [T,N] score matrix: TMA-fed tcgen05 bf16 MMAs accumulate 128x256 score
 * tiles in TMEM; epilogue warps keep, per query and per list, the
 * ALIVE_KNN_LIST_LEN best (score, frame) pairs.  plan.kernel == 1 (one realtime chunk,
 * t <= 32; realtime_inference.py:165): the operands are swapped - 128-frame library tiles
 * stream through a deep TMA ring as the M side of the MMA, the query chunk stays resident
 * in shared memory - and every CTA leaves one list per query.
 *   q_packed [items*t,d] bf16, lib_packed [items*n,d] bf16 (from alive_knn_pack)
 *   cand_score [items*t, plan.lists, 8] float32 (descending, -inf padded)
 *   cand_idx   [items*t, plan.lists, 8] int32   (-1 padded; frame index inside [items*n]) */
int alive_knn_search(const uint16_t* q_packed, const uint16_t* lib_packed,
                     const alive_knn_plan_t* plan_host,
                     float* cand_score, int32_t* cand_idx, alive_stream_t stream);

/* K2b - certificate + prune.  From the screened lists pick, per query, every
 * frame that can still belong to the exact top-k given the bf16 screening
 * error bound eps = lib_err + q_err[t] + lib_err*q_err[t] + slack(d) (slack: the tensor core's
 * float32 accumulation over d/16 instructions, select.cu accum_slack); a query whose
 * lists cannot prove completeness (or that needs more than r_max survivors,
 * or whose own norm is not finite, or when the library holds non-finite rows)
 * is appended to fb_list for the exact scan.
 *   sel_idx [t,r_max] int32, sel_n [t] int32, fb_list [t] int32, fb_count [1] int32 (zeroed here) */
int alive_knn_prune(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists,
                    int32_t k, int32_t d, const float* q_err, const float* q_norm, const uint32_t* lib_stats,
                    int32_t r_max, int32_t* sel_idx, int32_t* sel_n, int32_t* fb_list,
                    int32_t* fb_count, alive_stream_t stream);

/* K3 - exact rescoring of the survivors.  Mirrors common.py:102-105 on the
 * candidate set: each frame is normalised first (x / |x|, float32), then the
 * dot product is accumulated (in fp64, rounded once to float32); top-k by
 * (score desc, frame asc).  Queries with sel_n < 0 are skipped (fallback).
 *   top_score [t,k] float32, top_idx [t,k] int64 (= frame + idx_base) */
int alive_knn_rescore(const float* q_raw, const float* q_norm, int32_t t,
                      const float* lib_raw, const float* lib_norm, int32_t d,
                      const int32_t* sel_idx, const int32_t* sel_n, int32_t r_max, int32_t k,
                      int64_t idx_base, float* top_score, int64_t* top_idx, alive_stream_t stream);

/* K2b+K3+K4 fused, one CTA per query: alive_knn_prune + alive_knn_rescore (+ the
 * gather+mean+blend of alive_knn_gather_mean when out != NULL) in one launch.  Uncertified
 * queries are appended to fb_list (sel_n = -1) and left to alive_knn_exact.  t = ALL query
 * frames (items * t_item); with items > 1 the lists are kept per item: fb_list [items][t/items],
 * fb_count [items]. */
int alive_knn_finish(const float* cand_score, const int32_t* cand_idx, int32_t t, int32_t lists, int32_t k,
                     const float* q_raw, const float* q_norm, const float* q_err, const float* lib_raw,
                     const float* lib_norm, const uint32_t* lib_stats, int64_t n, int32_t d, int32_t r_max,
                     int64_t idx_base, float alpha, float* out, float* top_score, int64_t* top_idx,
                     int32_t* sel_n, int32_t* fb_list, int32_t* fb_count, int32_t items,
                     alive_stream_t stream);

/* Exact scan (no screen): common.py:102-105 for the queries listed in
 * q_list[0..*q_count) (both device; NULL/NULL = all t queries) against all n
 * frames, same arithmetic and tie rule as alive_knn_rescore; NaN similarities
 * rank first like torch.topk.  workspace: alive_knn_exact_workspace_bytes().
 * items > 1: t = items * t_item queries, item i scans frames [i*n, (i+1)*n) of lib_raw
 * ([items*n, d]); q_list is [items][t_item], q_count [items]; indices are global (i*n + frame).
 * out (nullable, [t,d] f32): when given, the scanned queries are also gathered
 * (common.py:107-109, same arithmetic as alive_knn_gather_mean with `alpha`). */
size_t alive_knn_exact_workspace_bytes(int32_t t, int64_t n, int32_t k, int32_t items);
int alive_knn_exact(const float* q_raw, const float* q_norm, int32_t t,
                    const float* lib_raw, const float* lib_norm, int64_t n, int32_t d, int32_t k,
                    const int32_t* q_list, const int32_t* q_count, int64_t idx_base,
                    void* workspace, float* top_score, int64_t* top_idx, float alpha, float* out,
                    int32_t items, alive_stream_t stream);

/* Multi-GPU merge: after an all-gather of every rank's exact local top-k,
 * scores/idx are [ranks,t,k]; writes the global top-k (score desc, frame asc). */
int alive_knn_merge(const float* scores, const int64_t* idx, int32_t ranks, int32_t t, int32_t k,
                    float* top_score, int64_t* top_idx, alive_stream_t stream);

/* The same merge on RECORDS - what one all-gather moves per rank: idx [t,k] int64 immediately
 * followed by score [t,k] float32 (t*k*12 bytes), the records of consecutive ranks `record_stride`
 * bytes apart.  Entries with idx < 0 are padding (a shard holding fewer than k frames). */
int alive_knn_merge_records(const void* records, int64_t record_stride, int32_t ranks, int32_t t,
                            int32_t k, float* top_score, int64_t* top_idx, alive_stream_t stream);

/* K4 - gather + mean + blend.  Replaces common.py:107-109
 * (voice_library.py:31-33): out[t,:] = mean_j raw[top_idx[t,j],:] summed
 * sequentially in descending-score order then divided by k, blended as
 * out*(1-alpha) + q_raw*alpha with separately rounded products.
 *   out [t,d] float32 row-major (the reference returns exactly this block
 *   viewed as [D,T]).
 *   q_norm [t] float32 or NULL: the norms alive_knn_pack wrote for the query frames.  When given,
 *   the query row is only read where it can change the result (alpha != 0, a non-finite query -
 *   0*inf is NaN in the reference - or a mean of exactly zero); NULL = always read it.  Same bits. */
int alive_knn_gather_mean(const float* lib_raw, int64_t n, int32_t d, const int64_t* top_idx,
                          int32_t t, int32_t k, const float* q_raw, const float* q_norm, float alpha,
                          float* out, alive_stream_t stream);

/* Sharded variant, step 1: rows[t,k,d] = raw[top_idx - row_lo] where
 * row_lo <= top_idx < row_lo + n, else 0 (so a sum over ranks is exact). */
int alive_knn_gather_rows(const float* lib_raw, int64_t n, int32_t d, int64_t row_lo,
                          const int64_t* top_idx, int32_t t, int32_t k, float* rows,
                          alive_stream_t stream);
/* Sharded variant, step 2: same mean + blend as alive_knn_gather_mean on gathered rows. */
int alive_knn_mean_blend(const float* rows, int32_t t, int32_t k, int32_t d, const float* q_raw,
                         float alpha, float* out, alive_stream_t stream);

/* Sharded variant over PEER memory (NVLink, CUDA IPC): shard_raw[r] (device array of `shards`
 * device pointers, each mapped into this process) holds frames bounds[r] .. bounds[r+1]-1
 * (bounds: device array of shards+1 int64).  Same arithmetic as alive_knn_gather_mean: every
 * rank computes the full, bit-identical [t,d] result with no collective after the merge. */
int alive_knn_gather_mean_peers(const float* const* shard_raw, const int64_t* bounds, int32_t shards,
                                int32_t d, const int64_t* top_idx, int32_t t, int32_t k,
                                const float* q_raw, const float* q_norm, float alpha, float* out,
                                alive_stream_t stream);

/* The final step of the row-sharded match as ONE kernel (common.py:105 across shards + :107-109):
 * for query rows [row0, row0+rows) of the t queries, merge the per-rank records (see
 * alive_knn_merge_records) into the global top-k and gather + mean + blend the k winning raw frames
 * straight from the GPU that owns them (shard_raw / bounds as in alive_knn_gather_mean_peers; a local
 * pointer is fine).  No collective follows the all-gather of the records.
 *   q_raw [t,d] / q_norm [t] (nullable) are indexed by the GLOBAL row; out [rows,d],
 *   top_score [rows,k] / top_idx [rows,k] (both nullable) by the row inside the range.
 *   d must be 128, 256, 512, 768, 1024 or 1536. */
int alive_knn_merge_gather(const void* records, int64_t record_stride, int32_t ranks, int32_t t, int32_t k,
                           int32_t row0, int32_t rows, const float* const* shard_raw,
                           const int64_t* bounds, int32_t shards, int32_t d, const float* q_raw,
                           const float* q_norm, float alpha, float* out, float* top_score,
                           int64_t* top_idx, alive_stream_t stream);

/* CUDA IPC plumbing for the peer-memory gather (host-side; one process per GPU, same box):
 * export the allocation containing dev_ptr (64-byte handle + byte offset of dev_ptr inside it),
 * open / close a handle exported by another process (mapped into the current device's context
 * with lazy peer access; the returned pointer is the allocation BASE). */
int alive_knn_ipc_export(const void* dev_ptr, uint8_t* handle64_host, int64_t* offset_host);
int alive_knn_ipc_open(const uint8_t* handle64_host, void** base_out_host);
int alive_knn_ipc_close(void* base);

/* Backward of VoiceLibrary.match w.r.t. tokens (voice_library.py:31-33 under
 * autograd): grad_rows[top_idx[t,j],:] += scale * grad_out[t,:], scale=(1-alpha)/k.
 * grad_rows [n,d] float32 must be zeroed by the caller. */
int alive_knn_scatter_grad(const float* grad_out, const int64_t* top_idx, int32_t t, int32_t k,
                           int32_t d, float scale, float* grad_rows, int64_t n,
                           alive_stream_t stream);

/* A packed library as produced by alive_knn_pack (all device pointers). */
typedef struct alive_knn_library {
  const uint16_t* packed;   /* [n,d] bf16 */
  const float* raw;         /* [n,d] f32  */
  const float* norms;       /* [n]   f32  */
  const uint32_t* stats;    /* [4]   u32  */
  int64_t n;                /* frames per item */
  int32_t d;
  int64_t row_base;         /* global index of frame 0 (row-sharded libraries), else 0 */
  int32_t items;            /* >= 1: independent libraries of n frames each, stored back to back */
  const uint16_t* lo;       /* [n,d] second plane (alive_knn_pack `lo`), or NULL */
  int32_t format;           /* ALIVE_KNN_FORMAT_* of packed / lo */
} alive_knn_library_t;

/* One-call pipeline = module/common.py:96-109 for `batch` x `t` query frames against one
 * packed library: pack queries -> search -> prune -> rescore -> exact scan of uncertified
 * queries -> gather+mean+blend.  source[b*stride_b + i*stride_t + j*stride_d] (the
 * reference's [B,D,T] layout: stride_b=D*T, stride_t=1, stride_d=T).
 *   lib->items > 1: `batch` must equal lib->items; batch item b is matched against frames
 *   [b*n, (b+1)*n) only (one launch for all items); indices are global (b*n + frame)
 *   mode: 0 auto (screen; exact scan when k > 8 or d % 64 != 0), 1 screen, 2 exact scan
 *   workspace: 256-byte aligned, at least offsets[11] bytes of alive_knn_match_layout
 *   out [batch*t, d] f32 (NULL = skip the gather, e.g. for a sharded library),
 *   top_idx [batch*t, k] int64 (global frame indices), top_score [batch*t, k] f32.
 *   ev_search_start/stop: optional cudaEvent_t recorded around the alive_knn_search launch
 *   (used by bench.py to time the dominant kernel inside the timed region), else NULL.
 * Everything is enqueued on `stream`; no host synchronisation (CUDA-graph capturable).
 * Uncertified queries (device-side lists, no host sync): a second tensor-core pass collects every frame that
 * reaches the query's cut - refined with the second bf16 planes when lib->lo is given - and rescores those
 * exactly; what overflows goes to the exhaustive scan.  Works per item for batched libraries.
 * alive_knn_match_layout fills 14 byte offsets into the workspace:
 *   0 q_raw 1 q_norm 2 q_packed 3 q_err 4 cand_score 5 cand_idx 6 collect-pass area 7 sel_n
 *   8 fb_list 9 fb_count (2*items words: uncertified after the screen | after the collect pass)
 *   10 exact scratch 11 TOTAL bytes 12 q_lo 13 q_err2. */
int alive_knn_match_layout(int32_t rows, int64_t n, int32_t d, int32_t k, int32_t r_max, int32_t mode,
                           int32_t num_sms, int32_t variant, int32_t items, int64_t* offsets14_host);
int alive_knn_match(const float* source, int32_t batch, int32_t t, int64_t stride_b, int64_t stride_t,
                    int64_t stride_d, const alive_knn_library_t* lib_host, int32_t k, float alpha,
                    int32_t r_max, int32_t mode, int32_t num_sms, int32_t variant, void* workspace,
                    size_t workspace_bytes, float* out, int64_t* top_idx, float* top_score,
                    void* ev_search_start, void* ev_search_stop, alive_stream_t stream);

/* The same pipeline for queries that arrive ALREADY packed (a producer - the content encoder,
 * module/content_encoder.py:8-25 - that runs alive_knn_pack as its own epilogue, e.g. once per utterance that is
 * then matched against several speakers' libraries, or on row-major [T, D] frames it kept channels-last):
 * q_raw [batch*t, d] f32, q_norm [batch*t], q_packed [batch*t, d] bf16, q_err [batch*t] exactly as alive_knn_pack
 * wrote them.  No K1 launch; everything else (search, certificate, rescoring, fallbacks, gather) as alive_knn_match. */
int alive_knn_match_packed(const float* q_raw, const float* q_norm, const uint16_t* q_packed, const float* q_err,
                           const uint16_t* q_lo, const float* q_err2,   /* second plane of the queries, or NULL / NULL */
                           int32_t batch, int32_t t, const alive_knn_library_t* lib_host, int32_t k, float alpha,
                           int32_t r_max, int32_t mode, int32_t num_sms, int32_t variant, void* workspace,
                           size_t workspace_bytes, float* out, int64_t* top_idx, float* top_score,
                           alive_stream_t stream);

/* Deferred fallback (the realtime loop again): with ALIVE_KNN_MODE_DEFER_FALLBACK or-ed into `mode`,
 * alive_knn_match / alive_knn_match_packed enqueue only pack -> search -> finish (-> the notify kernel, if armed):
 * the six launches of the fallback chain, idle whenever every query certifies, are NOT enqueued.  A caller that learns
 * (alive_knn_arm_notify's flag, bit 0) that some query was left uncertified enqueues them afterwards with
 * alive_knn_match_fallback - same arguments and SAME workspace as the front half; q_* = NULL when the front half was
 * alive_knn_match (its packed queries are in the workspace), else the packed queries again.  Results are identical to
 * the one-call form.  Without a notification there is no way to know: do not defer. */
#define ALIVE_KNN_MODE_DEFER_FALLBACK 0x100
int alive_knn_match_fallback(const float* q_raw, const float* q_norm, const uint16_t* q_packed, const float* q_err,
                             const uint16_t* q_lo, const float* q_err2, int32_t batch, int32_t t,
                             const alive_knn_library_t* lib_host, int32_t k, float alpha, int32_t r_max, int32_t mode,
                             int32_t num_sms, int32_t variant, void* workspace, size_t workspace_bytes, float* out,
                             int64_t* top_idx, float* top_score, alive_stream_t stream);

/* Chunk-loop runtime (realtime_inference.py:130-191; host-side, no kernels): launch an instantiated CUDA graph
 * (cudaGraphExec_t) on `stream` and record `event` (cudaEvent_t, nullable) behind it - one call per chunk; wait for
 * an event by polling (a blocking wait's wake-up costs more than a ~100 us chunk can spare). */
int alive_knn_graph_launch(void* graph_exec, alive_stream_t stream, void* event);
int alive_knn_event_wait(void* event);

/* Early result notification for the same loop.  alive_knn_arm_notify(host_flag, dev_counter): the NEXT
 * alive_knn_match / alive_knn_match_packed call of this thread (one shot; typically made once, under graph capture)
 * launches a one-thread kernel right behind its finish kernel that increments *dev_counter (8-byte device word) and
 * stores (counter << 1) | any_uncertified to *host_flag (8-byte word in MAPPED pinned host memory), after a
 * system-scope fence.  When the match writes its result rows into mapped host memory too, a host that sees the flag
 * change with bit 0 clear has the complete result of the chunk and need not wait for the (idle) fallback launches
 * still queued behind the finish kernel; bit 0 set = some query went to the fallback chain: wait for the stream.
 * alive_knn_flag_wait spins on the host word until it differs from `last` (returns the new value; -3 on timeout,
 * timeout_us <= 0 = no timeout).  NULL, NULL disarms. */
int alive_knn_arm_notify(void* host_flag, void* dev_counter);
int alive_knn_flag_wait(const void* host_flag, uint64_t last, uint64_t* value_out, int64_t timeout_us);

#ifdef __cplusplus
}
#endif
#endif /* ALIVE_KNN_H_ */
