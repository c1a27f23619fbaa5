// K2 - fused cosine-similarity contraction + running top list (sm_100a).
//
// Replaces the `torch.bmm` of module/common.py:104 (voice_library.py:28) and the
// scan of `torch.topk` (:105 / :29) of the reference, without ever writing the
// [T,N] score matrix to HBM.
//
// Shape of the computation
//   scores[t, n] = sum_d Q[t,d] * L[n,d]     Q [T,768] bf16, L [N,768] bf16, both
//                                            row-major = "K-major" for the MMA.
//   One work unit = (group of 128*kCtas queries) x (segment of whole 256-frame
//   tiles of the library).  A persistent CTA (or CTA pair) walks its units; per
//   tile it issues 48 tcgen05.mma (M=128*kCtas, N=256, K=16) into one of two
//   256-column TMEM accumulators while the 8 epilogue warps drain the other one.
//
// Warp roles (384 threads):
//   warp 0      TMA producer   (one lane): cp.async.bulk.tensor, 128B swizzle, mbarrier tx
//   warp 1      MMA issuer     (one lane, leader CTA only): tcgen05.mma + tcgen05.commit
//   warp 2      TMEM allocator (512 columns) / deallocator
//   warp 3      idle
//   warps 4-11  epilogue: warp w owns TMEM lanes 32*(w%4).. (= 32 queries) and
//               columns 128*((w-4)/4).. of every tile; each THREAD owns one query
//               row and keeps a sorted 8-entry (score, frame) list in registers.
//               tcgen05.ld 32x32b.x32 -> chunk max -> rare ordered insert.
//
// Output: per (query, list) the 8 best screened (score, frame) pairs; list id =
// 2*segment + column half.  The list minimum is an upper bound on every score
// the list dropped, which is what the certificate in select.cu needs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace alive {
namespace {

constexpr int kBlockM = ALIVE_KNN_TILE_M;   // 128 queries per CTA (TMEM lanes)
constexpr int kBlockN = ALIVE_KNN_TILE_N;   // 256 library frames per tile (TMEM columns)
constexpr int kBlockK = 64;                 // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kListLen = ALIVE_KNN_LIST_LEN;
constexpr int kNumEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = 32 * (kFirstEpiWarp + kNumEpiWarps);   // 384
constexpr int kTmemCols = 512;                                  // 2 accumulators x 256 columns
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;             // 16 KB per stage per CTA

// kRefine (collect mode only): every stage carries BOTH bf16 planes of both operands (hi = bf16(x^), lo = bf16(x^ - hi))
// and the MMA warp accumulates hi.hi + hi.lo + lo.hi - the split-bf16 refinement of the screen (error ~1e-5 instead
// of ~4e-3).  Twice the bytes per stage, so fewer stages; the collecting epilogue needs no insert scratch.
template <int kCtas, bool kRefine = false> struct Cfg {
  static constexpr int kPlanes = kRefine ? 2 : 1;
  static constexpr int kStages = kRefine ? ((kCtas == 1) ? 2 : 3) : ((kCtas == 1) ? 4 : 6);
  static constexpr int kBRows = kBlockN / kCtas;                       // library rows this CTA loads per tile
  static constexpr uint32_t kBBytes = kBRows * kBlockK * 2;            // 32 KB or 16 KB (one plane)
  static constexpr uint32_t kAStage = kABytes * kPlanes;               // bytes of A per stage (hi [+ lo])
  static constexpr uint32_t kBStage = kBBytes * kPlanes;
  static constexpr uint32_t kStageBytes = kAStage + kBStage;
  static constexpr uint32_t kTxBytes = kStageBytes * kCtas;            // bytes landing per stage, whole unit
  static constexpr uint32_t kSmemData = kStages * kStageBytes;
  static constexpr uint32_t kScratchBytes = kRefine ? 0 : 32 * kNumEpiWarps * 32 * 4;  // epilogue chunk scratch, 32 KB
  static constexpr uint32_t kSmemBytes = kSmemData + kScratchBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct SearchParams {
  int items;           // independent problems laid out back to back (1 = plain)
  int t;               // query frames per item
  int n;               // library frames per item
  int k_blocks;        // d / 64
  int m_units;
  int segments;
  int tiles_per_segment;
  int n_tiles;
  int lists;
  float* cand_score;
  int* cand_idx;
  int half;            // operands are fp16 (else bf16)
  int debug;           // diagnostics only: 1 = epilogue skips the scan, 2 = also skips the TMEM loads
  unsigned long long hint_q;     // L2 eviction policy of the query-tile loads
  unsigned long long hint_lib;   // L2 eviction policy of the library-tile loads
  unsigned int* sync_ctr;        // grid-wide pacing counter (zeroed before the launch), or NULL
  int sync_every;                // producers re-align every this many tiles ...
  int sync_rounds;               // ... for this many epochs (every CTA reaches them)
  // collect mode (knn_search_kernel<kCtas, true>): instead of running top lists, EVERY frame whose
  // screened score is not below the row's cut is appended to the row's candidate buffer
  const int* c_active;           // device [items]: live query rows of every item (rows beyond it are skipped)
  const float* c_cut;            // [items * t] per-row cut (the refined lower bound, or S_k - 2 eps of the first pass)
  int* c_cnt;                    // [items * t] candidates found (may exceed c_cap: overflow)
  int* c_idx;                    // [items * t, c_cap] frame indices
  int c_cap;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of `local_addr` (a shared::cta address) inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier given by a shared::cluster address (own or peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  // default (.release.cta) semantics as in CUTLASS' ClusterBarrier::arrive(cta_id): what is being
  // ordered are TMEM reads (tcgen05.wait::ld + tcgen05.fence::before_thread_sync), not global
  // memory - an explicit .release.cluster here costs a MEMBAR.ALL.GPU per tile per warp
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {   // ~10-15 s: far beyond any legitimate wait, short enough to fail fast
      printf("alive_knn_search: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
             blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// one lane of the (converged) warp; lets ptxas keep descriptors/addresses in uniform registers and
// issue UTMALDG / UTCHMMA / UTCBAR directly instead of wrapping each one in an ELECT+BRA.U.ANY loop
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
// L2 eviction-priority policies accepted by the .L2::cache_hint operand of cp.async.bulk.tensor
// (the encodings CUTLASS uses for TMA::CacheHintSm90)
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;

template <int kCtas>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1,
                                            uint64_t policy) {
  if constexpr (kCtas == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
  } else {
    // data lands in THIS CTA's smem, the transaction bytes are reported to the barrier at
    // `bar` (a shared::cluster address inside the leader CTA of the pair)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
  }
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int kCtas> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  if constexpr (kCtas == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
template <int kCtas> __device__ __forceinline__ void tmem_relinquish() {
  if constexpr (kCtas == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCtas> __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if constexpr (kCtas == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as 128-byte rows with the
// 128B swizzle TMA applies (8-row groups of 1024 B): start>>4 | LBO(16B, unused)=1 |
// SBO = 1024 B | version 1 (Blackwell) | layout SWIZZLE_128B (=2 in bits 61..63).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: D=f32 (bit 4), A and B formats at bits 7 and 10 (0 = fp16, 1 = bf16), both
// K-major, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool half) {
  return (1u << 4) | ((half ? 0u : 1u) << 7) | ((half ? 0u : 1u) << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
template <int kCtas>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCtas == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive once every previously issued tcgen05.mma of this thread has completed
template <int kCtas> __device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (kCtas == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    // same barrier offset in both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(static_cast<uint16_t>(3))
        : "memory");
  }
}

#define ALIVE_R32(v)                                                                               \
  v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14],    \
      v[15], v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23], v[24], v[25], v[26], v[27],    \
      v[28], v[29], v[30], v[31]

// 32 lanes x 32 consecutive fp32 columns: thread i receives TMEM lane (base lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld that also "touches" the destination registers so the compiler cannot
// schedule their first use above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// Ordered insert of (v, idx) into the descending list; requires v > s[kListLen-1].
// Equal scores keep the earlier (lower) frame first.
__device__ __forceinline__ void list_insert(float (&s)[kListLen], uint32_t (&id)[kListLen], float v, uint32_t idx) {
#pragma unroll
  for (int i = kListLen - 1; i >= 1; --i) {
    const bool above = v > s[i - 1];   // v also displaces entry i-1 -> entry i-1 shifts down into i
    const bool here = v > s[i];
    s[i] = above ? s[i - 1] : (here ? v : s[i]);
    id[i] = above ? id[i - 1] : (here ? idx : id[i]);
  }
  const bool top = v > s[0];
  s[0] = top ? v : s[0];
  id[0] = top ? idx : id[0];
}

// One 32-column chunk of one query row.  Fast path (the common case once the list has warmed
// up): one 3-input max tree, nothing beats the list minimum, done.  Slow path: a bit mask of
// the columns that beat it, the chunk parked in this thread's shared-memory scratch column,
// and a loop over the set bits with ONE copy of the ordered insert - the code stays small
// enough to live in the instruction cache (the fully unrolled variant was fetch-bound).
template <int kPark = 32>
__device__ __forceinline__ void scan_chunk(const uint32_t (&raw)[32], int col_base, int n_valid,
                                           float (&s)[kListLen], uint32_t (&id)[kListLen],
                                           float* __restrict__ scratch /* [kPark][256], this thread's column */) {
  static_assert(kPark == 32 || kPark == 16, "park 32 or 16 columns at a time");
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  if (col_base + 32 > n_valid) {   // ragged last tile of the library (warp-uniform branch)
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col_base + j >= n_valid) v[j] = -INFINITY;
  }
  float m = v[0];
#pragma unroll
  for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
  if (m > s[kListLen - 1]) {
    const float thr = s[kListLen - 1];
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) mask |= (v[j] > thr) ? (1u << j) : 0u;
    if constexpr (kPark == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) scratch[j * (kNumEpiWarps * 32)] = v[j];
      while (mask) {
        const int j = __ffs(static_cast<int>(mask)) - 1;
        mask &= mask - 1;
        const float x = scratch[j * (kNumEpiWarps * 32)];
        if (x > s[kListLen - 1]) list_insert(s, id, x, static_cast<uint32_t>(col_base + j));
      }
    } else {
      // half the scratch: the two 16-column halves are parked one after the other, ONE copy of the insert loop
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t mh = (mask >> (16 * h)) & 0xFFFFu;
        if (mh == 0) continue;
        if (h == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) scratch[j * (kNumEpiWarps * 32)] = v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) scratch[j * (kNumEpiWarps * 32)] = v[16 + j];
        }
        while (mh) {
          const int j = __ffs(static_cast<int>(mh)) - 1;
          mh &= mh - 1;
          const float x = scratch[j * (kNumEpiWarps * 32)];
          if (x > s[kListLen - 1]) list_insert(s, id, x, static_cast<uint32_t>(col_base + 16 * h + j));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int kCtas, bool kCollect, bool kRefine>
__global__ void __launch_bounds__(kThreads, 1)
knn_search_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_lib,
                  const __grid_constant__ CUtensorMap tmap_q_lo, const __grid_constant__ CUtensorMap tmap_lib_lo,
                  const SearchParams p) {
  static_assert(!kRefine || kCollect, "the refined accumulation exists in collect mode only");
  using C = Cfg<kCtas, kRefine>;
  constexpr int kStages = C::kStages;
  // collect mode: only the query units that hold live rows do any work (the counts live on the device)
  auto m_active_of = [&](int item) -> int {
    if constexpr (!kCollect) return p.m_units;
    const int live = min(p.c_active[item], p.t);
    return (live + kBlockM * kCtas - 1) / (kBlockM * kCtas);
  };
  if constexpr (kCollect) {
    pdl_launch_dependents();                 // (before the wait: an idle fallback chain drains without launch gaps)
    pdl_wait();
    int any = 0;
    for (int i = 0; i < p.items; ++i) any |= p.c_active[i];
    if (any <= 0) return;                    // nothing fell back: uniform exit before any setup
  }

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024 B alignment
  const uint32_t smem_a = smem_base;                              // per stage: A hi [, A lo]
  const uint32_t smem_b = smem_base + kStages * C::kAStage;       // per stage: B hi [, B lo]
  const uint32_t scratch_off = C::kSmemData;                     // [32][256] floats
  const uint32_t bars = smem_base + C::kSmemData + C::kScratchBytes;
  const uint32_t bar_full = bars;                                // kStages x 8 B
  const uint32_t bar_empty = bars + 8 * kStages;                 // kStages x 8 B
  const uint32_t bar_tfull = bars + 16 * kStages;                // 2 x 8 B
  const uint32_t bar_tempty = bars + 16 * kStages + 16;          // 2 x 8 B
  const uint32_t tmem_slot = bars + 16 * kStages + 32;           // 4 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtas == 1) ? 0u : cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int unit_stride = gridDim.x / kCtas;
  const int first_unit = blockIdx.x / kCtas;
  const int units_per_item = p.m_units * p.segments;
  const int total_units = units_per_item * p.items;
  pdl_launch_dependents();        // the finish kernel may be scheduled now; it waits for this grid on the device

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_lib);
    if constexpr (kRefine) {
      tma_prefetch_desc(&tmap_q_lo);
      tma_prefetch_desc(&tmap_lib_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);    // the leader's arrive.expect_tx; bytes of both CTAs land here
      mbar_init(bar_empty + 8 * s, 1);   // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);                        // one tcgen05.commit
      mbar_init(bar_tempty + 8 * a, kNumEpiWarps * kCtas);    // one arrive per epilogue warp of the unit
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<kCtas>(tmem_slot, kTmemCols);
    tmem_relinquish<kCtas>();
  }
  tcgen05_fence_before();
  if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // the whole warp walks the loop (warp-uniform control flow), one elected lane issues
    {
      pdl_wait();   // launched behind the query pack of the same call: its output is complete from here on
      const uint32_t full0 = (kCtas == 1) ? bar_full : map_to_cta(bar_full, 0);   // barrier lives in the leader
      uint32_t it = 0;
      int my_tiles = 0, epoch = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        const int item = unit / units_per_item;
        const int rem = unit - item * units_per_item;
        const int m_unit = rem % p.m_units;
        if (kCollect && m_unit >= m_active_of(item)) continue;
        const int seg = rem / p.m_units;
        const int tile0 = seg * p.tiles_per_segment;
        const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
        // item i: query rows [i*t, (i+1)*t), library rows [i*n, (i+1)*n); rows of a tile that spill
        // into the next item (or past the end: TMA zero fill) are masked in the epilogue
        const int q_row = item * p.t + (m_unit * kCtas + static_cast<int>(cta_rank)) * kBlockM;
        const int lib_row0 = item * p.n;
        for (int tile = tile0; tile < tile1; ++tile, ++my_tiles) {
          if (p.sync_ctr != nullptr && epoch < p.sync_rounds && my_tiles == (epoch + 1) * p.sync_every) {
            // Pacing, not correctness: CTAs that stream the same library segment drift apart (all of
            // them are MMA-bound, nobody ever catches up) until a tile has left L2 before its last
            // reader arrives.  Re-align the producers now and then; give up after ~20 us.
            if (elect_one_sync()) {
              atomicAdd(p.sync_ctr, 1u);
              const unsigned int target = static_cast<unsigned int>(epoch + 1) * gridDim.x;
              const long long t0 = clock64();
              while (*reinterpret_cast<volatile unsigned int*>(p.sync_ctr) < target && clock64() - t0 < 40000) {
              }
            }
            __syncwarp();
            ++epoch;
          }
          const int lib_row = lib_row0 + tile * kBlockN + static_cast<int>(cta_rank) * C::kBRows;
          for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
            const uint32_t stage = it % kStages;
            const uint32_t phase = (it / kStages) & 1u;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (elect_one_sync()) {
              // debug 3 (timing experiment only, results are garbage): query tiles are loaded for the
              // first library tile of a unit only - an upper bound for what a resident query tile would buy
              const bool load_q = p.debug != 3 || tile == tile0;
              if (leader) mbar_expect_tx(bar_full + 8 * stage, load_q ? C::kTxBytes : C::kBStage * kCtas);
              if (load_q)
                tma_load_2d<kCtas>(smem_a + stage * C::kAStage, &tmap_q, full0 + 8 * stage, kb * kBlockK, q_row, p.hint_q);
              tma_load_2d<kCtas>(smem_b + stage * C::kBStage, &tmap_lib, full0 + 8 * stage, kb * kBlockK, lib_row, p.hint_lib);
              if constexpr (kRefine) {
                tma_load_2d<kCtas>(smem_a + stage * C::kAStage + kABytes, &tmap_q_lo, full0 + 8 * stage, kb * kBlockK, q_row, p.hint_q);
                tma_load_2d<kCtas>(smem_b + stage * C::kBStage + C::kBBytes, &tmap_lib_lo, full0 + 8 * stage, kb * kBlockK, lib_row,
                                   p.hint_lib);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    // leader CTA only; warp-uniform loop, one elected lane issues the tcgen05 instructions
    if (leader) {
      const uint32_t idesc = make_idesc(kBlockM * kCtas, kBlockN, p.half != 0);
      uint32_t it = 0, tile_count = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        if (kCollect && (unit % units_per_item) % p.m_units >= m_active_of(unit / units_per_item)) continue;
        const int seg = (unit % units_per_item) / p.m_units;
        const int tile0 = seg * p.tiles_per_segment;
        const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
        for (int tile = tile0; tile < tile1; ++tile, ++tile_count) {
          const uint32_t acc = tile_count & 1u;
          const uint32_t acc_phase = (tile_count >> 1) & 1u;
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);   // epilogue drained this accumulator
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
            const uint32_t stage = it % kStages;
            const uint32_t phase = (it / kStages) & 1u;
            mbar_wait(bar_full + 8 * stage, phase);
            tcgen05_fence_after();
            if (elect_one_sync()) {
              const uint64_t adesc = make_smem_desc(smem_a + stage * C::kAStage);
              const uint64_t bdesc = make_smem_desc(smem_b + stage * C::kBStage);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
                umma_bf16<kCtas>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              }
              if constexpr (kRefine) {
                // the two cross terms of (hi + lo).(hi + lo); lo.lo (<= |dq||dr| ~ 3e-6) is part of the error bound
                const uint64_t adesc_lo = make_smem_desc(smem_a + stage * C::kAStage + kABytes);
                const uint64_t bdesc_lo = make_smem_desc(smem_b + stage * C::kBStage + C::kBBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  umma_bf16<kCtas>(tmem_d, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1u);
                  umma_bf16<kCtas>(tmem_d, adesc_lo + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
              }
              umma_commit<kCtas>(bar_empty + 8 * stage);   // smem slot reusable once these MMAs retire
              if (kb == p.k_blocks - 1) umma_commit<kCtas>(bar_tfull + 8 * acc);   // accumulator complete -> epilogue
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ======================================= epilogue =======================================
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - kFirstEpiWarp) >> 2;       // which 128 columns of each tile
    const uint32_t tempty0 = (kCtas == 1) ? bar_tempty : map_to_cta(bar_tempty, 0);
    const uint32_t taddr_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    float* scratch = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + scratch_off) +
                     (threadIdx.x - kFirstEpiWarp * 32);
    uint32_t tile_count = 0;
    for (int unit = first_unit; unit < total_units; unit += unit_stride) {
      const int item = unit / units_per_item;
      const int rem = unit - item * units_per_item;
      const int m_unit = rem % p.m_units;
      if (kCollect && m_unit >= m_active_of(item)) continue;
      const int seg = rem / p.m_units;
      const int tile0 = seg * p.tiles_per_segment;
      const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
      const int row_in_item = (m_unit * kCtas + static_cast<int>(cta_rank)) * kBlockM + quarter * 32 + lane;
      const bool row_valid = row_in_item < (kCollect ? min(p.c_active[item], p.t) : p.t);
      const int row = item * p.t + row_in_item;          // global query index
      const int col_item0 = item * p.n;                  // frame indices are global: item*n + frame
      const int n_valid = col_item0 + p.n;

      float s[kListLen];
      uint32_t id[kListLen];
#pragma unroll
      for (int i = 0; i < kListLen; ++i) {
        s[i] = row_valid ? -INFINITY : INFINITY;   // padded query rows never insert
        id[i] = 0xFFFFFFFFu;
      }
      // collect mode: this row's cut; a row that has overflowed its buffer stops collecting
      float c_cut = INFINITY;
      bool c_live = false;
      if constexpr (kCollect) {
        if (row_valid) {
          c_cut = p.c_cut[row];
          c_live = true;
        }
      }

      for (int tile = tile0; tile < tile1; ++tile, ++tile_count) {
        const uint32_t acc = tile_count & 1u;
        const uint32_t acc_phase = (tile_count >> 1) & 1u;
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = taddr_base + acc * kBlockN;
        const int col0 = col_item0 + tile * kBlockN + half * 128;
        if (p.debug == 2) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtas == 1) mbar_arrive(bar_tempty + 8 * acc);
            else mbar_arrive_cluster(tempty0 + 8 * acc);
          }
          continue;
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + 32 * c, v);
          tmem_ld_wait(v);
          if (c == 3) {
            // every TMEM read of this accumulator is done: hand it back before the last scan
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (kCtas == 1) mbar_arrive(bar_tempty + 8 * acc);
              else mbar_arrive_cluster(tempty0 + 8 * acc);
            }
          }
          if constexpr (kCollect) {
            if (c_live) {
              uint32_t mask = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j)     // NaN scores are collected too (they rank first)
                mask |= (!(__uint_as_float(v[j]) < c_cut) && col0 + 32 * c + j < n_valid) ? (1u << j) : 0u;
              if (mask) {
                int at = atomicAdd(p.c_cnt + row, __popc(mask));
                if (at >= p.c_cap) c_live = false;
                while (mask && at < p.c_cap) {
                  const int j = __ffs(static_cast<int>(mask)) - 1;
                  mask &= mask - 1;
                  p.c_idx[static_cast<size_t>(row) * p.c_cap + at++] = col0 + 32 * c + j;
                }
              }
            }
          } else if constexpr (!kRefine) {
            if (p.debug == 0 || p.debug == 3) scan_chunk(v, col0 + 32 * c, n_valid, s, id, scratch);
            else s[0] = fmaxf(s[0], __uint_as_float(v[0] ^ v[13] ^ v[31]));
          }
        }
      }

      if (!kCollect && row_valid) {
        const size_t o = (static_cast<size_t>(row) * p.lists + static_cast<size_t>(seg * 2 + half)) * kListLen;
        float4* ps = reinterpret_cast<float4*>(p.cand_score + o);
        int4* pi = reinterpret_cast<int4*>(p.cand_idx + o);
        ps[0] = make_float4(s[0], s[1], s[2], s[3]);
        ps[1] = make_float4(s[4], s[5], s[6], s[7]);
        pi[0] = make_int4(static_cast<int>(id[0]), static_cast<int>(id[1]), static_cast<int>(id[2]), static_cast<int>(id[3]));
        pi[1] = make_int4(static_cast<int>(id[4]), static_cast<int>(id[5]), static_cast<int>(id[6]), static_cast<int>(id[7]));
      }
    }
  }

  // ------------------------------------------ teardown ------------------------------------------
  tcgen05_fence_before();
  if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<kCtas>(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel with a (partly) RESIDENT query block, specialised for a compile-time number of channel blocks
// (kKB = d/64; 12 for the 768-channel features of the reference).
//
// Why: under ncu the kernel above keeps the tensor pipe 98 % busy - what bounds it is the SM clock the 1 kW power cap
// allows, and a good part of the power moves operands: 1.23 TB per cfg4 launch cross the L2->SM crossbar, half of
// them the SAME 128 x 768 query block fetched again for every library tile (an experiment that skipped those reloads
// ran 8 % faster at higher clocks).  Here the first kRes channel blocks of the unit's query rows are loaded ONCE per
// unit and stay in shared memory; the ring carries 16 KB slots: one (library) per resident channel block, two (query,
// library) per streamed one - kUses = kRes + 2 (kKB - kRes) slot loads per tile instead of 2 kKB.
//
// How (measured at cfg4, ms per step, kernel above = 105.9): what decides is the look-ahead of the ring in CHANNEL
// BLOCKS.  Six slots behind six leading resident blocks cover 6 blocks in the resident half of a tile but only 3 in
// the streamed half: 109.1 ms with run-time slot arithmetic, 107.8 ms fully static - slower than no residency, at
// HIGHER clocks (the tensor pipe idles).  Eight resident + four slots: 117.8 ms.  Making every OTHER channel block
// resident (kInterleave) gives a uniform demand of 3 slots per 2 blocks = 4 blocks of look-ahead everywhere:
// 102.5 ms, +4.3 % over the kernel above.  Everything is static: kUses is a multiple of kSlots, every use of every
// tile therefore hits the same slot, the loops over channel blocks are fully unrolled with constant addresses, and
// the only run-time state is one parity bit per tile.
//   smem: A_res [kRes][16 KB] | slots [kSlots][16 KB] | epilogue scratch 32 KB | barriers
//   barriers: full/empty per slot, tfull/tempty per accumulator, ares_full (the resident block has landed),
//             ares_free (every MMA of the unit that read it has retired: the next unit's block may overwrite it)
// Same MMA sequence, same accumulators, same epilogue as above: bit-identical lists.
// ---------------------------------------------------------------------------------------------
template <int kKB, int kRes, int kSlots, bool kInterleave = false> struct ResCfg {
  static constexpr int kUses = kRes + 2 * (kKB - kRes);           // slot loads per library tile
  static constexpr int kRounds = kUses / kSlots;                   // trips around the ring per tile
  static_assert(kRes <= kKB && kUses % kSlots == 0, "resident kernel: the slot pattern must repeat every tile");
  static_assert(!kInterleave || kKB == 2 * kRes, "interleaved pattern: every other channel block is resident");
  static constexpr uint32_t kSmemData = (kRes + kSlots) * kABytes;
  static constexpr uint32_t kScratch = 32 * kNumEpiWarps * 32 * 4;
  static constexpr uint32_t kSmemBytes = kSmemData + kScratch + 1024 + 256;
  static_assert(kSmemBytes <= 227 * 1024, "resident kernel: shared memory");
  // which channel blocks are resident: the first kRes, or (kInterleave) the even ones - a uniform slot demand
  __host__ __device__ static constexpr bool is_res(int kb) { return kInterleave ? (kb & 1) == 0 : kb < kRes; }
  __host__ __device__ static constexpr int res_idx(int kb) { return kInterleave ? kb / 2 : kb; }
  // use index (position in the per-tile load sequence) of the library (B) load of channel block kb, and of its
  // query (A) load when that block streams
  __host__ __device__ static constexpr int use_b(int kb) {
    return kInterleave ? 3 * (kb / 2) + ((kb & 1) ? 2 : 0) : (kb < kRes ? kb : kRes + 2 * (kb - kRes) + 1);
  }
  __host__ __device__ static constexpr int use_a(int kb) { return kInterleave ? 3 * (kb / 2) + 1 : kRes + 2 * (kb - kRes); }
  // the inverse: channel block and operand of use u
  __host__ __device__ static constexpr int use_kb(int u) {
    return kInterleave ? 2 * (u / 3) + (u % 3 != 0 ? 1 : 0) : (u < kRes ? u : kRes + (u - kRes) / 2);
  }
  __host__ __device__ static constexpr bool use_is_a(int u) {
    return kInterleave ? u % 3 == 1 : (u >= kRes && ((u - kRes) & 1) == 0);
  }
};

__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {   // small at the (unrolled) call site
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

template <int kKB, int kRes, int kSlots, bool kInterleave>
__global__ void __launch_bounds__(kThreads, 1)
knn_search_resident_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_lib,
                           const SearchParams p) {
  constexpr int kCtas = 2;
  using RC = ResCfg<kKB, kRes, kSlots, kInterleave>;
  constexpr uint32_t kSlotBytes = kABytes;                        // 16 KB: 128 rows x 64 channels of either operand
  constexpr int kBRows = kBlockN / kCtas;                         // library rows this CTA loads per tile

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_res = smem_base;
  const uint32_t smem_slot = smem_base + kRes * kSlotBytes;
  const uint32_t scratch_off = RC::kSmemData;
  const uint32_t bars = smem_base + RC::kSmemData + RC::kScratch;
  const uint32_t bar_full = bars;                                // kSlots x 8 B
  const uint32_t bar_empty = bars + 8 * kSlots;                  // kSlots x 8 B
  const uint32_t bar_tfull = bars + 16 * kSlots;                 // 2 x 8 B
  const uint32_t bar_tempty = bars + 16 * kSlots + 16;           // 2 x 8 B
  const uint32_t bar_ares_full = bars + 16 * kSlots + 32;        // 8 B
  const uint32_t bar_ares_free = bars + 16 * kSlots + 40;        // 8 B
  const uint32_t tmem_slot = bars + 16 * kSlots + 48;            // 4 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int unit_stride = gridDim.x / kCtas;
  const int first_unit = blockIdx.x / kCtas;
  const int units_per_item = p.m_units * p.segments;
  const int total_units = units_per_item * p.items;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_lib);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kNumEpiWarps * kCtas);
    }
    mbar_init(bar_ares_full, 1);
    mbar_init(bar_ares_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<kCtas>(tmem_slot, kTmemCols);
    tmem_relinquish<kCtas>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    pdl_wait();
    const uint32_t full0 = map_to_cta(bar_full, 0);              // the full barriers live in the leader
    const uint32_t ares_full0 = map_to_cta(bar_ares_full, 0);
    int my_tiles = 0, epoch = 0, my_units = 0;
    for (int unit = first_unit; unit < total_units; unit += unit_stride, ++my_units) {
      const int item = unit / units_per_item;
      const int rem = unit - item * units_per_item;
      const int m_unit = rem % p.m_units;
      const int seg = rem / p.m_units;
      const int tile0 = seg * p.tiles_per_segment;
      const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
      const int q_row = item * p.t + (m_unit * kCtas + static_cast<int>(cta_rank)) * kBlockM;
      const int lib_row0 = item * p.n;
      // the resident block of this unit: the previous unit's MMAs must have retired before it is overwritten
      if (my_units > 0) mbar_wait(bar_ares_free, (static_cast<uint32_t>(my_units) - 1u) & 1u);
      if (elect_one_sync()) {
        if (leader) mbar_expect_tx(bar_ares_full, kRes * kSlotBytes * kCtas);
#pragma unroll
        for (int kb = 0; kb < kKB; ++kb)
          if (RC::is_res(kb))
            tma_load_2d<kCtas>(smem_res + RC::res_idx(kb) * kSlotBytes, &tmap_q, ares_full0, kb * kBlockK, q_row, p.hint_q);
      }
      __syncwarp();
      for (int tile = tile0; tile < tile1; ++tile, ++my_tiles) {
        if (p.sync_ctr != nullptr && epoch < p.sync_rounds && my_tiles == (epoch + 1) * p.sync_every) {
          // pacing, not correctness (see knn_search_kernel)
          if (elect_one_sync()) {
            atomicAdd(p.sync_ctr, 1u);
            const unsigned int target = static_cast<unsigned int>(epoch + 1) * gridDim.x;
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(p.sync_ctr) < target && clock64() - t0 < 40000) {
            }
          }
          __syncwarp();
          ++epoch;
        }
        const int lib_row = lib_row0 + tile * kBlockN + static_cast<int>(cta_rank) * kBRows;
        // parity of use u of this CTA's n-th tile: ring round n * kRounds + u / kSlots
        const uint32_t tile_par = (static_cast<uint32_t>(my_tiles) * RC::kRounds) & 1u;
#pragma unroll
        for (int u = 0; u < RC::kUses; ++u) {
          const int slot = u % kSlots;
          const uint32_t par = (tile_par + static_cast<uint32_t>(u / kSlots)) & 1u;
          const bool is_a = RC::use_is_a(u);
          const int kb = RC::use_kb(u);
          mbar_wait_lean(bar_empty + 8 * slot, par ^ 1u);
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(bar_full + 8 * slot, kSlotBytes * kCtas);
            if (is_a) tma_load_2d<kCtas>(smem_slot + slot * kSlotBytes, &tmap_q, full0 + 8 * slot, kb * kBlockK, q_row, p.hint_q);
            else tma_load_2d<kCtas>(smem_slot + slot * kSlotBytes, &tmap_lib, full0 + 8 * slot, kb * kBlockK, lib_row, p.hint_lib);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (leader) {
      const uint32_t idesc = make_idesc(kBlockM * kCtas, kBlockN, p.half != 0);
      const uint64_t desc_res = make_smem_desc(smem_res);        // +1024 per 16 KB (the address field counts 16 B)
      const uint64_t desc_slot = make_smem_desc(smem_slot);
      uint32_t tile_count = 0;
      int my_units = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride, ++my_units) {
        const int seg = (unit % units_per_item) / p.m_units;
        const int tile0 = seg * p.tiles_per_segment;
        const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
        mbar_wait(bar_ares_full, static_cast<uint32_t>(my_units) & 1u);       // the unit's resident query block
        tcgen05_fence_after();
        for (int tile = tile0; tile < tile1; ++tile, ++tile_count) {
          const uint32_t acc = tile_count & 1u;
          const uint32_t acc_phase = (tile_count >> 1) & 1u;
          const uint32_t tile_par = (tile_count * RC::kRounds) & 1u;
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBlockN;
#pragma unroll
          for (int kb = 0; kb < kKB; ++kb) {
            constexpr uint32_t kDescStep = kSlotBytes >> 4;
            const bool streamed = !RC::is_res(kb);
            const int ub = RC::use_b(kb);
            const int ua = RC::use_a(kb);
            if (streamed) mbar_wait_lean(bar_full + 8 * (ua % kSlots), (tile_par + static_cast<uint32_t>(ua / kSlots)) & 1u);
            mbar_wait_lean(bar_full + 8 * (ub % kSlots), (tile_par + static_cast<uint32_t>(ub / kSlots)) & 1u);
            tcgen05_fence_after();
            if (elect_one_sync()) {
              const uint64_t adesc = streamed ? desc_slot + static_cast<uint64_t>((ua % kSlots) * kDescStep)
                                              : desc_res + static_cast<uint64_t>(RC::res_idx(kb) * kDescStep);
              const uint64_t bdesc = desc_slot + static_cast<uint64_t>((ub % kSlots) * kDescStep);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                umma_bf16<kCtas>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              if (streamed) umma_commit<kCtas>(bar_empty + 8 * (ua % kSlots));
              umma_commit<kCtas>(bar_empty + 8 * (ub % kSlots));
              if (kb == kKB - 1) {
                umma_commit<kCtas>(bar_tfull + 8 * acc);
                if (tile == tile1 - 1) umma_commit<kCtas>(bar_ares_free);      // the resident block may be replaced
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ======================================= epilogue (as in knn_search_kernel) =======================================
    const int quarter = warp & 3;
    const int half = (warp - kFirstEpiWarp) >> 2;
    const uint32_t tempty0 = map_to_cta(bar_tempty, 0);
    const uint32_t taddr_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    float* scratch = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + scratch_off) +
                     (threadIdx.x - kFirstEpiWarp * 32);
    uint32_t tile_count = 0;
    for (int unit = first_unit; unit < total_units; unit += unit_stride) {
      const int item = unit / units_per_item;
      const int rem = unit - item * units_per_item;
      const int m_unit = rem % p.m_units;
      const int seg = rem / p.m_units;
      const int tile0 = seg * p.tiles_per_segment;
      const int tile1 = min(tile0 + p.tiles_per_segment, p.n_tiles);
      const int row_in_item = (m_unit * kCtas + static_cast<int>(cta_rank)) * kBlockM + quarter * 32 + lane;
      const bool row_valid = row_in_item < p.t;
      const int row = item * p.t + row_in_item;
      const int col_item0 = item * p.n;
      const int n_valid = col_item0 + p.n;
      float s[kListLen];
      uint32_t id[kListLen];
#pragma unroll
      for (int i = 0; i < kListLen; ++i) {
        s[i] = row_valid ? -INFINITY : INFINITY;
        id[i] = 0xFFFFFFFFu;
      }
      for (int tile = tile0; tile < tile1; ++tile, ++tile_count) {
        const uint32_t acc = tile_count & 1u;
        const uint32_t acc_phase = (tile_count >> 1) & 1u;
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = taddr_base + acc * kBlockN;
        const int col0 = col_item0 + tile * kBlockN + half * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + 32 * c, v);
          tmem_ld_wait(v);
          if (c == 3) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);
          }
          scan_chunk(v, col0 + 32 * c, n_valid, s, id, scratch);
        }
      }
      if (row_valid) {
        const size_t o = (static_cast<size_t>(row) * p.lists + static_cast<size_t>(seg * 2 + half)) * kListLen;
        float4* ps = reinterpret_cast<float4*>(p.cand_score + o);
        int4* pi = reinterpret_cast<int4*>(p.cand_idx + o);
        ps[0] = make_float4(s[0], s[1], s[2], s[3]);
        ps[1] = make_float4(s[4], s[5], s[6], s[7]);
        pi[0] = make_int4(static_cast<int>(id[0]), static_cast<int>(id[1]), static_cast<int>(id[2]), static_cast<int>(id[3]));
        pi[1] = make_int4(static_cast<int>(id[4]), static_cast<int>(id[5]), static_cast<int>(id[6]), static_cast<int>(id[7]));
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<kCtas>(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// "skinny" kernel: at most 32 query frames (one realtime chunk, BASELINE cfg2).  The library streams
// through ONCE and nothing else is worth staging, so the roles of the operands are swapped:
//   A (M = 128) = a 128-frame library tile, 16 KB per 64-channel block, a deep TMA ring of its own;
//   B (N = 32)  = the query chunk, resident in shared memory for the whole launch (d/64 x 4 KB).
// Every byte of the ring is library data (the tiled kernel re-fetches a 16 KB query tile with every
// 32 KB of library), so ~25% more HBM traffic is in flight per SM - this regime is bound by exactly that.
// The accumulator is [128 frames (TMEM lanes) x 32 queries (columns)]: 16 of them fit in TMEM, the
// MMA warp runs up to 16 tiles ahead.  An epilogue warp reads its 32-lane quarter (thread = frame),
// transposes the 32 x 32 block through shared memory (thread = query) and folds it into the
// per-query top list it keeps in registers; the four warps' lists are merged at the end, so the
// launch leaves ONE list per CTA and query: lists = grid.
// ---------------------------------------------------------------------------------------------
constexpr int kSkinnyQ = 32;                                   // query rows (MMA N)
constexpr int kSkinnyThreads = 256;                            // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 epilogue
constexpr int kSkinnySlots = 16;                               // accumulators of 32 TMEM columns
constexpr uint32_t kSkinnyQBytes = kSkinnyQ * kBlockK * 2;     // 4 KB per channel block
constexpr uint32_t kSkinnyTbFloats = 32 * 33;                  // transpose buffer per epilogue warp
constexpr uint32_t kSkinnySmemMax = 227 * 1024;

struct SkinnyParams {
  int t, n, k_blocks, n_tiles, stages, lists, contig, half;
  float* cand_score;
  int* cand_idx;
};

__host__ __device__ inline uint32_t skinny_fixed_bytes(int k_blocks) {
  return static_cast<uint32_t>(k_blocks) * kSkinnyQBytes + 4 * kSkinnyTbFloats * 4 + 1024 /*align slack*/ + 1024 /*barriers*/;
}

__global__ void __launch_bounds__(kSkinnyThreads, 1)
knn_search_skinny_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_lib,
                         const SkinnyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;                                              // [k_blocks][32 x 128 B]
  const uint32_t smem_a = smem_q + static_cast<uint32_t>(p.k_blocks) * kSkinnyQBytes;   // [stages][128 x 128 B]
  const uint32_t tb_off = static_cast<uint32_t>(p.k_blocks) * kSkinnyQBytes + static_cast<uint32_t>(p.stages) * kABytes;
  const uint32_t bars = smem_base + tb_off + 4 * kSkinnyTbFloats * 4;
  const uint32_t bar_full = bars;                                   // stages x 8 B
  const uint32_t bar_empty = bars + 8 * p.stages;                   // stages x 8 B
  const uint32_t bar_tfull = bars + 16 * p.stages;                  // 16 x 8 B
  const uint32_t bar_tempty = bar_tfull + 8 * kSkinnySlots;         // 16 x 8 B
  const uint32_t bar_q = bar_tempty + 8 * kSkinnySlots;             // 8 B
  const uint32_t tmem_slot = bar_q + 8;                             // 4 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();        // the finish kernel may be scheduled now; it waits for this grid on the device

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_lib);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < kSkinnySlots; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);      // one tcgen05.commit
      mbar_init(bar_tempty + 8 * a, 4);     // one arrive per epilogue warp
    }
    mbar_init(bar_q, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, kTmemCols);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // the library does not depend on the query pack that may still be running in front of this launch
    // (programmatic dependent launch): start streaming right away
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(bar_full + 8 * stage, kABytes);
          if (p.contig)
            tma_load_2d<1>(smem_a + stage * kABytes, &tmap_lib, bar_full + 8 * stage, 0, (tile * p.k_blocks + kb) * kBlockM,
                           kL2EvictFirst);
          else
            tma_load_2d<1>(smem_a + stage * kABytes, &tmap_lib, bar_full + 8 * stage, kb * kBlockK, tile * kBlockM,
                           kL2EvictFirst);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(p.stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    const uint32_t idesc = make_idesc(kBlockM, kSkinnyQ, p.half != 0);
    mbar_wait(bar_q, 0);
    tcgen05_fence_after();
    uint32_t stage = 0, phase = 0, slot = 0, slot_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(bar_tempty + 8 * slot, slot_phase ^ 1u);   // the epilogue drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + slot * kSkinnyQ;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint64_t adesc = make_smem_desc(smem_a + stage * kABytes);
          const uint64_t bdesc = make_smem_desc(smem_q + kb * kSkinnyQBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16<1>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit<1>(bar_empty + 8 * stage);
          if (kb == p.k_blocks - 1) umma_commit<1>(bar_tfull + 8 * slot);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(p.stages)) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++slot == kSkinnySlots) {
        slot = 0;
        slot_phase ^= 1u;
      }
    }
  } else if (warp == 3) {
    // ================================ query chunk (resident B operand) ================================
    pdl_wait();                                                     // the query pack of this call has finished
    if (elect_one_sync()) {
      mbar_expect_tx(bar_q, static_cast<uint32_t>(p.k_blocks) * kSkinnyQBytes);
      for (int kb = 0; kb < p.k_blocks; ++kb)
        tma_load_2d<1>(smem_q + kb * kSkinnyQBytes, &tmap_q, bar_q, kb * kBlockK, 0, kL2EvictNormal);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int quarter = warp & 3;                                   // TMEM lanes 32*quarter .. +31
    float* tb = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + tb_off) + quarter * kSkinnyTbFloats;
    const bool q_valid = lane < p.t;                                // lane = query once transposed
    float s[kListLen];
    uint32_t id[kListLen];
#pragma unroll
    for (int i = 0; i < kListLen; ++i) {
      s[i] = q_valid ? -INFINITY : INFINITY;                        // padded queries never insert
      id[i] = 0xFFFFFFFFu;
    }
    uint32_t slot = 0, slot_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(bar_tfull + 8 * slot, slot_phase);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * kSkinnyQ, v);
      tmem_ld_wait(v);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
      if (++slot == kSkinnySlots) {
        slot = 0;
        slot_phase ^= 1u;
      }
      // thread = frame (32*quarter + lane of this tile) -> thread = query
      const int row0 = tile * kBlockM + quarter * 32;
      const bool row_ok = row0 + lane < p.n;                        // ragged last tile: TMA zero fill
#pragma unroll
      for (int j = 0; j < 32; ++j) tb[lane * 33 + j] = row_ok ? __uint_as_float(v[j]) : -INFINITY;
      __syncwarp();
      float m = tb[lane];
#pragma unroll
      for (int r = 1; r < 32; ++r) m = fmaxf(m, tb[r * 33 + lane]);
      if (m > s[kListLen - 1]) {
#pragma unroll 1
        for (int r = 0; r < 32; ++r) {
          const float x = tb[r * 33 + lane];
          if (x > s[kListLen - 1]) list_insert(s, id, x, static_cast<uint32_t>(row0 + r));
        }
      }
      __syncwarp();
    }
    // ---- merge the four quarters' lists: every warp parks its lists, warp 4 folds them ----
    // (named barrier 1 over the 128 epilogue threads; the transpose buffers are free now)
    float* park_s = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + tb_off);   // [4][8][32]
    uint32_t* park_i = reinterpret_cast<uint32_t*>(park_s + 4 * kListLen * 32);                            // [4][8][32]
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int i = 0; i < kListLen; ++i) {
      park_s[(quarter * kListLen + i) * 32 + lane] = s[i];
      park_i[(quarter * kListLen + i) * 32 + lane] = id[i];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (quarter == 0 && q_valid) {
      pdl_wait();      // (already satisfied: the accumulators needed the query chunk) orders the list stores
#pragma unroll 1
      for (int e = kListLen; e < 4 * kListLen; ++e) {
        const float x = park_s[e * 32 + lane];
        if (x > s[kListLen - 1]) list_insert(s, id, x, park_i[e * 32 + lane]);
      }
      const size_t o = (static_cast<size_t>(lane) * p.lists + blockIdx.x) * kListLen;
      float4* ps = reinterpret_cast<float4*>(p.cand_score + o);
      int4* pi = reinterpret_cast<int4*>(p.cand_idx + o);
      ps[0] = make_float4(s[0], s[1], s[2], s[3]);
      ps[1] = make_float4(s[4], s[5], s[6], s[7]);
      pi[0] = make_int4(static_cast<int>(id[0]), static_cast<int>(id[1]), static_cast<int>(id[2]), static_cast<int>(id[3]));
      pi[1] = make_int4(static_cast<int>(id[4]), static_cast<int>(id[5]), static_cast<int>(id[6]), static_cast<int>(id[7]));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<1>(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows, d] bf16 row-major -> 2-D map, box = 64 elements (128 B) x box_rows, 128B swizzle, OOB = 0
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t d, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return -3;
  }
  cuuint64_t gdim[2] = {d, rows};
  cuuint64_t gstride[1] = {d * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu d=%llu box_rows=%u base=%p)",
              static_cast<int>(r), static_cast<unsigned long long>(rows), static_cast<unsigned long long>(d),
              box_rows, base);
    return -3;
  }
  return 0;
}

// pacing counters: a small ring of device words per DEVICE, one word per in-flight launch
unsigned int* pacing_slot(cudaStream_t stream) {
  static std::mutex mu;
  static unsigned int* base[64] = {};
  static std::atomic<unsigned int> seq{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  unsigned int* b;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev] && cudaMalloc(&base[dev], 64 * sizeof(unsigned int)) != cudaSuccess) {
      base[dev] = nullptr;
      return nullptr;
    }
    b = base[dev];
  }
  unsigned int* slot = b + (seq.fetch_add(1, std::memory_order_relaxed) % 64);
  if (cudaMemsetAsync(slot, 0, sizeof(unsigned int), stream) != cudaSuccess) return nullptr;
  return slot;
}

struct CollectArgs {
  const int* active;
  const float* cut;
  int* cnt;
  int* idx;
  int cap;
  const uint16_t* q_lo;      // second bf16 planes (refined accumulation), or NULL
  const uint16_t* lib_lo;
};

template <int kCtas, bool kCollect = false, bool kRefine = false>
int launch_search(const uint16_t* q, const uint16_t* lib, const alive_knn_plan_t& plan, float* cand_score,
                  int32_t* cand_idx, int after_query_pack, cudaStream_t stream, const CollectArgs* ca = nullptr) {
  using C = Cfg<kCtas, kRefine>;
  CUtensorMap mq, ml, mq_lo, ml_lo;
  const uint64_t items = static_cast<uint64_t>(plan.items < 1 ? 1 : plan.items);
  int rc = make_map(&mq, q, items * static_cast<uint64_t>(plan.t), static_cast<uint64_t>(plan.d), kBlockM);
  if (rc) return rc;
  rc = make_map(&ml, lib, items * static_cast<uint64_t>(plan.n), static_cast<uint64_t>(plan.d), C::kBRows);
  if (rc) return rc;
  mq_lo = mq;
  ml_lo = ml;
  if constexpr (kRefine) {
    rc = make_map(&mq_lo, ca->q_lo, items * static_cast<uint64_t>(plan.t), static_cast<uint64_t>(plan.d), kBlockM);
    if (rc) return rc;
    rc = make_map(&ml_lo, ca->lib_lo, items * static_cast<uint64_t>(plan.n), static_cast<uint64_t>(plan.d), C::kBRows);
    if (rc) return rc;
  }
  SearchParams p;
  p.items = static_cast<int>(items);
  p.t = plan.t;
  p.n = static_cast<int>(plan.n);
  p.half = plan.format == ALIVE_KNN_FORMAT_FP16 ? 1 : 0;
  p.k_blocks = plan.d / kBlockK;
  p.m_units = plan.m_units;
  p.segments = plan.segments;
  p.tiles_per_segment = plan.tiles_per_segment;
  p.n_tiles = plan.n_tiles;
  p.lists = plan.lists;
  p.cand_score = cand_score;
  p.cand_idx = cand_idx;
  {
    const char* dbg = getenv("ALIVE_KNN_DEBUG_EPILOGUE");
    p.debug = dbg ? atoi(dbg) : 0;
    // defaults: query tiles are re-read for every library tile (keep), library tiles stream
    auto pick = [](const char* name, unsigned long long dflt) -> unsigned long long {
      const char* v = getenv(name);
      if (!v) return dflt;
      switch (atoi(v)) {
        case 1: return kL2EvictFirst;
        case 2: return kL2EvictLast;
        default: return kL2EvictNormal;
      }
    };
    p.hint_q = pick("ALIVE_KNN_HINT_Q", kL2EvictNormal);
    p.hint_lib = pick("ALIVE_KNN_HINT_LIB", kL2EvictNormal);
    // producer pacing (see the kernel): only worth it when several waves of units share segments
    const char* se = getenv("ALIVE_KNN_SYNC_EVERY");
    p.sync_every = se ? atoi(se) : 128;   // measured at cfg4: DRAM reads 57 -> 23 GB per launch, +6.6 % throughput
    p.sync_ctr = nullptr;
    p.sync_rounds = 0;
    const long long total_units = static_cast<long long>(plan.m_units) * plan.segments * static_cast<long long>(items);
    const long long clusters = plan.grid / kCtas;
    const int last_seg_tiles = plan.n_tiles - (plan.segments - 1) * plan.tiles_per_segment;
    const long long min_tiles = (total_units / clusters) * (last_seg_tiles < plan.tiles_per_segment ? last_seg_tiles : plan.tiles_per_segment);
    p.c_active = nullptr;
    p.c_cut = nullptr;
    p.c_cnt = nullptr;
    p.c_idx = nullptr;
    p.c_cap = 0;
    if constexpr (kCollect) {
      p.c_active = ca->active;
      p.c_cut = ca->cut;
      p.c_cnt = ca->cnt;
      p.c_idx = ca->idx;
      p.c_cap = ca->cap;
    }
    // (no pacing in collect mode: units are skipped by a device-side count, CTAs would wait for absentees)
    if (!kCollect && p.sync_every > 0 && (plan.m_units > 1 || items > 1) && min_tiles / p.sync_every >= 1) {
      p.sync_ctr = pacing_slot(stream);
      p.sync_rounds = static_cast<int>(min_tiles / p.sync_every);
      if ((static_cast<long long>(p.sync_rounds) + 1) * plan.grid >= (1ll << 32)) p.sync_ctr = nullptr;
    }
  }

  // CTA pairs that visit several library tiles per unit keep (part of) their query block resident
  bool resident = false;
  if constexpr (kCtas == 2 && !kCollect && !kRefine) {
    const char* rs = getenv("ALIVE_KNN_RESIDENT");
    resident = (rs ? atoi(rs) != 0 : true) && p.debug == 0 && plan.tiles_per_segment >= 2 && plan.d == 768;
  }
  if (resident) {
    const char* rs = getenv("ALIVE_KNN_RESIDENT");
    const int rv = rs ? atoi(rs) : 1;
    auto go = [&](auto kern, uint32_t smem, PerDeviceOnce& once) -> int {
      const int rc_attr = once.run([&]() -> int {
        ALIVE_CHECK_CUDA((cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))));
        return 0;
      });
      if (rc_attr) return rc_attr;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(plan.grid));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = after_query_pack ? 2 : 1;
      ALIVE_CHECK_CUDA((cudaLaunchKernelEx(&cfg, kern, mq, ml, p)));
      return 0;
    };
    static PerDeviceOnce o1, o2;
    // ALIVE_KNN_RESIDENT=2: the first six channel blocks resident instead of every other one (kept for the A/B in
    // 3 channel blocks of look-ahead in the streamed half of each tile, 1.8 % SLOWER than no residency)
    if (rv == 2) return go(knn_search_resident_kernel<12, 6, 6, false>, ResCfg<12, 6, 6, false>::kSmemBytes, o2);
    return go(knn_search_resident_kernel<12, 6, 6, true>, ResCfg<12, 6, 6, true>::kSmemBytes, o1);
  }

  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(knn_search_kernel<kCtas, kCollect, kRefine>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::kSmemBytes))));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(plan.grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = after_query_pack ? 2 : 1;
  ALIVE_CHECK_CUDA((cudaLaunchKernelEx(&cfg, knn_search_kernel<kCtas, kCollect, kRefine>, mq, ml, mq_lo, ml_lo, p)));
  return 0;
}

int launch_skinny(const uint16_t* q, const uint16_t* lib, const alive_knn_plan_t& plan, float* cand_score,
                  int32_t* cand_idx, int after_query_pack, cudaStream_t stream) {
  ALIVE_REQUIRE(plan.items == 1 && plan.t >= 1 && plan.t <= kSkinnyQ && plan.grid >= 1 && plan.lists == plan.grid &&
                    plan.n_tiles == static_cast<int32_t>((plan.n + kBlockM - 1) / kBlockM) && plan.grid <= plan.n_tiles,
                "alive_knn_search: bad skinny plan");
  CUtensorMap mq, ml;
  int rc = make_map(&mq, q, static_cast<uint64_t>(plan.t), static_cast<uint64_t>(plan.d), kSkinnyQ);
  if (rc) return rc;
  rc = make_map(&ml, lib, static_cast<uint64_t>(plan.n), static_cast<uint64_t>(plan.d), kBlockM);
  if (rc) return rc;
  const char* contig = getenv("ALIVE_KNN_SKINNY_CONTIG");   // timing experiment only (results are garbage):
  const bool contig_exp = contig && atoi(contig) == 1;      // every 16 KB stage is ONE contiguous chunk of HBM
  if (contig_exp) {
    rc = make_map(&ml, lib, static_cast<uint64_t>(plan.n) * (plan.d / kBlockK), kBlockK, kBlockM);
    if (rc) return rc;
  }
  SkinnyParams p;
  p.contig = contig_exp ? 1 : 0;
  p.half = plan.format == ALIVE_KNN_FORMAT_FP16 ? 1 : 0;
  p.t = plan.t;
  p.n = static_cast<int>(plan.n);
  p.k_blocks = plan.d / kBlockK;
  p.n_tiles = plan.n_tiles;
  p.lists = plan.lists;
  p.cand_score = cand_score;
  p.cand_idx = cand_idx;
  const uint32_t fixed = skinny_fixed_bytes(p.k_blocks);
  ALIVE_REQUIRE(fixed + 4 * kABytes <= kSkinnySmemMax, "alive_knn_search: d = %d is too wide for the skinny kernel", plan.d);
  int stages = static_cast<int>((kSkinnySmemMax - fixed) / kABytes);
  if (stages > 16) stages = 16;
  {
    const char* st = getenv("ALIVE_KNN_SKINNY_STAGES");   // experiments: shallower ring
    if (st && atoi(st) >= 2 && atoi(st) < stages) stages = atoi(st);
  }
  p.stages = stages;
  const size_t smem = fixed + static_cast<size_t>(stages) * kABytes;
  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA(cudaFuncSetAttribute(knn_search_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(kSkinnySmemMax)));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(plan.grid));
  cfg.blockDim = dim3(kSkinnyThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = after_query_pack ? 1 : 0;
  ALIVE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, knn_search_skinny_kernel, mq, ml, p));
  return 0;
}

}  // namespace
}  // namespace alive

extern "C" int alive_knn_plan_batched(int32_t items, int32_t t, int64_t n, int32_t d, int32_t num_sms, int32_t variant,
                                      alive_knn_plan_t* plan) {
  using namespace alive;
  ALIVE_REQUIRE(plan != nullptr, "alive_knn_plan: plan is NULL");
  ALIVE_REQUIRE(items >= 1, "alive_knn_plan: items must be >= 1 (got %d)", items);
  ALIVE_REQUIRE(t >= 1, "alive_knn_plan: t must be >= 1 (got %d)", t);
  ALIVE_REQUIRE(n >= 1 && n < (1ll << 31) - 512, "alive_knn_plan: n out of range (%lld)", static_cast<long long>(n));
  ALIVE_REQUIRE(static_cast<long long>(items) * n < (1ll << 31) - 512 && static_cast<long long>(items) * t < (1ll << 31) - 512,
                "alive_knn_plan: items * n and items * t must stay below 2^31");
  // d <= 1536: the exact scan behind every screened call (alive_knn_exact) and the certificate's accumulation
  // slack (select.cu accum_slack) are both sized for at most 1536 channels - refuse here, before anything is launched
  ALIVE_REQUIRE(d >= 64 && d % 64 == 0 && d <= 1536, "alive_knn_plan: d must be a multiple of 64, <= 1536 (got %d)", d);
  ALIVE_REQUIRE(num_sms >= 2, "alive_knn_plan: num_sms must be >= 2");
  // default: the CTA-pair kernel (cta_group::2) once there is more than one 128-query tile - it
  // moves a third less operand data per flop and is ~10% faster under the power cap; a single
  // tile (streaming chunks) runs on one CTA per unit
  const int skinny_stages = static_cast<int>((kSkinnySmemMax - skinny_fixed_bytes(d / kBlockK)) / kABytes);
  const bool skinny_ok = items == 1 && t <= kSkinnyQ && skinny_stages >= 4;
  if (variant == 0) {
    const char* sk = getenv("ALIVE_KNN_SKINNY");      // A/B switch: 0 keeps the tiled kernel for small chunks
    variant = (skinny_ok && !(sk && atoi(sk) == 0)) ? 3 : (t > kBlockM) ? 2 : 1;
  }
  ALIVE_REQUIRE(variant >= 1 && variant <= 3, "alive_knn_plan: variant must be 0, 1, 2 or 3");
  plan->kernel = 0;
  plan->format = ALIVE_KNN_FORMAT_BF16;      // the caller sets the format of ITS operands
  if (variant == 3) {
    ALIVE_REQUIRE(skinny_ok, "alive_knn_plan: the skinny kernel needs one item, t <= %d and d <= %d (got t=%d d=%d items=%d)",
                  kSkinnyQ, 2048, t, d, items);
    plan->t = t;
    plan->n = n;
    plan->d = d;
    plan->items = 1;
    plan->kernel = 1;
    plan->ctas_per_unit = 1;
    plan->m_units = 1;
    plan->n_tiles = static_cast<int32_t>((n + kBlockM - 1) / kBlockM);
    plan->grid = plan->n_tiles < num_sms ? plan->n_tiles : num_sms;
    plan->segments = plan->grid;
    plan->tiles_per_segment = (plan->n_tiles + plan->grid - 1) / plan->grid;
    plan->lists = plan->grid;
    return 0;
  }
  const int ctas = variant;
  int slots = num_sms / ctas;   // units that run concurrently
  {
    // The query groups of one library segment should run in the SAME wave (they share the segment's
    // tiles through L2): with a few groups per item, give up the slots that would make segments straddle
    // waves (at most 1/16 of the machine).  ALIVE_KNN_GRID_ALIGN=0/1 (A/B runs).
    static const int align = getenv("ALIVE_KNN_GRID_ALIGN") ? atoi(getenv("ALIVE_KNN_GRID_ALIGN")) : 0;
    const int mu = (t + kBlockM * ctas - 1) / (kBlockM * ctas);
    if (align && items > 1 && mu > 1 && mu <= slots && (slots % mu) * 16 <= slots) slots -= slots % mu;
  }
  plan->t = t;
  plan->n = n;
  plan->d = d;
  plan->items = items;
  plan->ctas_per_unit = ctas;
  plan->m_units = (t + kBlockM * ctas - 1) / (kBlockM * ctas);
  plan->n_tiles = static_cast<int32_t>((n + kBlockN - 1) / kBlockN);
  // Segments (per item): at least 16 when the library is long (more, shorter lists make the
  // completeness certificate easy), enough units to fill every SM, and as few idle tile-slots in
  // the last wave as possible.
  const int n_tiles = plan->n_tiles;
  const long long mi = static_cast<long long>(plan->m_units) * items;    // query groups over all items
  // candidates: from 16 segments (enough lists for the certificate) up to 4x what fills the
  // machine; the cost model below picks the cheapest, ties go to FEWER segments (every unit start
  // pays for a cold top list in the epilogue)
  const int s_fill = static_cast<int>((slots + mi - 1) / mi);
  int s_lo = 16;
  if (s_lo > n_tiles) s_lo = n_tiles;
  int s_hi = 4 * (s_fill > 16 ? s_fill : 16);
  if (s_hi > n_tiles) s_hi = n_tiles;
  long long best_cost = -1;
  int best_tps = 1;
  for (int s = s_lo; s <= s_hi; ++s) {
    const int tps = (n_tiles + s - 1) / s;
    const int s_eff = (n_tiles + tps - 1) / tps;
    const long long units = mi * s_eff;
    const long long waves = (units + slots - 1) / slots;
    const long long cost = waves * tps;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_tps = tps;
    }
  }
  plan->tiles_per_segment = best_tps;
  plan->segments = (n_tiles + best_tps - 1) / best_tps;
  plan->lists = plan->segments * 2;
  const long long units = mi * plan->segments;
  plan->grid = static_cast<int32_t>((units < slots ? units : slots) * ctas);
  return 0;
}

extern "C" int alive_knn_plan(int32_t t, int64_t n, int32_t d, int32_t num_sms, int32_t variant,
                              alive_knn_plan_t* plan) {
  return alive_knn_plan_batched(1, t, n, d, num_sms, variant, plan);
}

namespace alive {
// second screen pass of the one-call pipeline: see SearchParams (collect mode) and select.cu
int collect_impl(const uint16_t* qc_packed, const uint16_t* lib_packed, const alive_knn_plan_t* plan,
                 const int32_t* active_rows, const float* cut, int32_t* cnt, int32_t* idx, int32_t cap,
                 const uint16_t* qc_lo, const uint16_t* lib_lo, alive_stream_t stream) {
  ALIVE_REQUIRE(plan && plan->kernel == 0, "collect pass: needs a tiled plan");
  ALIVE_REQUIRE((qc_lo == nullptr) == (lib_lo == nullptr), "collect pass: the second planes of queries and library go together");
  ALIVE_REQUIRE(lib_lo == nullptr || ((reinterpret_cast<uintptr_t>(qc_lo) | reinterpret_cast<uintptr_t>(lib_lo)) & 15) == 0,
                "collect pass: second planes must be 16-byte aligned");
  CollectArgs ca{active_rows, cut, cnt, idx, cap, qc_lo, lib_lo};
  const bool refine = lib_lo != nullptr;
  if (plan->ctas_per_unit == 1) {
    if (refine) return launch_search<1, true, true>(qc_packed, lib_packed, *plan, nullptr, nullptr, 1, as_stream(stream), &ca);
    return launch_search<1, true, false>(qc_packed, lib_packed, *plan, nullptr, nullptr, 1, as_stream(stream), &ca);
  }
  if (refine) return launch_search<2, true, true>(qc_packed, lib_packed, *plan, nullptr, nullptr, 1, as_stream(stream), &ca);
  return launch_search<2, true, false>(qc_packed, lib_packed, *plan, nullptr, nullptr, 1, as_stream(stream), &ca);
}

int search_impl(const uint16_t* q_packed, const uint16_t* lib_packed, const alive_knn_plan_t* plan, float* cand_score,
                int32_t* cand_idx, int after_query_pack, alive_stream_t stream) {
  ALIVE_REQUIRE(plan && q_packed && lib_packed && cand_score && cand_idx, "alive_knn_search: NULL argument");
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(q_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(lib_packed) & 15) == 0,
                "alive_knn_search: packed operands must be 16-byte aligned");
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(cand_score) & 15) == 0 && (reinterpret_cast<uintptr_t>(cand_idx) & 15) == 0,
                "alive_knn_search: candidate buffers must be 16-byte aligned");
  ALIVE_REQUIRE(plan->d % 64 == 0 && plan->grid > 0 && plan->grid % plan->ctas_per_unit == 0, "alive_knn_search: bad plan");
  ALIVE_REQUIRE(plan->format == ALIVE_KNN_FORMAT_BF16 || plan->format == ALIVE_KNN_FORMAT_FP16, "alive_knn_search: bad plan format");
  if (plan->kernel == 1) return launch_skinny(q_packed, lib_packed, *plan, cand_score, cand_idx, after_query_pack, as_stream(stream));
  ALIVE_REQUIRE(plan->kernel == 0, "alive_knn_search: unknown plan kernel %d", plan->kernel);
  if (plan->ctas_per_unit == 1) return launch_search<1>(q_packed, lib_packed, *plan, cand_score, cand_idx, after_query_pack, as_stream(stream));
  if (plan->ctas_per_unit == 2) return launch_search<2>(q_packed, lib_packed, *plan, cand_score, cand_idx, after_query_pack, as_stream(stream));
  set_error("alive_knn_search: ctas_per_unit must be 1 or 2");
  return -1;
}
}  // namespace alive

extern "C" int alive_knn_search(const uint16_t* q_packed, const uint16_t* lib_packed,
                                const alive_knn_plan_t* plan, float* cand_score, int32_t* cand_idx,
                                alive_stream_t stream) {
  return alive::search_impl(q_packed, lib_packed, plan, cand_score, cand_idx, 0, stream);
}
