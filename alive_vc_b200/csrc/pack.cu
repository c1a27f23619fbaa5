// K1 - L2-normalise and pack frames (library once, queries per call).
//
// Reference lines replaced: module/common.py:100-104 (transpose views, torch.norm,
// `x / norm`) - the reference redoes this for the WHOLE library on every call
// (SURVEY §0); here it runs once per library and once per query batch.
//
// Input  x[i*stride_n + j*stride_d]  (reference layout [1,D,N]: stride_n=1, stride_d=N)
// Output raw    [n,d] f32  row-major copy of the raw frames (gather + exact rescoring)
//        norms  [n]   f32  sqrt(sum x^2), sum in fp64
//        packed [n,d] bf16 row-major x/|x| (IEEE division, then round-to-nearest-even)
//        err    [n]   f32  || bf16(x/|x|) - x/|x| ||_2   (screening error bound input)
//        stats  [0] max of the err bits, [1] count of non-finite normalised rows
//
// HBM-bound: algorithmic bytes per frame = d*(4 read + 4 raw + 2 packed) + 8 = 7,688 B at d=768.
// Kernels (dispatch in pack_impl at the end of this file; DESIGN.md §4 K1 has the measurements):
//   pack_cm2_kernel    channel-major libraries: 32-frame [d][36] tile filled by 16-byte cp.async, 16 warps x 2 frames
//   pack_cm_kernel     the same tile finished by 8 warps x 4 frames (A/B: ALIVE_KNN_PACK_CM2=0)
//   pack_rm_kernel     row-major frames: one warp per frame, the row stays in registers
//   pack_kernel<8/32>  any strides / alignment, batches of query items (FrameMap)
//   pack_frame_kernel  <= 512 frames (a streaming chunk): one CTA per frame
// All of them compute the same bits: fp64 norm, correctly rounded x/|x|, RN-even bf16.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace alive {
namespace {

constexpr int kPackThreads = 256;  // 8 warps

// stats[0] = max of the per-frame error norms, stats[1] = number of non-finite rows.  One atomic per
// FRAME on a single address serialises the whole kernel (measured: ~2.7 ns per atomic = 2.7 ms per
// million frames, more than the HBM time of the pack itself), so frames are reduced per CTA in shared
// memory and the CTA touches the global word only when it would raise it (a stale read only costs a
// redundant atomic).
struct CtaStats {
  unsigned int max_bits;
  unsigned int bad;
  unsigned int max2_bits;
};
__device__ __forceinline__ void cta_stats_init(CtaStats* cs) {
  if (threadIdx.x == 0) {
    cs->max_bits = 0u;
    cs->bad = 0u;
    cs->max2_bits = 0u;
  }
}
__device__ __forceinline__ void cta_stats_add(CtaStats* cs, float e, bool finite, float e2nd) {
  if (finite) {
    atomicMax(&cs->max_bits, __float_as_uint(e));
    atomicMax(&cs->max2_bits, __float_as_uint(e2nd));
  } else {
    atomicAdd(&cs->bad, 1u);
  }
}
// after a __syncthreads() that follows every cta_stats_add of the CTA
__device__ __forceinline__ void cta_stats_publish(const CtaStats* cs, unsigned int* __restrict__ stats) {
  if (threadIdx.x == 0 && stats != nullptr) {
    if (cs->max_bits > __ldcg(&stats[0])) atomicMax(&stats[0], cs->max_bits);
    if (cs->bad) atomicAdd(&stats[1], cs->bad);
    if (cs->max2_bits > __ldcg(&stats[2])) atomicMax(&stats[2], cs->max2_bits);
  }
}

// Optional second bf16 plane (the split-bf16 refinement of the screen, search_sm100.cu collect mode):
//   lo   [n,d] bf16 = bf16_rn(a - hi)   where a = x/|x| (float32) and hi = bf16_rn(a) is the `packed` plane
//   err2 [n]   f32  = || a - hi - lo ||_2 (rounded up); stats[2] = max of its bits over the finite rows
// Both subtractions are exact in float32 (a value minus its own rounding).
struct Refine {
  uint16_t* lo;
  float* err2;
};

// The 16-bit format of the two planes (alive_knn.h ALIVE_KNN_FORMAT_*): bf16 (what the north-star text names) or
// IEEE fp16.  Normalised frames live in [-1, 1], so fp16's range is no constraint and its 11-bit significand rounds
// 8x finer than bf16's 8 bits at the same tensor-core rate (kind::f16 takes either): the certificate's band shrinks
// 8x.  Everything downstream is format-agnostic - the error norms are measured from the values actually stored.
template <bool kHalf> struct Plane;
template <> struct Plane<false> {
  static __device__ __forceinline__ unsigned short one(float a) { return __bfloat16_as_ushort(__float2bfloat16_rn(a)); }
  static __device__ __forceinline__ float val(unsigned short h) { return __uint_as_float(static_cast<unsigned>(h) << 16); }
  static __device__ __forceinline__ unsigned pack2(float a, float b) {      // low half = a
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&p);
  }
  static __device__ __forceinline__ float lo_val(unsigned w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi_val(unsigned w) { return __uint_as_float(w & 0xffff0000u); }
};
template <> struct Plane<true> {
  static __device__ __forceinline__ unsigned short one(float a) { return __half_as_ushort(__float2half_rn(a)); }
  static __device__ __forceinline__ float val(unsigned short h) { return __half2float(__ushort_as_half(h)); }
  static __device__ __forceinline__ unsigned pack2(float a, float b) {
    const __half2 p = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&p);
  }
  static __device__ __forceinline__ float lo_val(unsigned w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ float hi_val(unsigned w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }
};
// second-plane bits of an element; fh = float(hi); *r2 = a - hi - lo
template <bool kHalf>
__device__ __forceinline__ unsigned short split_lo(float a, float fh, float* r2) {
  const float r1 = a - fh;
  const unsigned short l = Plane<kHalf>::one(r1);
  *r2 = r1 - Plane<kHalf>::val(l);
  return l;
}
__device__ __forceinline__ float round_up_norm2(float sumsq) { return sqrtf(sumsq) * 1.0001f + 1e-12f; }

// Query batches [B, D, T] arrive as B items of `item_frames` frames, `stride_b` elements apart (one
// launch for all items; output rows b*item_frames + t).  A library is one item.
struct FrameMap {
  int item_frames;
  long long stride_b, stride_n;
};
__device__ __forceinline__ const float* frame_ptr(const float* __restrict__ x, const FrameMap& m, long long f) {
  const int fi = static_cast<int>(f);
  const int b = fi / m.item_frames;
  return x + b * m.stride_b + (fi - b * m.item_frames) * m.stride_n;
}

// kFrames = frames per CTA: 32 for libraries (128-byte coalesced reads of the channel-major
// input), 8 for small query batches (more CTAs; 32-byte sectors are still fully used).
template <int kFrames, bool kHalf>
__global__ void __launch_bounds__(kPackThreads)
pack_kernel(const float* __restrict__ x, long long n, int d, const FrameMap fm, long long stride_d,
            float* __restrict__ raw, float* __restrict__ norms, uint16_t* __restrict__ packed,
            float* __restrict__ err, unsigned int* __restrict__ stats, int* __restrict__ zero_words, int n_zero,
            int async_stage, const Refine rf) {
  extern __shared__ float tile[];          // [d][kFrames + 1]
  __shared__ CtaStats cta_stats;
  cta_stats_init(&cta_stats);
  pdl_launch_dependents();                 // a search launched behind this pack may start streaming the library
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kPackThreads) zero_words[i] = 0;
  const int ld = kFrames + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f0 = static_cast<long long>(blockIdx.x) * kFrames;
  const int nf = static_cast<int>(min(static_cast<long long>(kFrames), n - f0));

  if (fm.stride_n == 1 || stride_d != 1) {
    // frames are the fast axis (or fully strided): lane -> (frame, channel sub-row)
    constexpr int kRowsPerWarp = 32 / kFrames;
    const int f = lane % kFrames, jsub = lane / kFrames;
    const bool ok = f < nf;
    const float* src = ok ? frame_ptr(x, fm, f0 + f) : x;
    if (async_stage) {
      // every 4-byte element of the CTA's tile is requested before anything waits (cp.async straight
      // into shared memory, no register staging): ~d*kFrames*4 B in flight per CTA instead of a
      // handful of loads per thread - the staging loop was latency-bound, not bandwidth-bound
      const unsigned tile_s = static_cast<unsigned>(__cvta_generic_to_shared(tile));
      for (int j = warp * kRowsPerWarp + jsub; j < d; j += (kPackThreads / 32) * kRowsPerWarp) {
        if (ok)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + 4u * (j * ld + f)),
                       "l"(src + j * stride_d)
                       : "memory");
        else
          tile[j * ld + f] = 0.f;
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      for (int j = warp * kRowsPerWarp + jsub; j < d; j += (kPackThreads / 32) * kRowsPerWarp)
        tile[j * ld + f] = ok ? src[j * stride_d] : 0.f;
    }
  } else {
    // already row-major frames: lane = channel
    const unsigned tile_s = static_cast<unsigned>(__cvta_generic_to_shared(tile));
    for (int f = warp; f < kFrames; f += kPackThreads / 32) {
      const float* src = f < nf ? frame_ptr(x, fm, f0 + f) : x;
      if (async_stage && f < nf) {
        for (int j = lane; j < d; j += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + 4u * (j * ld + f)), "l"(src + j)
                       : "memory");
      } else {
        for (int j = lane; j < d; j += 32) tile[j * ld + f] = (f < nf) ? src[j] : 0.f;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();

  // each warp finishes whole frames: 4 frames per warp
  for (int f = warp; f < nf; f += kPackThreads / 32) {
    const long long row = f0 + f;
    double ss = 0.0;
    for (int j = lane; j < d; j += 32) {
      const float v = tile[j * ld + f];
      raw[row * d + j] = v;
      ss += static_cast<double>(v) * static_cast<double>(v);
    }
    ss = warp_sum_f64(ss);
    const float nrm = static_cast<float>(sqrt(ss));
    double e2 = 0.0, e22 = 0.0;
    bool finite = true;
    // two channels per lane so the bf16 row is written as 4-byte words
    for (int j = 2 * lane; j < d; j += 64) {
      const float a = __fdiv_rn(tile[j * ld + f], nrm);
      const float b = __fdiv_rn(tile[(j + 1) * ld + f], nrm);
      const unsigned short ha = Plane<kHalf>::one(a), hb = Plane<kHalf>::one(b);
      const float fa = Plane<kHalf>::val(ha), fb = Plane<kHalf>::val(hb);
      const float da = fa - a, db = fb - b;
      e2 += static_cast<double>(da) * da + static_cast<double>(db) * db;
      finite = finite && isfinite(a) && isfinite(b);
      *reinterpret_cast<unsigned*>(packed + row * d + j) = static_cast<unsigned>(ha) | (static_cast<unsigned>(hb) << 16);
      float ra, rb;
      const unsigned la = split_lo<kHalf>(a, fa, &ra), lb = split_lo<kHalf>(b, fb, &rb);
      e22 += static_cast<double>(ra) * ra + static_cast<double>(rb) * rb;
      if (rf.lo) *reinterpret_cast<unsigned*>(rf.lo + row * d + j) = la | (lb << 16);
    }
    e2 = warp_sum_f64(e2);
    e22 = warp_sum_f64(e22);
    finite = __all_sync(0xffffffffu, finite);
    if (lane == 0) {
      norms[row] = nrm;
      // round the error norm UP a little: it feeds a bound that must not be under-estimated
      float e = finite ? static_cast<float>(sqrt(e2)) * 1.0001f + 1e-9f : 0.f;
      const float e2nd = finite ? round_up_norm2(static_cast<float>(e22)) : 0.f;
      if (err) err[row] = e;
      if (rf.err2) rf.err2[row] = e2nd;
      cta_stats_add(&cta_stats, e, finite, e2nd);
    }
  }
  __syncthreads();
  cta_stats_publish(&cta_stats, stats);
}

// ---------------------------------------------------------------------------------------------
// Library-sized inputs: two layout-specific kernels.  The generic kernel above spent ~1200 warp
// instructions per frame (ncu: issue-bound at 53 % of HBM peak; 4-byte LDGSTS at 8 cycles each, a
// full IEEE division per element, three fp64 conversions per element); these two do the same
// arithmetic in ~400.
// ---------------------------------------------------------------------------------------------

// x / nrm, correctly rounded, for 2^-40 <= nrm <= 2^40 and r = __frcp_rn(nrm) (correctly rounded
// reciprocal): quotient estimate + ONE exact-remainder correction (Markstein: with a correctly rounded
// reciprocal and a faithful estimate, q + r*(x - q*nrm) rounds to the correctly rounded quotient;
// checked in exact rational arithmetic on adversarial mantissas and bit for bit against __fdiv_rn by
// test_fast_pack_kernels_equal_the_generic_kernel); a quotient too small for the remainder to be exact
// (or a signed zero) takes the full IEEE division.
__device__ __forceinline__ float div_by_norm(float x, float nrm, float r) {
  float q = __fmul_rn(x, r);
  float e = __fmaf_rn(-q, nrm, x);
  q = __fmaf_rn(e, r, q);
#ifdef ALIVE_PACK_TWO_CORRECTIONS    // not needed (see above); costs 2-3 % of the pack rate
  e = __fmaf_rn(-q, nrm, x);
  q = __fmaf_rn(e, r, q);
#endif
  return (fabsf(q) >= 0x1p-40f) ? q : __fdiv_rn(x, nrm);
}
__device__ __forceinline__ bool norm_is_tame(float nrm) { return nrm >= 0x1p-40f && nrm <= 0x1p40f; }

__device__ __forceinline__ void finish_frame(long long row, float nrm, float e2, float e22, bool finite,
                                             float* __restrict__ norms, float* __restrict__ err, const Refine& rf,
                                             CtaStats* cta_stats) {
  norms[row] = nrm;
  // round the error norm UP a little: it feeds a bound that must not be under-estimated (the fp32 sum of
  // d squares is within 2e-6 relative of the exact one)
  const float e = finite ? sqrtf(e2) * 1.0001f + 1e-9f : 0.f;
  const float e2nd = finite ? round_up_norm2(e22) : 0.f;
  if (err) err[row] = e;
  if (rf.err2) rf.err2[row] = e2nd;
  cta_stats_add(cta_stats, e, finite, e2nd);
}

// Channel-major input (the reference's [D, N], stride_n == 1, 16-byte aligned rows): a tile = 32
// frames; 16-byte cp.async copies (4 frames of one channel) fill a [d][36] tile, every warp then owns
// 4 frames and reads them back as float4 (conflict-free: quarter-warp rows are 144 B apart).
// Default: one CTA per tile, two CTAs per SM (one stages while the other finishes).  kDouble (opt-in
// experiment, slower): persistent CTAs, one per SM, with TWO tiles in shared memory (d <= 768: 2 x 108 KB).
constexpr int kCmLd = 36;

__device__ __forceinline__ void cm_stage(float* tile, const float* __restrict__ x, long long f0, int nf, int d,
                                         long long stride_d) {
  const unsigned tile_s = static_cast<unsigned>(__cvta_generic_to_shared(tile));
  if (nf == 32) {
    const int g4 = threadIdx.x & 7, jsub = threadIdx.x >> 3;          // 8 groups of 4 frames x blockDim/8 channel rows
    const int rows = static_cast<int>(blockDim.x >> 3);               // (32 with 256 threads, 64 with 512)
    const float* src = x + f0 + 4 * g4 + static_cast<long long>(jsub) * stride_d;
    unsigned dst = tile_s + 4u * (jsub * kCmLd + 4 * g4);
    const long long src_step = rows * stride_d;
    const unsigned dst_step = 4u * rows * kCmLd;
#pragma unroll 4
    for (int j = jsub; j < d; j += rows) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      src += src_step;
      dst += dst_step;
    }
  } else {
    // ragged last tile: element-wise, missing frames read as zeros
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool ok = lane < nf;
    const float* src = x + f0 + lane;
    for (int j = warp; j < d; j += static_cast<int>(blockDim.x >> 5)) {
      if (ok)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + 4u * (j * kCmLd + lane)),
                     "l"(src + j * stride_d)
                     : "memory");
      else
        tile[j * kCmLd + lane] = 0.f;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// warp w: frames 4w..4w+3 of the tile, lane: channels lane, lane+32, ...
template <bool kHalf>
__device__ __forceinline__ void cm_compute(const float* tile, long long f0, int nf, int d, float* __restrict__ raw,
                                           float* __restrict__ norms, uint16_t* __restrict__ packed,
                                           float* __restrict__ err, const Refine& rf, CtaStats* cta_stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fw = 4 * warp;
  const long long row0 = f0 + fw;
  const int nv = max(0, min(4, nf - fw));
  if (nv == 0) return;
  const float4* t4 = reinterpret_cast<const float4*>(tile) + warp;     // tile[j*36 + 4w] = t4[j*9]
  double ss0 = 0.0, ss1 = 0.0, ss2 = 0.0, ss3 = 0.0;
  if (nv == 4) {                       // (every group but the last of a ragged tile: no per-store predicates)
    float* r0 = raw + row0 * d;
#pragma unroll 4
    for (int j = lane; j < d; j += 32) {
      const float4 v = t4[j * (kCmLd / 4)];
      r0[j] = v.x;
      r0[j + d] = v.y;
      r0[j + 2 * d] = v.z;
      r0[j + 3 * d] = v.w;
      ss0 += static_cast<double>(v.x) * static_cast<double>(v.x);
      ss1 += static_cast<double>(v.y) * static_cast<double>(v.y);
      ss2 += static_cast<double>(v.z) * static_cast<double>(v.z);
      ss3 += static_cast<double>(v.w) * static_cast<double>(v.w);
    }
  } else {
    for (int j = lane; j < d; j += 32) {
      const float4 v = t4[j * (kCmLd / 4)];
      raw[row0 * d + j] = v.x;
      if (nv > 1) raw[(row0 + 1) * d + j] = v.y;
      if (nv > 2) raw[(row0 + 2) * d + j] = v.z;
      ss0 += static_cast<double>(v.x) * static_cast<double>(v.x);
      ss1 += static_cast<double>(v.y) * static_cast<double>(v.y);
      ss2 += static_cast<double>(v.z) * static_cast<double>(v.z);
      ss3 += static_cast<double>(v.w) * static_cast<double>(v.w);
    }
  }
  float nrm[4];
  nrm[0] = static_cast<float>(sqrt(warp_sum_f64(ss0)));
  nrm[1] = static_cast<float>(sqrt(warp_sum_f64(ss1)));
  nrm[2] = static_cast<float>(sqrt(warp_sum_f64(ss2)));
  nrm[3] = static_cast<float>(sqrt(warp_sum_f64(ss3)));
  bool tame[4], finite[4];
  float rinv[4], e2[4], e22[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tame[c] = norm_is_tame(nrm[c]);
    rinv[c] = tame[c] ? __frcp_rn(nrm[c]) : 0.f;
    finite[c] = true;
    e2[c] = 0.f;
    e22[c] = 0.f;
  }
  if (tame[0] && tame[1] && tame[2] && tame[3]) {
    // Common case, branch-free per element (the per-element `tame` / small-quotient branches of the first version
    // cost ~50 instructions per element: the kernel was issue-bound at half the HBM rate).  Lane pairs
    // (2m, 2m+1) hold adjacent channels: they swap half of their four bf16 values so that the even lane
    // writes frames 0,1 and the odd lane frames 2,3 as 4-byte words (two channels each) - two 32-bit stores per
    // lane and iteration instead of four 16-bit ones.
    const bool odd = (lane & 1) != 0;
    const int my_f = odd ? 2 : 0;                              // first of the two frames this lane stores
    unsigned* pkA = reinterpret_cast<unsigned*>(packed + (row0 + my_f) * d);
    unsigned* pkB = reinterpret_cast<unsigned*>(packed + (row0 + my_f + 1) * d);
    const bool okA = my_f < nv, okB = my_f + 1 < nv;
    const bool want_lo = rf.lo != nullptr;
    unsigned* loA = want_lo ? reinterpret_cast<unsigned*>(rf.lo + (row0 + my_f) * d) : nullptr;
    unsigned* loB = want_lo ? reinterpret_cast<unsigned*>(rf.lo + (row0 + my_f + 1) * d) : nullptr;
    // one 32-channel step; kTail: the last, partial step of a d that is not a multiple of 32 (d is even: the two
    // lanes of a pair are live together; dead lanes still take part in the exchange)
    auto step = [&](int j, bool live) {
      const float4 v = live ? t4[j * (kCmLd / 4)] : make_float4(1.f, 1.f, 1.f, 1.f);
      const float xs[4] = {v.x, v.y, v.z, v.w};
      float a[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float q0 = __fmul_rn(xs[c], rinv[c]);
        const float e = __fmaf_rn(-q0, nrm[c], xs[c]);
        a[c] = __fmaf_rn(e, rinv[c], q0);
      }
      // zeros (their sign must survive) and quotients too small for an exact remainder: rare, one test per step
      if (!(fminf(fminf(fabsf(a[0]), fabsf(a[1])), fminf(fabsf(a[2]), fabsf(a[3]))) >= 0x1p-40f)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (xs[c] == 0.f) a[c] = __fmul_rn(xs[c], rinv[c]);
          else if (!(fabsf(a[c]) >= 0x1p-40f)) a[c] = __fdiv_rn(xs[c], nrm[c]);
        }
      }
      // two packing conversions: (frame 0, frame 1) and (frame 2, frame 3) of MY channel, low half = first frame
      const unsigned w01 = Plane<kHalf>::pack2(a[0], a[1]), w23 = Plane<kHalf>::pack2(a[2], a[3]);
      // a - hi per frame (exact): what the second plane has to carry
      const float r0 = a[0] - Plane<kHalf>::lo_val(w01), r1 = a[1] - Plane<kHalf>::hi_val(w01);
      const float r2 = a[2] - Plane<kHalf>::lo_val(w23), r3 = a[3] - Plane<kHalf>::hi_val(w23);
      if (live) {
        e2[0] = fmaf(r0, r0, e2[0]);
        e2[1] = fmaf(r1, r1, e2[1]);
        e2[2] = fmaf(r2, r2, e2[2]);
        e2[3] = fmaf(r3, r3, e2[3]);
      }
      const unsigned keep = odd ? w23 : w01;                                       // my channel, the frames I store
      const unsigned got = __shfl_xor_sync(0xffffffffu, odd ? w01 : w23, 1);       // partner's channel, my frames
      const unsigned ev = odd ? got : keep, od = odd ? keep : got;                 // even / odd channel of the pair
      const int w = j >> 1;                                                         // word index of the channel pair
      if (okA && live) pkA[w] = __byte_perm(ev, od, 0x5410);                        // frame A: (even ch, odd ch)
      if (okB && live) pkB[w] = __byte_perm(ev, od, 0x7632);                        // frame B
      {
        // second plane: lo = bf16_rn(a - hi), residual a - hi - lo feeds err2 (always) and the plane is stored on request
        const unsigned l01 = Plane<kHalf>::pack2(r0, r1), l23 = Plane<kHalf>::pack2(r2, r3);
        if (live) {
          const float s0 = r0 - Plane<kHalf>::lo_val(l01), s1 = r1 - Plane<kHalf>::hi_val(l01);
          const float s2 = r2 - Plane<kHalf>::lo_val(l23), s3 = r3 - Plane<kHalf>::hi_val(l23);
          e22[0] = fmaf(s0, s0, e22[0]);
          e22[1] = fmaf(s1, s1, e22[1]);
          e22[2] = fmaf(s2, s2, e22[2]);
          e22[3] = fmaf(s3, s3, e22[3]);
        }
        if (want_lo) {                                                               // (warp-uniform)
          const unsigned lkeep = odd ? l23 : l01;
          const unsigned lgot = __shfl_xor_sync(0xffffffffu, odd ? l01 : l23, 1);
          const unsigned lev = odd ? lgot : lkeep, lod = odd ? lkeep : lgot;
          if (okA && live) loA[w] = __byte_perm(lev, lod, 0x5410);
          if (okB && live) loB[w] = __byte_perm(lev, lod, 0x7632);
        }
      }
    };
    const int d_full = d & ~31;
#pragma unroll 2
    for (int j0 = 0; j0 < d_full; j0 += 32) step(j0 + lane, true);
    if (d_full < d) step(d_full + lane, d_full + lane < d);
  } else {
    // a zero / huge / tiny / non-finite frame in this group of four: full IEEE division, element by element
    unsigned short* pk16 = reinterpret_cast<unsigned short*>(packed);
    for (int j = lane; j < d; j += 32) {
      const float4 v = t4[j * (kCmLd / 4)];
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a;
        if (tame[c]) {
          a = div_by_norm(xs[c], nrm[c], rinv[c]);
        } else {
          a = __fdiv_rn(xs[c], nrm[c]);
          finite[c] = finite[c] && isfinite(a);
        }
        const unsigned short hb = Plane<kHalf>::one(a);
        const float fh = Plane<kHalf>::val(hb);
        const float da = fh - a;
        e2[c] = fmaf(da, da, e2[c]);
        if (c < nv) pk16[(row0 + c) * d + j] = hb;
        float r2nd;
        const unsigned short lb = split_lo<kHalf>(a, fh, &r2nd);
        e22[c] = fmaf(r2nd, r2nd, e22[c]);
        if (rf.lo && c < nv) rf.lo[(row0 + c) * d + j] = lb;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e2[c] += __shfl_xor_sync(0xffffffffu, e2[c], o);
      e22[c] += __shfl_xor_sync(0xffffffffu, e22[c], o);
    }
    finite[c] = __all_sync(0xffffffffu, finite[c]);
    if (lane == 0 && c < nv) finish_frame(row0 + c, nrm[c], e2[c], e22[c], finite[c], norms, err, rf, cta_stats);
  }
}

template <bool kDouble, bool kHalf>
__global__ void __launch_bounds__(kPackThreads)
pack_cm_kernel(const float* __restrict__ x, long long n, int d, long long stride_d, float* __restrict__ raw,
               float* __restrict__ norms, uint16_t* __restrict__ packed, float* __restrict__ err,
               unsigned int* __restrict__ stats, int* __restrict__ zero_words, int n_zero, const Refine rf) {
  extern __shared__ __align__(16) float tile[];   // [kDouble ? 2 : 1][d][36]
  __shared__ CtaStats cta_stats;
  cta_stats_init(&cta_stats);
  pdl_launch_dependents();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kPackThreads) zero_words[i] = 0;
  const long long n_tiles = (n + 31) / 32;
  const int tile_floats = d * kCmLd;
  long long t = blockIdx.x;
  if (t < n_tiles) cm_stage(tile, x, t * 32, static_cast<int>(min(32ll, n - t * 32)), d, stride_d);
  for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
    const float* cur = tile + (kDouble ? (it & 1) * tile_floats : 0);
    const long long tn = t + gridDim.x;
    if (kDouble && tn < n_tiles) {
      cm_stage(tile + ((it + 1) & 1) * tile_floats, x, tn * 32, static_cast<int>(min(32ll, n - tn * 32)), d, stride_d);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    cm_compute<kHalf>(cur, t * 32, static_cast<int>(min(32ll, n - t * 32)), d, raw, norms, packed, err, rf, &cta_stats);
    __syncthreads();                 // every warp is done with `cur` before it is refilled
    if (!kDouble && tn < n_tiles)
      cm_stage(tile, x, tn * 32, static_cast<int>(min(32ll, n - tn * 32)), d, stride_d);
  }
  cta_stats_publish(&cta_stats, stats);
}

// The same tile finished by SIXTEEN warps, two frames each (512-thread CTAs, two per SM, <= 64 registers): twice the
// warps of cm_compute on the same shared-memory footprint.  The kernel above sits at 0.82 of the HBM peak with 24 % of
// the warp slots in use (105 registers, two 8-warp CTAs per SM: its finishing pass is latency-bound) - the row-major
// kernel showed what occupancy is worth in this family.  Identical arithmetic per element; a lane pair (2m, 2m+1)
// exchanges its packed word so that the even lane stores frame 0 and the odd lane frame 1 as 32-bit words of two
// adjacent channels.
template <bool kHalf>
__device__ __forceinline__ void cm_compute2(const float* tile, long long f0, int nf, int d, float* __restrict__ raw,
                                            float* __restrict__ norms, uint16_t* __restrict__ packed,
                                            float* __restrict__ err, const Refine& rf, CtaStats* cta_stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fw = 2 * warp;
  const long long row0 = f0 + fw;
  const int nv = max(0, min(2, nf - fw));
  if (nv == 0) return;
  const float2* t2 = reinterpret_cast<const float2*>(tile) + warp;      // tile[j*36 + 2w] = t2[j*18]
  double ss0 = 0.0, ss1 = 0.0;
  {
    float* r0 = raw + row0 * d;
    const bool two = nv > 1;
#pragma unroll 4
    for (int j = lane; j < d; j += 32) {
      const float2 v = t2[j * (kCmLd / 2)];
      r0[j] = v.x;
      if (two) r0[j + d] = v.y;
      ss0 += static_cast<double>(v.x) * static_cast<double>(v.x);
      ss1 += static_cast<double>(v.y) * static_cast<double>(v.y);
    }
  }
  float nrm[2];
  nrm[0] = static_cast<float>(sqrt(warp_sum_f64(ss0)));
  nrm[1] = static_cast<float>(sqrt(warp_sum_f64(ss1)));
  bool tame[2], finite[2];
  float rinv[2], e2[2], e22[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    tame[c] = norm_is_tame(nrm[c]);
    rinv[c] = tame[c] ? __frcp_rn(nrm[c]) : 0.f;
    finite[c] = true;
    e2[c] = 0.f;
    e22[c] = 0.f;
  }
  if (tame[0] && tame[1]) {
    const bool odd = (lane & 1) != 0;
    const bool ok = (odd ? 1 : 0) < nv;                        // the frame this lane stores exists
    unsigned* pk = reinterpret_cast<unsigned*>(packed + (row0 + (odd ? 1 : 0)) * d);
    const bool want_lo = rf.lo != nullptr;
    unsigned* lo = want_lo ? reinterpret_cast<unsigned*>(rf.lo + (row0 + (odd ? 1 : 0)) * d) : nullptr;
    auto step = [&](int j, bool live) {
      const float2 v = live ? t2[j * (kCmLd / 2)] : make_float2(1.f, 1.f);
      const float xs[2] = {v.x, v.y};
      float a[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float q0 = __fmul_rn(xs[c], rinv[c]);
        const float e = __fmaf_rn(-q0, nrm[c], xs[c]);
        a[c] = __fmaf_rn(e, rinv[c], q0);
      }
      if (!(fminf(fabsf(a[0]), fabsf(a[1])) >= 0x1p-40f)) {    // zeros / tiny quotients: rare, see cm_compute
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (xs[c] == 0.f) a[c] = __fmul_rn(xs[c], rinv[c]);
          else if (!(fabsf(a[c]) >= 0x1p-40f)) a[c] = __fdiv_rn(xs[c], nrm[c]);
        }
      }
      const unsigned w01 = Plane<kHalf>::pack2(a[0], a[1]);      // my channel: frame 0 low half, frame 1 high half
      const float r0 = a[0] - Plane<kHalf>::lo_val(w01), r1 = a[1] - Plane<kHalf>::hi_val(w01);
      if (live) {
        e2[0] = fmaf(r0, r0, e2[0]);
        e2[1] = fmaf(r1, r1, e2[1]);
      }
      const unsigned got = __shfl_xor_sync(0xffffffffu, w01, 1);   // the partner's channel, both frames
      const int w = j >> 1;                                        // word index of the channel pair
      // even lane: frame 0 of (my channel, partner's); odd lane: frame 1 of (partner's channel, mine)
      if (ok && live) pk[w] = odd ? __byte_perm(got, w01, 0x7632) : __byte_perm(w01, got, 0x5410);
      const unsigned l01 = Plane<kHalf>::pack2(r0, r1);
      if (live) {
        const float s0 = r0 - Plane<kHalf>::lo_val(l01), s1 = r1 - Plane<kHalf>::hi_val(l01);
        e22[0] = fmaf(s0, s0, e22[0]);
        e22[1] = fmaf(s1, s1, e22[1]);
      }
      if (want_lo) {                                               // (warp-uniform)
        const unsigned lgot = __shfl_xor_sync(0xffffffffu, l01, 1);
        if (ok && live) lo[w] = odd ? __byte_perm(lgot, l01, 0x7632) : __byte_perm(l01, lgot, 0x5410);
      }
    };
    const int d_full = d & ~31;
#pragma unroll 2
    for (int j0 = 0; j0 < d_full; j0 += 32) step(j0 + lane, true);
    if (d_full < d) step(d_full + lane, d_full + lane < d);
  } else {
    unsigned short* pk16 = reinterpret_cast<unsigned short*>(packed);
    for (int j = lane; j < d; j += 32) {
      const float2 v = t2[j * (kCmLd / 2)];
      const float xs[2] = {v.x, v.y};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float a;
        if (tame[c]) {
          a = div_by_norm(xs[c], nrm[c], rinv[c]);
        } else {
          a = __fdiv_rn(xs[c], nrm[c]);
          finite[c] = finite[c] && isfinite(a);
        }
        const unsigned short hb = Plane<kHalf>::one(a);
        const float fh = Plane<kHalf>::val(hb);
        const float da = fh - a;
        e2[c] = fmaf(da, da, e2[c]);
        if (c < nv) pk16[(row0 + c) * d + j] = hb;
        float r2nd;
        const unsigned short lb = split_lo<kHalf>(a, fh, &r2nd);
        e22[c] = fmaf(r2nd, r2nd, e22[c]);
        if (rf.lo && c < nv) rf.lo[(row0 + c) * d + j] = lb;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e2[c] += __shfl_xor_sync(0xffffffffu, e2[c], o);
      e22[c] += __shfl_xor_sync(0xffffffffu, e22[c], o);
    }
    finite[c] = __all_sync(0xffffffffu, finite[c]);
    if (lane == 0 && c < nv) finish_frame(row0 + c, nrm[c], e2[c], e22[c], finite[c], norms, err, rf, cta_stats);
  }
}

constexpr int kCm2Threads = 512;
template <bool kHalf>
__global__ void __launch_bounds__(kCm2Threads, 2)
pack_cm2_kernel(const float* __restrict__ x, long long n, int d, long long stride_d, float* __restrict__ raw,
                float* __restrict__ norms, uint16_t* __restrict__ packed, float* __restrict__ err,
                unsigned int* __restrict__ stats, int* __restrict__ zero_words, int n_zero, const Refine rf) {
  extern __shared__ __align__(16) float tile[];   // [d][36]
  __shared__ CtaStats cta_stats;
  cta_stats_init(&cta_stats);
  pdl_launch_dependents();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kCm2Threads) zero_words[i] = 0;
  const long long n_tiles = (n + 31) / 32;
  long long t = blockIdx.x;
  if (t < n_tiles) cm_stage(tile, x, t * 32, static_cast<int>(min(32ll, n - t * 32)), d, stride_d);
  for (; t < n_tiles; t += gridDim.x) {
    const long long tn = t + gridDim.x;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    cm_compute2<kHalf>(tile, t * 32, static_cast<int>(min(32ll, n - t * 32)), d, raw, norms, packed, err, rf, &cta_stats);
    __syncthreads();                 // every warp is done with the tile before it is refilled
    if (tn < n_tiles) cm_stage(tile, x, tn * 32, static_cast<int>(min(32ll, n - tn * 32)), d, stride_d);
  }
  cta_stats_publish(&cta_stats, stats);
}

// Row-major input (a producer's [n, D]: stride_d == 1, 16-byte aligned rows, d % 4 == 0): nothing to
// transpose - one warp per frame, the 3 KB row lives in registers between the norm and the division.
constexpr int kRmMaxV = 12;                         // float4 per lane: d <= 1536 (kV = 6: d <= 768, half the registers)
template <bool kHalf, int kV>
__global__ void __launch_bounds__(kPackThreads, kV <= 6 ? 4 : 3)
pack_rm_kernel(const float* __restrict__ x, long long n, int d, const FrameMap fm, float* __restrict__ raw,
               float* __restrict__ norms, uint16_t* __restrict__ packed, float* __restrict__ err,
               unsigned int* __restrict__ stats, int* __restrict__ zero_words, int n_zero, const Refine rf) {
  pdl_launch_dependents();
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kPackThreads) zero_words[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ CtaStats cta_stats;
  cta_stats_init(&cta_stats);
  __syncthreads();
  const long long row = static_cast<long long>(blockIdx.x) * (kPackThreads / 32) + warp;
  if (row < n) {
  const int d4 = d >> 2;
  const float4* src = reinterpret_cast<const float4*>(frame_ptr(x, fm, row));
  float4* dst_raw = reinterpret_cast<float4*>(raw + row * d);
  float4 v[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int q = lane + 32 * i;
    v[i] = (q < d4) ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  double ss = 0.0;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int q = lane + 32 * i;
    if (q < d4) {
      dst_raw[q] = v[i];
      ss += static_cast<double>(v[i].x) * static_cast<double>(v[i].x);
      ss += static_cast<double>(v[i].y) * static_cast<double>(v[i].y);
      ss += static_cast<double>(v[i].z) * static_cast<double>(v[i].z);
      ss += static_cast<double>(v[i].w) * static_cast<double>(v[i].w);
    }
  }
  const float nrm = static_cast<float>(sqrt(warp_sum_f64(ss)));
  const bool tame = norm_is_tame(nrm);
  const float rinv = tame ? __frcp_rn(nrm) : 0.f;
  bool finite = true;
  float e2 = 0.f, e22 = 0.f;
  uint2* dst_pk = reinterpret_cast<uint2*>(packed + row * d);
  uint2* dst_lo = rf.lo ? reinterpret_cast<uint2*>(rf.lo + row * d) : nullptr;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int q = lane + 32 * i;
    if (q < d4) {
      const float xs[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      unsigned short hb[4], lb[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a;
        if (tame) {
          a = div_by_norm(xs[c], nrm, rinv);
        } else {
          a = __fdiv_rn(xs[c], nrm);
          finite = finite && isfinite(a);
        }
        hb[c] = Plane<kHalf>::one(a);
        const float fh = Plane<kHalf>::val(hb[c]);
        const float da = fh - a;
        e2 = fmaf(da, da, e2);
        float r2nd;
        lb[c] = split_lo<kHalf>(a, fh, &r2nd);
        e22 = fmaf(r2nd, r2nd, e22);
      }
      dst_pk[q] = make_uint2(static_cast<unsigned>(hb[0]) | (static_cast<unsigned>(hb[1]) << 16),
                             static_cast<unsigned>(hb[2]) | (static_cast<unsigned>(hb[3]) << 16));
      if (dst_lo)
        dst_lo[q] = make_uint2(static_cast<unsigned>(lb[0]) | (static_cast<unsigned>(lb[1]) << 16),
                               static_cast<unsigned>(lb[2]) | (static_cast<unsigned>(lb[3]) << 16));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    e22 += __shfl_xor_sync(0xffffffffu, e22, o);
  }
  finite = __all_sync(0xffffffffu, finite);
  if (lane == 0) finish_frame(row, nrm, e2, e22, finite, norms, err, rf, &cta_stats);
  }  // row < n
  __syncthreads();
  cta_stats_publish(&cta_stats, stats);
}

// Tiny batches (streaming chunks, T <= 512): one CTA per frame, three channels per thread, two
// block reductions - one load round trip instead of a 768-row staging loop on a handful of CTAs.
template <bool kHalf>
__global__ void __launch_bounds__(kPackThreads)
pack_frame_kernel(const float* __restrict__ x, long long n, int d, const FrameMap fm, long long stride_d,
                  float* __restrict__ raw, float* __restrict__ norms, uint16_t* __restrict__ packed,
                  float* __restrict__ err, unsigned int* __restrict__ stats, int* __restrict__ zero_words, int n_zero,
                  const Refine rf) {
  __shared__ double red[kPackThreads / 32];
  __shared__ double red2[kPackThreads / 32];
  pdl_launch_dependents();                 // a search launched behind this pack may start streaming the library
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kPackThreads) zero_words[i] = 0;
  __shared__ int fin[kPackThreads / 32];
  const long long row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kPer = 6;                      // d <= 1536
  const float* xrow = frame_ptr(x, fm, row);
  float v[kPer];
  double ss = 0.0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int j = threadIdx.x + i * kPackThreads;
    v[i] = (j < d) ? xrow[j * stride_d] : 0.f;
    if (j < d) raw[row * d + j] = v[i];
    ss += static_cast<double>(v[i]) * static_cast<double>(v[i]);
  }
  ss = warp_sum_f64(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < kPackThreads / 32; ++w) tot += red[w];     // same order in every thread
  const float nrm = static_cast<float>(sqrt(tot));
  __syncthreads();
  double e2 = 0.0, e22 = 0.0;
  bool finite = true;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int j = threadIdx.x + i * kPackThreads;
    if (j < d) {
      const float a = __fdiv_rn(v[i], nrm);
      const unsigned short h = Plane<kHalf>::one(a);
      const float fh = Plane<kHalf>::val(h);
      const float da = fh - a;
      e2 += static_cast<double>(da) * da;
      finite = finite && isfinite(a);
      packed[row * d + j] = h;
      float r2nd;
      const unsigned short lb = split_lo<kHalf>(a, fh, &r2nd);
      e22 += static_cast<double>(r2nd) * r2nd;
      if (rf.lo) rf.lo[row * d + j] = lb;
    }
  }
  e2 = warp_sum_f64(e2);
  e22 = warp_sum_f64(e22);
  finite = __all_sync(0xffffffffu, finite);
  if (lane == 0) {
    red[warp] = e2;
    red2[warp] = e22;
    fin[warp] = finite ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double et = 0.0, et2 = 0.0;
    bool ok = true;
    for (int w = 0; w < kPackThreads / 32; ++w) {
      et += red[w];
      et2 += red2[w];
      ok = ok && fin[w];
    }
    norms[row] = nrm;
    const float e = ok ? static_cast<float>(sqrt(et)) * 1.0001f + 1e-9f : 0.f;
    const float e2nd = ok ? round_up_norm2(static_cast<float>(et2)) : 0.f;
    if (err) err[row] = e;
    if (rf.err2) rf.err2[row] = e2nd;
    if (stats) {
      if (!ok) atomicAdd(&stats[1], 1u);
      else {
        if (__float_as_uint(e) > __ldcg(&stats[0])) atomicMax(&stats[0], __float_as_uint(e));
        if (__float_as_uint(e2nd) > __ldcg(&stats[2])) atomicMax(&stats[2], __float_as_uint(e2nd));
      }
    }
  }
}

}  // namespace
}  // namespace alive

namespace alive {
namespace {
template <bool kHalf>
int pack_dispatch(const float* x, int64_t n, int32_t d, const FrameMap& fm, bool uniform, int64_t stride_n, int64_t stride_d,
                  int64_t stride_b, float* raw, float* norms, uint16_t* pk, float* err, uint32_t* stats, int32_t* zero_words,
                  int32_t n_zero, const Refine& rf, cudaStream_t stream) {
  static PerDeviceOnce attr_once;
  {
    const int rc_attr = attr_once.run([]() -> int {
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(pack_kernel<32, kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 33 * 4)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(pack_kernel<8, kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 9 * 4)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(pack_cm_kernel<false, kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * kCmLd * 4)));
      ALIVE_CHECK_CUDA((cudaFuncSetAttribute(pack_cm_kernel<true, kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 768 * kCmLd * 4)));
      return 0;
    });
    if (rc_attr) return rc_attr;
  }
  // ALIVE_KNN_PACK_ASYNC=0: register-staged loads (the first version; kept for A/B runs)
  static const int async_stage = !(getenv("ALIVE_KNN_PACK_ASYNC") && atoi(getenv("ALIVE_KNN_PACK_ASYNC")) == 0);
  // layout-specific kernels (ALIVE_KNN_PACK_FAST=0: the generic kernel everywhere, for A/B runs)
  static const int fast = !(getenv("ALIVE_KNN_PACK_FAST") && atoi(getenv("ALIVE_KNN_PACK_FAST")) == 0);
  const bool x16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const bool out16 = (reinterpret_cast<uintptr_t>(raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(pk) & 7) == 0;
  if (fast && n > 512 && stride_d == 1 && d % 4 == 0 && stride_n % 4 == 0 && (uniform || stride_b % 4 == 0) && x16 && out16) {
    // d <= 768 runs the instance with half the row registers (60 instead of 80-92): four CTAs per SM instead of three
    // took this kernel from 0.87 to 1.0 of the measured copy peak (0.78 -> 0.96 without the second plane); a persistent
    // software-pipelined variant (next row requested before the current one is finished) was measured before that
    // and is gone: 0.82-0.85, it needs the registers of two rows
    {
      if (d <= 4 * 32 * 6)
        pack_rm_kernel<kHalf, 6><<<static_cast<unsigned>((n + 7) / 8), kPackThreads, 0, stream>>>(
            x, n, d, fm, raw, norms, pk, err, stats, zero_words, n_zero, rf);
      else
        pack_rm_kernel<kHalf, kRmMaxV><<<static_cast<unsigned>((n + 7) / 8), kPackThreads, 0, stream>>>(
            x, n, d, fm, raw, norms, pk, err, stats, zero_words, n_zero, rf);
    }
  } else if (fast && uniform && n > 8192 && stride_n == 1 && stride_d % 4 == 0 && x16) {
    const size_t smem = static_cast<size_t>(d) * kCmLd * sizeof(float);
    const long long n_tiles = (n + 31) / 32;
    // ALIVE_KNN_PACK_DOUBLE=1: the persistent double-buffered variant (an experiment that lost: with one
    // 8-warp CTA per SM the finishing pass cannot hide its arithmetic latency - 55 % of the HBM peak
    // against 73-77 % for two independent CTAs per SM)
    static const int dbl = getenv("ALIVE_KNN_PACK_DOUBLE") && atoi(getenv("ALIVE_KNN_PACK_DOUBLE")) == 1;
    int num_sms = 148;
    if (dbl) {      // (per device: only the experiment needs it)
      int dev = 0;
      ALIVE_CHECK_CUDA(cudaGetDevice(&dev));
      ALIVE_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (dbl && d <= 768 && n_tiles >= 4ll * num_sms) {
      pack_cm_kernel<true, kHalf><<<static_cast<unsigned>(num_sms), kPackThreads, 2 * smem, stream>>>(
          x, n, d, stride_d, raw, norms, pk, err, stats, zero_words, n_zero, rf);
    } else {
      // 16 warps per tile by default (0.84-0.89 of the HBM peak against 0.82-0.855 for the 8-warp kernel, same bits);
      // ALIVE_KNN_PACK_CM2=0 runs the 8-warp kernel for A/B
      static const int cm2 = getenv("ALIVE_KNN_PACK_CM2") ? atoi(getenv("ALIVE_KNN_PACK_CM2")) : 1;
      if (cm2) {
        static PerDeviceOnce cm2_once;
        const int rc2 = cm2_once.run([]() -> int {
          ALIVE_CHECK_CUDA((cudaFuncSetAttribute(pack_cm2_kernel<kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * kCmLd * 4)));
          return 0;
        });
        if (rc2) return rc2;
        pack_cm2_kernel<kHalf><<<static_cast<unsigned>(n_tiles), kCm2Threads, smem, stream>>>(
            x, n, d, stride_d, raw, norms, pk, err, stats, zero_words, n_zero, rf);
      } else
      pack_cm_kernel<false, kHalf><<<static_cast<unsigned>(n_tiles), kPackThreads, smem, stream>>>(
          x, n, d, stride_d, raw, norms, pk, err, stats, zero_words, n_zero, rf);
    }
  } else if (n <= 512) {
    pack_frame_kernel<kHalf><<<static_cast<unsigned>(n), kPackThreads, 0, stream>>>(x, n, d, fm, stride_d, raw, norms, pk, err,
                                                                                   stats, zero_words, n_zero, rf);
  } else if (n <= 8192) {
    const size_t smem = static_cast<size_t>(d) * 9 * sizeof(float);
    pack_kernel<8, kHalf><<<static_cast<unsigned>((n + 7) / 8), kPackThreads, smem, stream>>>(
        x, n, d, fm, stride_d, raw, norms, pk, err, stats, zero_words, n_zero, async_stage, rf);
  } else {
    const size_t smem = static_cast<size_t>(d) * 33 * sizeof(float);
    pack_kernel<32, kHalf><<<static_cast<unsigned>((n + 31) / 32), kPackThreads, smem, stream>>>(
        x, n, d, fm, stride_d, raw, norms, pk, err, stats, zero_words, n_zero, async_stage, rf);
  }
  ALIVE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

int pack_impl(const float* x, int64_t n, int32_t d, int64_t stride_n, int64_t stride_d, float* raw, float* norms,
              uint16_t* packed, float* err, uint32_t* stats, int32_t* zero_words, int32_t n_zero,
              alive_stream_t stream, int64_t item_frames, int64_t stride_b, uint16_t* lo, float* err2, int32_t format) {
  ALIVE_REQUIRE(x && raw && norms && packed, "alive_knn_pack: NULL argument");
  ALIVE_REQUIRE(lo == nullptr || (reinterpret_cast<uintptr_t>(lo) & 7) == 0, "alive_knn_pack: lo must be 8-byte aligned");
  ALIVE_REQUIRE(format == ALIVE_KNN_FORMAT_BF16 || format == ALIVE_KNN_FORMAT_FP16, "alive_knn_pack: unknown plane format %d", format);
  const Refine rf{lo, err2};
  if (item_frames <= 0 || item_frames >= n) {       // one item: the plain [n] frame sequence
    item_frames = n > 0 ? n : 1;
    stride_b = 0;
  }
  ALIVE_REQUIRE(n % item_frames == 0, "alive_knn_pack: n must be a multiple of the frames per item");
  const bool one_item = item_frames == n || n == 0;
  // items laid out with a uniform frame stride are one plain sequence
  const bool uniform = one_item || stride_b == item_frames * stride_n;
  const FrameMap fm{static_cast<int>(uniform ? (n > 0 ? n : 1) : item_frames), uniform ? 0 : stride_b, stride_n};
  ALIVE_REQUIRE(n >= 0 && n < (1ll << 31), "alive_knn_pack: n out of range (%lld)", static_cast<long long>(n));
  ALIVE_REQUIRE(d >= 2 && d % 2 == 0 && d <= 1536, "alive_knn_pack: d must be even and <= 1536 (got %d)", d);
  ALIVE_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 3) == 0, "alive_knn_pack: packed must be 4-byte aligned");
  if (n == 0) {
    if (n_zero > 0) ALIVE_CHECK_CUDA(cudaMemsetAsync(zero_words, 0, sizeof(int32_t) * n_zero, as_stream(stream)));
    return 0;
  }
  if (format == ALIVE_KNN_FORMAT_FP16)
    return pack_dispatch<true>(x, n, d, fm, uniform, stride_n, stride_d, stride_b, raw, norms, packed, err, stats, zero_words,
                               n_zero, rf, as_stream(stream));
  return pack_dispatch<false>(x, n, d, fm, uniform, stride_n, stride_d, stride_b, raw, norms, packed, err, stats, zero_words,
                              n_zero, rf, as_stream(stream));
}
}  // namespace alive

extern "C" int alive_knn_pack(const float* x, int64_t n, int32_t d, int64_t stride_n, int64_t stride_d,
                              float* raw, float* norms, uint16_t* packed, float* err, uint32_t* stats,
                              uint16_t* lo, float* err2, int32_t format, alive_stream_t stream) {
  return alive::pack_impl(x, n, d, stride_n, stride_d, raw, norms, packed, err, stats, nullptr, 0, stream, 0, 0, lo, err2,
                          format);
}
