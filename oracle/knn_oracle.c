/* CPU oracle, C restatement - TEST INFRASTRUCTURE ONLY (see oracle/knn_oracle.py for the rules: only
 * tests/, __graft_entry__.smoke() and bench.py's CPU arm may build, load or call anything under oracle/).
 *
 * Restates uthree/ALiVE-VC module/common.py:96-109 (= module/voice_library.py:15-33) step by step,
 * independently of numpy/torch, so that the numpy oracle and this one check each other and larger
 * parity cases finish in seconds (OpenMP over query frames):
 *
 *   :100-101  source [B,D,T], reference [B,D,N] are read through their channel-major layout
 *   :102-103  L2 norm of every frame (no epsilon)
 *   :104      (s/|s|) . (r/|r|): frames normalised in float32 (IEEE division), products accumulated in
 *             double and rounded once - the value the reference's float32 sgemm approximates (the two
 *             differ by a few 1e-8; index parity is defined modulo ties within 1e-6)
 *   :105      top-k, largest first, NaN above everything, ties -> lowest index
 *   :107      RAW frames of the k winners summed sequentially in float32 in descending-score order,
 *             divided by k (the bit-exact model of torch's mean(dim=2))
 *   :108-109  result*(1-alpha) + input*alpha with separately rounded products
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/c_oracle.py); -ffp-contract=off keeps the
 * compiler from fusing the separately rounded float32 operations above.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int better(float sa, int64_t ia, float sb, int64_t ib) {
  const int na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return ia < ib;
}

/* source [B,D,T] contiguous, reference [RB,D,N] contiguous with RB == B or RB == 1 (a shared library,
 * VoiceLibrary.match).  out [B,D,T] (may be NULL), idx [B,T,k], val [B,T,k] (may be NULL).
 * Returns 0, or -1 for "selected index k out of range" (k > N or N == 0), -2 for a batch mismatch. */
int alive_oracle_match(const float* source, const float* reference, int32_t B, int32_t RB, int32_t D, int32_t T,
                       int32_t N, int32_t k, float alpha, float* out, int64_t* idx, float* val) {
  if (RB != B && RB != 1) return -2;
  if (k < 1 || k > N || N == 0) return -1;
  if (T == 0) return 0;
  const float a1 = (float)(1.0 - (double)alpha), a0 = alpha;
  for (int32_t b = 0; b < B; ++b) {
    const float* src = source + (size_t)b * D * T;
    const float* ref = reference + (size_t)(RB == 1 ? 0 : b) * D * N;
    /* library frames, normalised in float32, row-major for the scan */
    float* rn = (float*)malloc((size_t)N * D * sizeof(float));
    if (!rn) return -3;
#pragma omp parallel for schedule(static)
    for (int32_t n = 0; n < N; ++n) {
      double ss = 0.0;
      for (int32_t j = 0; j < D; ++j) ss += (double)ref[(size_t)j * N + n] * (double)ref[(size_t)j * N + n];
      const float nrm = (float)sqrt(ss);
      for (int32_t j = 0; j < D; ++j) rn[(size_t)n * D + j] = ref[(size_t)j * N + n] / nrm;
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int32_t t = 0; t < T; ++t) {
      float* sn = (float*)malloc((size_t)D * sizeof(float));
      float* bs = (float*)malloc((size_t)k * sizeof(float));
      int64_t* bi = (int64_t*)malloc((size_t)k * sizeof(int64_t));
      double ss = 0.0;
      for (int32_t j = 0; j < D; ++j) ss += (double)src[(size_t)j * T + t] * (double)src[(size_t)j * T + t];
      const float nrm = (float)sqrt(ss);
      for (int32_t j = 0; j < D; ++j) sn[j] = src[(size_t)j * T + t] / nrm;
      int32_t have = 0;
      for (int32_t n = 0; n < N; ++n) {
        double acc = 0.0;
        const float* r = rn + (size_t)n * D;
        for (int32_t j = 0; j < D; ++j) acc += (double)sn[j] * (double)r[j];
        const float s = (float)acc;
        if (have < k || better(s, n, bs[have - 1], bi[have - 1])) {
          int32_t p = have < k ? have++ : k - 1;
          while (p > 0 && better(s, n, bs[p - 1], bi[p - 1])) {
            bs[p] = bs[p - 1];
            bi[p] = bi[p - 1];
            --p;
          }
          bs[p] = s;
          bi[p] = n;
        }
      }
      for (int32_t r = 0; r < k; ++r) {
        idx[((size_t)b * T + t) * k + r] = bi[r];
        if (val) val[((size_t)b * T + t) * k + r] = bs[r];
      }
      if (out) {
        for (int32_t j = 0; j < D; ++j) {
          float acc = ref[(size_t)j * N + bi[0]];
          for (int32_t r = 1; r < k; ++r) acc = acc + ref[(size_t)j * N + bi[r]];
          const float mean = acc / (float)k;
          const float p1 = mean * a1;
          const float p0 = src[(size_t)j * T + t] * a0;
          out[((size_t)b * D + j) * T + t] = p1 + p0;
        }
      }
      free(sn);
      free(bs);
      free(bi);
    }
    free(rn);
  }
  return 0;
}
