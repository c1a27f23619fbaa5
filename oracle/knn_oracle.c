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
 *             double (dot_f64) and rounded once - the value the reference's float32 sgemm approximates (the
 *             two differ by a few 1e-8; index parity is defined modulo ties within 1e-6)
 *   :105      top-k, largest first, NaN above everything, ties -> lowest index
 *   :107      RAW frames of the k winners summed sequentially in float32 in descending-score order,
 *             divided by k (the bit-exact model of torch's mean(dim=2))
 *   :108-109  result*(1-alpha) + input*alpha with separately rounded products
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/c_oracle.py); -ffp-contract=off keeps the
 * compiler from fusing the separately rounded float32 operations above.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>

/* sum of the exact products sn[j]*r[j] in double, rounded once by the caller.  Eight independent partial sums
 * (so the loop runs at throughput, not at the latency of one add chain; the large parity cases of
 * tests/test_gpu_fullsize.py need ~1e12 of these) combined at the end: every product of two floats is exact in
 * double and each partial sum carries < 1e-14 of rounding error, far below the 1e-6 tie tolerance of the parity
 * rule - the summation order is not part of the definition (the reference's own sgemm has none either). */
static double dot_f64(const float* a, const float* b, int32_t d) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
  int32_t j = 0;
  for (; j + 8 <= d; j += 8) {
    s0 += (double)a[j] * (double)b[j];
    s1 += (double)a[j + 1] * (double)b[j + 1];
    s2 += (double)a[j + 2] * (double)b[j + 2];
    s3 += (double)a[j + 3] * (double)b[j + 3];
    s4 += (double)a[j + 4] * (double)b[j + 4];
    s5 += (double)a[j + 5] * (double)b[j + 5];
    s6 += (double)a[j + 6] * (double)b[j + 6];
    s7 += (double)a[j + 7] * (double)b[j + 7];
  }
  for (; j < d; ++j) s0 += (double)a[j] * (double)b[j];
  return ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
}

static int better(float sa, int64_t ia, float sb, int64_t ib) {
  const int na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return ia < ib;
}

/* source [B,D,T] contiguous, reference [RB,D,N] contiguous with RB == B or RB == 1 (a shared library,
 * VoiceLibrary.match).  out [B,D,T] (may be NULL), idx [B,T,k], val [B,T,k] (may be NULL).
 * rows != 0: the same data handed over frame-major (source [B,T,D], reference [RB,N,D], out [B,T,D]) - what
 * the GPU side keeps; saves two transposes of multi-GB libraries in the large parity tests, same arithmetic.
 * Returns 0, or -1 for "selected index k out of range" (k > N or N == 0), -2 for a batch mismatch. */
int alive_oracle_match2(const float* source, const float* reference, int32_t B, int32_t RB, int32_t D, int32_t T,
                        int32_t N, int32_t k, float alpha, float* out, int64_t* idx, float* val, int32_t rows) {
  if (RB != B && RB != 1) return -2;
  if (k < 1 || k > N || N == 0) return -1;
  if (T == 0) return 0;
  const float a1 = (float)(1.0 - (double)alpha), a0 = alpha;
  /* element (channel j, frame f): source src[j*SD + f*ST], library ref[j*RD + f*RN], result out[j*SD + f*ST] */
  const size_t SD = rows ? 1 : (size_t)T, ST = rows ? (size_t)D : 1;
  const size_t RD = rows ? 1 : (size_t)N, RN = rows ? (size_t)D : 1;
  for (int32_t b = 0; b < B; ++b) {
    const float* src = source + (size_t)b * D * T;
    const float* ref = reference + (size_t)(RB == 1 ? 0 : b) * D * N;
    /* library frames, normalised in float32, row-major for the scan */
    float* rn = (float*)malloc((size_t)N * D * sizeof(float));
    if (!rn) return -3;
    /* (blocks of 64 frames: the channel-major rows are read 256 bytes at a time instead of one float per line) */
#pragma omp parallel for schedule(static)
    for (int32_t n0 = 0; n0 < N; n0 += 64) {
      const int32_t nb = N - n0 < 64 ? N - n0 : 64;
      double ss[64];
      float nrm[64];
      for (int32_t i = 0; i < nb; ++i) ss[i] = 0.0;
      for (int32_t j = 0; j < D; ++j)                       /* per frame: j ascending, as before */
        for (int32_t i = 0; i < nb; ++i) ss[i] += (double)ref[j * RD + (n0 + i) * RN] * (double)ref[j * RD + (n0 + i) * RN];
      for (int32_t i = 0; i < nb; ++i) nrm[i] = (float)sqrt(ss[i]);
      for (int32_t j = 0; j < D; ++j)
        for (int32_t i = 0; i < nb; ++i) rn[(size_t)(n0 + i) * D + j] = ref[j * RD + (n0 + i) * RN] / nrm[i];
    }
    /* a task = a block of QB query frames that streams the normalised library ONCE (the frame in cache is
     * scored against all QB queries): same arithmetic as one query at a time, 1/QB of the memory traffic */
    int32_t QB = (T + omp_get_max_threads() - 1) / omp_get_max_threads();
    if (QB < 1) QB = 1;
    if (QB > 8) QB = 8;
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t tb = 0; tb < T; tb += QB) {
      const int32_t nq = T - tb < QB ? T - tb : QB;
      float* sn = (float*)malloc((size_t)nq * D * sizeof(float));
      float* bs = (float*)malloc((size_t)nq * k * sizeof(float));
      int64_t* bi = (int64_t*)malloc((size_t)nq * k * sizeof(int64_t));
      int32_t* have = (int32_t*)calloc((size_t)nq, sizeof(int32_t));
      for (int32_t q = 0; q < nq; ++q) {
        const int32_t t = tb + q;
        double ss = 0.0;
        for (int32_t j = 0; j < D; ++j) ss += (double)src[j * SD + t * ST] * (double)src[j * SD + t * ST];
        const float nrm = (float)sqrt(ss);
        for (int32_t j = 0; j < D; ++j) sn[(size_t)q * D + j] = src[j * SD + t * ST] / nrm;
      }
      for (int32_t n = 0; n < N; ++n) {
        const float* r = rn + (size_t)n * D;
        for (int32_t q = 0; q < nq; ++q) {
          const float s = (float)dot_f64(sn + (size_t)q * D, r, D);
          float* qs = bs + (size_t)q * k;
          int64_t* qi = bi + (size_t)q * k;
          if (have[q] < k || better(s, n, qs[have[q] - 1], qi[have[q] - 1])) {
            int32_t p = have[q] < k ? have[q]++ : k - 1;
            while (p > 0 && better(s, n, qs[p - 1], qi[p - 1])) {
              qs[p] = qs[p - 1];
              qi[p] = qi[p - 1];
              --p;
            }
            qs[p] = s;
            qi[p] = n;
          }
        }
      }
      for (int32_t q = 0; q < nq; ++q) {
        const int32_t t = tb + q;
        const float* qs = bs + (size_t)q * k;
        const int64_t* qi = bi + (size_t)q * k;
        for (int32_t r = 0; r < k; ++r) {
          idx[((size_t)b * T + t) * k + r] = qi[r];
          if (val) val[((size_t)b * T + t) * k + r] = qs[r];
        }
        if (out) {
          for (int32_t j = 0; j < D; ++j) {
            float acc = ref[j * RD + (size_t)qi[0] * RN];
            for (int32_t r = 1; r < k; ++r) acc = acc + ref[j * RD + (size_t)qi[r] * RN];
            const float mean = acc / (float)k;
            const float p1 = mean * a1;
            const float p0 = src[j * SD + t * ST] * a0;
            out[(size_t)b * D * T + j * SD + t * ST] = p1 + p0;
          }
        }
      }
      free(sn);
      free(bs);
      free(bi);
      free(have);
    }
    free(rn);
  }
  return 0;
}

int alive_oracle_match(const float* source, const float* reference, int32_t B, int32_t RB, int32_t D, int32_t T,
                       int32_t N, int32_t k, float alpha, float* out, int64_t* idx, float* val) {
  return alive_oracle_match2(source, reference, B, RB, D, T, N, k, alpha, out, idx, val, 0);
}
