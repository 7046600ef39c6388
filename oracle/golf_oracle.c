/*
 * golf_oracle.c -- CPU restatement of the sample-recurrent parts of GOLF's
 * synthesis hot path.  TEST INFRASTRUCTURE ONLY: nothing under golf_b200/ may
 * import, link or execute this file; it exists so tests/, smoke() and the
 * cpu_baseline leg of bench.py have something independent to check the CUDA
 * path against and to time on the host cores.
 *
 * What is restated and where it comes from (paths under /root/reference):
 *
 *  - oracle_sample_wise_lpc_{f32,f64}: the time-varying all-pole recurrence
 *    that models/filters.py:112 (and :789, models/lru/lru.py:15) obtains from
 *    the third-party package `torchlpc` (requirements.txt:19, unpinned, not
 *    vendored, not installable here).  PARITY UNPINNED for this function: the
 *    reference tree holds no golden value for it.  Restated from its call-site
 *    contract: y[b,t] = x[b,t] - sum_{i<M} A[b,t,i] * y[b,t-1-i], zero (or zi)
 *    initial state, sign fixed by models/lpc.py:11-16 and filters.py:189-193,
 *    taps visited newest-first with a rounded multiply and a rounded subtract
 *    per tap (the shape of torchlpc's numba CPU loop), batch rows in parallel
 *    (its prange).
 *
 *  - oracle_allpole_lti_f32: the per-channel LTI recurrence behind
 *    torchaudio.functional.lfilter as used by models/lpc.py:11-16,118.  The
 *    native loop lives in libtorchaudio (installed, 2.11.0); its arithmetic was
 *    pinned empirically: taps visited OLDEST-first, separate rounded multiply
 *    and subtract -- this restatement is bit-identical to lfilter on CPU
 *    (tests/test_oracle.py::test_lti_matches_torchaudio_bitwise).
 *
 *  - oracle_biquad_cascade_f32: models/lpc.py:115-118, a loop of 3-tap lfilter
 *    calls (each normalised by its own a0 like lfilter does).
 *
 *  - oracle_linear_upsample_f32: models/audiotensor/audiotensor.py:11-17
 *    (F.interpolate, mode="linear", align_corners=True), ATen CPU arithmetic
 *    pinned empirically: src = fl(scale*dst), l1 = src - floor(src),
 *    l0 = 1 - l1, out = fma(l0, x[i0], fl(l1 * x[i1])).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- time-varying all-pole recurrence (torchlpc.sample_wise_lpc contract) ---- */

#define DEFINE_SAMPLE_WISE_LPC(NAME, REAL)                                         \
  ORACLE_API void NAME(const REAL *x, const REAL *A, const REAL *zi, REAL *y,      \
                       int64_t B, int64_t T, int64_t M) {                          \
    _Pragma("omp parallel for schedule(static)")                                   \
    for (int64_t b = 0; b < B; ++b) {                                              \
      REAL *buf = (REAL *)malloc((size_t)(T + M) * sizeof(REAL));                  \
      /* buf[M + t] = y[t]; buf[M-1-j] = y[-1-j] = zi[b, j] */                     \
      for (int64_t j = 0; j < M; ++j) buf[M - 1 - j] = zi ? zi[b * M + j] : (REAL)0; \
      const REAL *xb = x + b * T;                                                  \
      const REAL *Ab = A + b * T * M;                                              \
      for (int64_t t = 0; t < T; ++t) {                                            \
        REAL acc = xb[t];                                                          \
        const REAL *at = Ab + t * M;                                               \
        const REAL *yt = buf + M + t - 1;                                          \
        for (int64_t i = 0; i < M; ++i) {                                          \
          REAL p = at[i] * yt[-i];                                                 \
          acc = acc - p;                                                           \
        }                                                                          \
        buf[M + t] = acc;                                                          \
      }                                                                            \
      memcpy(y + b * T, buf + M, (size_t)T * sizeof(REAL));                        \
      free(buf);                                                                   \
    }                                                                              \
  }

DEFINE_SAMPLE_WISE_LPC(oracle_sample_wise_lpc_f32, float)
DEFINE_SAMPLE_WISE_LPC(oracle_sample_wise_lpc_f64, double)

/* ---- per-channel LTI all-pole recurrence (torchaudio lfilter IIR core) ---- */
/* x: [C, N], a: [C, M] (a_1..a_M, a_0 == 1 already divided out), y: [C, N]. */
ORACLE_API void oracle_allpole_lti_f32(const float *x, const float *a, float *y,
                                       int64_t C, int64_t N, int64_t M) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    float *buf = (float *)calloc((size_t)(N + M), sizeof(float));
    const float *xc = x + c * N;
    const float *ac = a + c * M;
    for (int64_t n = 0; n < N; ++n) {
      float acc = xc[n];
      /* buf[n + k] = y[n - M + k]; oldest first: k = 0 pairs with a_M */
      for (int64_t k = 0; k < M; ++k) {
        float p = ac[M - 1 - k] * buf[n + k];
        acc = acc - p;
      }
      buf[n + M] = acc;
    }
    memcpy(y + c * N, buf + M, (size_t)N * sizeof(float));
    free(buf);
  }
}

ORACLE_API void oracle_allpole_lti_f64(const double *x, const double *a, double *y,
                                       int64_t C, int64_t N, int64_t M) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    double *buf = (double *)calloc((size_t)(N + M), sizeof(double));
    const double *xc = x + c * N;
    const double *ac = a + c * M;
    for (int64_t n = 0; n < N; ++n) {
      double acc = xc[n];
      for (int64_t k = 0; k < M; ++k) acc -= ac[M - 1 - k] * buf[n + k];
      buf[n + M] = acc;
    }
    memcpy(y + c * N, buf + M, (size_t)N * sizeof(double));
    free(buf);
  }
}

/* ---- cascade of K second-order all-pole sections (models/lpc.py:115-118) ---- */
/* x: [C, N]; biquads: [C, K, 3] = (a0, a1, a2) per section; y: [C, N].
 * lfilter divides a (and b = [1,0,0]) by a0 first, so every section computes
 *   v[n] = u[n] / a0 ... ; here b0/a0 multiplies the input and (a1/a0, a2/a0)
 * are the recurrence taps, visited oldest-first like oracle_allpole_lti_f32. */
ORACLE_API void oracle_biquad_cascade_f32(const float *x, const float *biquads, float *y,
                                          int64_t C, int64_t N, int64_t K) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    float *cur = (float *)malloc((size_t)N * sizeof(float));
    memcpy(cur, x + c * N, (size_t)N * sizeof(float));
    for (int64_t k = 0; k < K; ++k) {
      const float *q = biquads + (c * K + k) * 3;
      float a1 = q[1] / q[0], a2 = q[2] / q[0], b0 = 1.0f / q[0];
      float y1 = 0.f, y2 = 0.f;
      for (int64_t n = 0; n < N; ++n) {
        float acc = cur[n] * b0;
        float p2 = a2 * y2;
        acc = acc - p2;
        float p1 = a1 * y1;
        acc = acc - p1;
        y2 = y1;
        y1 = acc;
        cur[n] = acc;
      }
    }
    memcpy(y + c * N, cur, (size_t)N * sizeof(float));
    free(cur);
  }
}

/* ---- linear upsample, align_corners=True (audiotensor.py:11-17) ---- */
/* x: [R, n] -> out: [R, (n-1)*hop + 1] */
ORACLE_API void oracle_linear_upsample_f32(const float *x, float *out, int64_t R,
                                           int64_t n, int64_t hop) {
  int64_t L = (n - 1) * hop + 1;
  float scale = (L > 1) ? (float)(n - 1) / (float)(L - 1) : 0.f;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < R; ++r) {
    const float *xr = x + r * n;
    float *o = out + r * L;
    for (int64_t d = 0; d < L; ++d) {
      float src = scale * (float)d;
      int64_t i0 = (int64_t)floorf(src);
      if (i0 > n - 1) i0 = n - 1;
      float l1 = src - (float)i0;
      if (l1 < 0.f) l1 = 0.f;
      if (l1 > 1.f) l1 = 1.f;
      float l0 = 1.f - l1;
      int64_t i1 = i0 + (i0 < n - 1 ? 1 : 0);
      float p1 = l1 * xr[i1];
      o[d] = fmaf(l0, xr[i0], p1);
    }
  }
}

/* ---- fused convenience: GOLF-ss filter on frame-rate controls ---- */
/* (models/filters.py:99-113)  ex: [B, Tex], gain: [B, F], a: [B, F, M];
 * L = min(Tex, (F-1)*hop+1); y: [B, L].  Upsamples with the pinned ATen
 * arithmetic above, multiplies, runs the recurrence.  Coefficients are
 * interpolated on the fly (no [B, L, M] temporary) so this is also the leg
 * timed as the CPU baseline. */
ORACLE_API void oracle_lpc_ss_f32(const float *ex, const float *gain, const float *a,
                                  float *y, int64_t B, int64_t Tex, int64_t F,
                                  int64_t M, int64_t hop) {
  int64_t Lup = (F - 1) * hop + 1;
  int64_t L = Tex < Lup ? Tex : Lup;
  float scale = (Lup > 1) ? (float)(F - 1) / (float)(Lup - 1) : 0.f;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    float *buf = (float *)calloc((size_t)(L + M), sizeof(float));
    const float *gb = gain + b * F;
    const float *ab = a + b * F * M;
    const float *xb = ex + b * Tex;
    for (int64_t t = 0; t < L; ++t) {
      float src = scale * (float)t;
      int64_t i0 = (int64_t)floorf(src);
      if (i0 > F - 1) i0 = F - 1;
      float l1 = src - (float)i0;
      if (l1 < 0.f) l1 = 0.f;
      if (l1 > 1.f) l1 = 1.f;
      float l0 = 1.f - l1;
      int64_t i1 = i0 + (i0 < F - 1 ? 1 : 0);
      float g = fmaf(l0, gb[i0], l1 * gb[i1]);
      float acc = xb[t] * g;
      const float *a0 = ab + i0 * M, *a1 = ab + i1 * M;
      const float *yt = buf + M + t - 1;
      for (int64_t i = 0; i < M; ++i) {
        float p1 = l1 * a1[i];
        float c = fmaf(l0, a0[i], p1);
        float p = c * yt[-i];
        acc = acc - p;
      }
      buf[M + t] = acc;
    }
    memcpy(y + b * L, buf + M, (size_t)L * sizeof(float));
    free(buf);
  }
}

/* same in double from float inputs: the "truth" the parity tests report against */
ORACLE_API void oracle_lpc_ss_f64(const float *ex, const float *gain, const float *a,
                                  double *y, int64_t B, int64_t Tex, int64_t F,
                                  int64_t M, int64_t hop) {
  int64_t Lup = (F - 1) * hop + 1;
  int64_t L = Tex < Lup ? Tex : Lup;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    double *buf = (double *)calloc((size_t)(L + M), sizeof(double));
    const float *gb = gain + b * F;
    const float *ab = a + b * F * M;
    const float *xb = ex + b * Tex;
    for (int64_t t = 0; t < L; ++t) {
      int64_t i0 = t / hop;
      if (i0 > F - 1) i0 = F - 1;
      double l1 = (double)(t - i0 * hop) / (double)hop;
      double l0 = 1.0 - l1;
      int64_t i1 = i0 + (i0 < F - 1 ? 1 : 0);
      double g = l0 * gb[i0] + l1 * gb[i1];
      double acc = (double)xb[t] * g;
      const float *a0 = ab + i0 * M, *a1 = ab + i1 * M;
      const double *yt = buf + M + t - 1;
      for (int64_t i = 0; i < M; ++i) acc -= (l0 * a0[i] + l1 * a1[i]) * yt[-i];
      buf[M + t] = acc;
    }
    memcpy(y + b * L, buf + M, (size_t)L * sizeof(double));
    free(buf);
  }
}
