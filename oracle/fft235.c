/* fft235.c -- plain-C in-place batched complex FFT (Stockham autosort, mixed radix 2/3/4/5 plus a
 * generic O(p^2) butterfly for any other prime factor), float and double.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/). It stands in for FFTW3, which the reference links
 * (tensorflow_nufft/cc/kernels/fftw_api.h:27-205; call sites nufft_plan.cc:336,413) but which is
 * not in this image. It backs (a) the `fftw3.h` shim under oracle/ref_build so that the reference's
 * own CPU plan can run, and (b) the C restatement in oracle/nufft_oracle.c. The product path never
 * calls it (the engine uses cuFFT).
 *
 * Semantics follow fftw_plan_many_dft as the reference uses it (nufft_plan.cc:413-426): `rank`
 * dims n[0..rank-1] in row-major order (last fastest), `howmany` transforms `dist` elements apart,
 * in place, unnormalised, exponent sign `sign` (-1 forward, +1 backward).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FFT235_BLOCK 16

static int fft235_factorize(int n, int* fac) {
  int nf = 0;
  while (n % 4 == 0) { fac[nf++] = 4; n /= 4; }
  while (n % 2 == 0) { fac[nf++] = 2; n /= 2; }
  while (n % 3 == 0) { fac[nf++] = 3; n /= 3; }
  while (n % 5 == 0) { fac[nf++] = 5; n /= 5; }
  for (int p = 7; n > 1; p += 2) {
    while (n % p == 0) { fac[nf++] = p; n /= p; }
  }
  return nf;
}

#define DEFINE_FFT235(T, SUFFIX)                                                                   \
  typedef struct { T re, im; } cplx_##SUFFIX;                                                      \
  /* One Stockham pass of radix r: n_cur = current sub-length, s = current stride. */              \
  static void pass_##SUFFIX(int n, int n_cur, int s, int r, const cplx_##SUFFIX* x,                \
                            cplx_##SUFFIX* y, const cplx_##SUFFIX* w) {                            \
    int m = n_cur / r;                                                                             \
    int wstep = n / n_cur; /* w[k*wstep] = exp(sign*2*pi*i*k/n_cur) */                             \
    int rstep = n / r;     /* w[k*rstep] = exp(sign*2*pi*i*k/r)     */                             \
    for (int p = 0; p < m; ++p) {                                                                  \
      for (int q = 0; q < s; ++q) {                                                                \
        cplx_##SUFFIX a[64];                                                                       \
        for (int t = 0; t < r; ++t) a[t] = x[q + s * (p + t * m)];                                 \
        if (r == 2) {                                                                              \
          cplx_##SUFFIX b0 = {a[0].re + a[1].re, a[0].im + a[1].im};                               \
          cplx_##SUFFIX b1 = {a[0].re - a[1].re, a[0].im - a[1].im};                               \
          cplx_##SUFFIX t1 = w[p * wstep];                                                         \
          y[q + s * (2 * p)] = b0;                                                                 \
          y[q + s * (2 * p + 1)].re = b1.re * t1.re - b1.im * t1.im;                               \
          y[q + s * (2 * p + 1)].im = b1.re * t1.im + b1.im * t1.re;                               \
        } else {                                                                                   \
          for (int u = 0; u < r; ++u) {                                                            \
            T br = 0, bi = 0;                                                                      \
            for (int t = 0; t < r; ++t) {                                                          \
              cplx_##SUFFIX o = w[((long)t * u % r) * rstep];                                      \
              br += a[t].re * o.re - a[t].im * o.im;                                               \
              bi += a[t].re * o.im + a[t].im * o.re;                                               \
            }                                                                                      \
            cplx_##SUFFIX tw = w[((long)p * u % n_cur) * wstep];                                   \
            y[q + s * (r * p + u)].re = br * tw.re - bi * tw.im;                                   \
            y[q + s * (r * p + u)].im = br * tw.im + bi * tw.re;                                   \
          }                                                                                        \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
  /* 1D FFT of length n on buf0 (scratch buf1). Returns pointer to the buffer holding the result. */\
  static cplx_##SUFFIX* fft1d_##SUFFIX(int n, cplx_##SUFFIX* buf0, cplx_##SUFFIX* buf1,            \
                                       const cplx_##SUFFIX* w, const int* fac, int nfac) {         \
    int n_cur = n, s = 1;                                                                          \
    cplx_##SUFFIX *x = buf0, *y = buf1;                                                            \
    for (int i = 0; i < nfac; ++i) {                                                               \
      pass_##SUFFIX(n, n_cur, s, fac[i], x, y, w);                                                 \
      n_cur /= fac[i];                                                                             \
      s *= fac[i];                                                                                 \
      cplx_##SUFFIX* tmp = x; x = y; y = tmp;                                                      \
    }                                                                                              \
    return x;                                                                                      \
  }                                                                                                \
  void fft235_##SUFFIX(T* data_, int rank, const int* dims, int howmany, long dist, int sign,      \
                       int nthreads) {                                                             \
    cplx_##SUFFIX* data = (cplx_##SUFFIX*)data_;                                                   \
    int n3[3] = {1, 1, 1};                                                                         \
    for (int i = 0; i < rank; ++i) n3[3 - rank + i] = dims[i];                                     \
    long tot = (long)n3[0] * n3[1] * n3[2];                                                        \
    long stride3[3] = {(long)n3[1] * n3[2], (long)n3[2], 1};                                       \
    if (nthreads < 1) nthreads = 1;                                                                \
    for (int ax = 2; ax >= 0; --ax) {                                                              \
      int n = n3[ax];                                                                              \
      if (n == 1) continue;                                                                        \
      long st = stride3[ax];                                                                       \
      int fac[64];                                                                                 \
      int nfac = fft235_factorize(n, fac);                                                         \
      cplx_##SUFFIX* w = (cplx_##SUFFIX*)malloc(sizeof(cplx_##SUFFIX) * n);                        \
      for (int k = 0; k < n; ++k) {                                                                \
        double ang = sign * 2.0 * M_PI * (double)k / (double)n;                                    \
        w[k].re = (T)cos(ang);                                                                     \
        w[k].im = (T)sin(ang);                                                                     \
      }                                                                                            \
      /* A "line" is addressed by (outer, inner): element k at base + k*st, base = outer*st*n +    \
       * inner, inner in [0, st). Lines with adjacent `inner` are gathered together. */             \
      long n_outer = tot / (st * n);                                                               \
      long blocks_per_outer = (st + FFT235_BLOCK - 1) / FFT235_BLOCK;                              \
      long nblocks = (long)howmany * n_outer * blocks_per_outer;                                   \
      _Pragma("omp parallel num_threads(nthreads)")                                                \
      {                                                                                            \
        cplx_##SUFFIX* b0 = (cplx_##SUFFIX*)malloc(sizeof(cplx_##SUFFIX) * n * FFT235_BLOCK);      \
        cplx_##SUFFIX* b1 = (cplx_##SUFFIX*)malloc(sizeof(cplx_##SUFFIX) * n);                     \
        cplx_##SUFFIX* b2 = (cplx_##SUFFIX*)malloc(sizeof(cplx_##SUFFIX) * n);                     \
        _Pragma("omp for schedule(static)")                                                        \
        for (long blk = 0; blk < nblocks; ++blk) {                                                 \
          long bi = blk % blocks_per_outer;                                                        \
          long rest = blk / blocks_per_outer;                                                      \
          long outer = rest % n_outer;                                                             \
          long batch = rest / n_outer;                                                             \
          long inner0 = bi * FFT235_BLOCK;                                                         \
          int nl = (int)((st - inner0) < FFT235_BLOCK ? (st - inner0) : FFT235_BLOCK);             \
          cplx_##SUFFIX* base = data + batch * dist + outer * st * n + inner0;                     \
          if (st == 1) {                                                                           \
            memcpy(b1, base, sizeof(cplx_##SUFFIX) * n);                                           \
            cplx_##SUFFIX* r = fft1d_##SUFFIX(n, b1, b2, w, fac, nfac);                            \
            memcpy(base, r, sizeof(cplx_##SUFFIX) * n);                                            \
          } else {                                                                                 \
            for (int k = 0; k < n; ++k)                                                            \
              for (int l = 0; l < nl; ++l) b0[(long)l * n + k] = base[(long)k * st + l];           \
            for (int l = 0; l < nl; ++l) {                                                         \
              memcpy(b1, b0 + (long)l * n, sizeof(cplx_##SUFFIX) * n);                             \
              cplx_##SUFFIX* r = fft1d_##SUFFIX(n, b1, b2, w, fac, nfac);                          \
              memcpy(b0 + (long)l * n, r, sizeof(cplx_##SUFFIX) * n);                              \
            }                                                                                      \
            for (int k = 0; k < n; ++k)                                                            \
              for (int l = 0; l < nl; ++l) base[(long)k * st + l] = b0[(long)l * n + k];           \
          }                                                                                        \
        }                                                                                          \
        free(b0); free(b1); free(b2);                                                              \
      }                                                                                            \
      free(w);                                                                                     \
    }                                                                                              \
  }

DEFINE_FFT235(float, f32)
DEFINE_FFT235(double, f64)
