/* nufft_oracle.c -- CPU restatement of the reference's NUFFT algorithm for the tfft.nufft hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under tensorflow_nufft_b200/ may call into this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg (and only when oracle/_ref is
 * absent) use it, as the checker. Plain C (gnu11), float and double instantiated from one macro
 * body. Compiled WITHOUT -ffast-math and without FMA contraction (-ffp-contract=off, x86-64
 * baseline), like the reference library (/root/reference/Makefile:38).
 *
 * PINNING: checked against (a) the reference's own CPU plan compiled from /root/reference
 * (oracle/_ref/libref.so; tests/test_oracle.py::test_port_matches_compiled_reference_*), and
 * (b) the golden vectors that library produced (tests/golden/, made by tests/golden/make_golden.py).
 *
 * The variant restated is "reference CPU code driven with the GPU plan's parameters"
 * (SURVEY.md finding 3): upsampling factor 2.0, direct exp(sqrt) kernel evaluation.
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/tensorflow_nufft/cc/kernels/).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void fft235_f32(float* data, int rank, const int* dims, int howmany, long dist, int sign, int nthreads);
void fft235_f64(double* data, int rank, const int* dims, int howmany, long dist, int sign, int nthreads);

/* next_smooth_int, nufft_util.cc:119-133 (== next_smooth_integer, nufft_plan.h:628-649). */
int oracle_next_smooth_int(int n) {
  if (n <= 2) return 2;
  if (n % 2 == 1) n += 1;
  int p = n - 2, d = 2;
  while (d > 1) {
    p += 2;
    d = p;
    while (d % 2 == 0) d /= 2;
    while (d % 3 == 0) d /= 3;
    while (d % 5 == 0) d /= 5;
  }
  return p;
}

/* Gauss-Legendre nodes/weights on [-1,1], ascending, weights normalised to sum 2. Stands in for
 * legendre_compute_glr (legendre_rule_fast.cc:28, LGPL -- not copied): Newton on the Legendre
 * recurrence. Agreement with the reference's nodes is checked through kernel_fseries outputs. */
void oracle_gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double t = cos(M_PI * (i + 0.75) / (n + 0.5));
    double dp = 0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; ++k) { double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = pk; }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
      double dt = p1 / dp;
      t -= dt;
      if (fabs(dt) < 1e-16 * fmax(1.0, fabs(t))) break;
    }
    { double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; ++k) { double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k; p0 = p1; p1 = pk; }
      dp = n * (t * p1 - p0) / (t * t - 1.0); }
    x[n - 1 - i] = t; x[i] = -t;
    w[n - 1 - i] = w[i] = 2.0 / ((1.0 - t) * (1.0 + t) * dp * dp);
  }
  if (n % 2 == 1) x[n / 2] = 0.0;
  double s = 0;
  for (int i = 0; i < n; ++i) s += w[i];
  for (int i = 0; i < n; ++i) w[i] = 2.0 * w[i] / s;
}

#define ORACLE_MAX_W 16

#define DEFINE_ORACLE(F, SFX, CPLX, CEXP, CLOG, EXPF, COSF, SINF, LOG10F, LOGF, FMODF, CEILF, FLOORF, FABSF, EPS, FFT) \
                                                                                                    \
  /* setup_spreader, nufft_plan.cu.cc:3040-3099 (CPU twin: nufft_plan.h:739-780 +                    \
   * nufft_plan.cc:885-947): width from tol, beta, c, all in FloatType. */                          \
  int oracle_kernel_params_##SFX(double tol_, double sigma, double* beta_out, double* c_out) {       \
    F eps = (F)tol_;                                                                                 \
    if (eps < (F)EPS) eps = (F)EPS;                                                                  \
    int ns = (int)ceil(-LOG10F(eps / (F)10.0));                                                      \
    if (sigma != 2.0) ns = (int)ceil(-LOGF(eps) / ((F)3.14159265358979329 * sqrt(1.0 - 1.0 / sigma))); \
    if (ns < 2) ns = 2;                                                                              \
    if (ns > ORACLE_MAX_W) ns = ORACLE_MAX_W;                                                        \
    F c = (F)(4.0 / (F)(ns * ns));                                                                   \
    F bon = (F)2.30;                                                                                 \
    if (ns == 2) bon = (F)2.20;                                                                      \
    if (ns == 3) bon = (F)2.26;                                                                      \
    if (ns == 4) bon = (F)2.38;                                                                      \
    if (sigma != 2.0) { F gamma = (F)0.97; bon = (F)(gamma * (F)3.14159265358979329 * (1 - 1 / (2 * sigma))); } \
    F beta = bon * (F)ns;                                                                            \
    *beta_out = (double)beta;                                                                        \
    *c_out = (double)c;                                                                              \
    return ns;                                                                                       \
  }                                                                                                  \
                                                                                                    \
  /* FoldAndRescale functors, nufft_plan.h:676-734. range 0 STRICT, 1 EXTENDED, 2 INFINITE. */       \
  void oracle_fold_rescale_##SFX(long M, const F* in, F* out, int range, int nf) {                   \
    const F pi = (F)3.14159265358979329, twopi = (F)6.283185307179586, i2pi = (F)0.159154943091895336; \
    for (long i = 0; i < M; ++i) {                                                                   \
      F x = in[i], s;                                                                                \
      if (range == 0) s = x + pi;                                                                    \
      else if (range == 1) { if (x > pi) s = x - pi; else if (x < -pi) s = x + (F)3.0 * pi; else s = x + pi; } \
      else { s = FMODF(x + pi, twopi); if (s < (F)0.0) s += twopi; }                                 \
      out[i] = s * i2pi * (F)nf;                                                                     \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* Bin-sort. rounding 0: CalcBinSizeNoGhost + CalcInvertofGlobalSortIdx kernels, nufft_plan.cu.cc:160-296,  \
   * with STABLE within-bin order (what the racy atomics give when they resolve in index order);   \
   * rounding 1: binsort_singlethread, nufft_plan.cc:475-531. pts = [rank][M] folded coords. */      \
  void oracle_binsort_##SFX(int rank, long M, const F* pts, const int* nf, const int* bin, int rounding, \
                            int* idx, int* bin_start, int* bin_sizes) {                              \
    int nb[3] = {1, 1, 1};                                                                           \
    long nbtot = 1;                                                                                  \
    for (int d = 0; d < rank; ++d) {                                                                 \
      nb[d] = rounding == 0 ? (nf[d] + bin[d] - 1) / bin[d] : nf[d] / bin[d] + 1;                    \
      nbtot *= nb[d];                                                                                \
    }                                                                                                \
    int* key = (int*)malloc(sizeof(int) * (M > 0 ? M : 1));                                          \
    memset(bin_sizes, 0, sizeof(int) * nbtot);                                                       \
    for (long i = 0; i < M; ++i) {                                                                   \
      int k = 0, mul = 1;                                                                            \
      for (int d = 0; d < rank; ++d) {                                                               \
        F x = pts[(long)d * M + i];                                                                  \
        int b;                                                                                       \
        if (rounding == 0) {                                                                         \
          b = (int)FLOORF(x / (F)bin[d]);                                                            \
          b = b >= nb[d] ? b - 1 : b;                                                                \
          b = b < 0 ? 0 : b;                                                                         \
        } else {                                                                                     \
          b = (int)(x / (F)bin[d]);                                                                  \
        }                                                                                            \
        k += mul * b;                                                                                \
        mul *= nb[d];                                                                                \
      }                                                                                              \
      key[i] = k;                                                                                    \
      bin_sizes[k]++;                                                                                \
    }                                                                                                \
    int run = 0;                                                                                     \
    for (long b = 0; b < nbtot; ++b) { bin_start[b] = run; run += bin_sizes[b]; }                    \
    int* cursor = (int*)malloc(sizeof(int) * nbtot);                                                 \
    memcpy(cursor, bin_start, sizeof(int) * nbtot);                                                  \
    for (long i = 0; i < M; ++i) idx[cursor[key[i]]++] = (int)i;                                     \
    free(cursor);                                                                                    \
    free(key);                                                                                       \
  }                                                                                                  \
                                                                                                    \
  /* evaluate_kernel, nufft_util.cc:64-69 (c*x*x in FloatType, then double). */                      \
  static F es_scalar_##SFX(F x, F beta, F c, F hw) {                                                 \
    if (FABSF(x) >= hw) return (F)0.0;                                                               \
    F cxx = c * x * x;                                                                               \
    return (F)exp(beta * sqrt(1.0 - cxx));                                                           \
  }                                                                                                  \
                                                                                                    \
  /* set_kernel_args + evaluate_kernel_vector, nufft_plan.cc:1244-1289 (direct evaluation). */       \
  static void es_vector_##SFX(F* ker, F x1, int ns, F beta, F c, F hw) {                             \
    for (int i = 0; i < ns; ++i) {                                                                   \
      F a = x1 + (F)i;                                                                               \
      F e = (F)(beta * sqrt(1.0 - c * a * a));                                                       \
      F k = (F)EXPF(e);                                                                              \
      if (FABSF(a) >= hw) k = (F)0.0;                                                                \
      ker[i] = k;                                                                                    \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* kernel_fseries_1d, nufft_util.cc:71-117, including the per-thread chunked phase winding        \
   * (nt = min(nout, num_threads) chunks, each restarted with pow()). */                             \
  void oracle_kernel_fseries_##SFX(int nf, int ns, double beta_, double c_, int num_threads, F* out) { \
    F beta = (F)beta_, c = (F)c_;                                                                    \
    F hw = (F)(ns / 2.0);                                                                            \
    int q = (int)(2 + 3.0 * hw);                                                                     \
    double z[2 * 100], w[2 * 100];                                                                   \
    F f[100];                                                                                        \
    CPLX a[100], aj[100];                                                                            \
    oracle_gauss_legendre(2 * q, z, w);                                                              \
    for (int n = 0; n < q; ++n) {                                                                    \
      z[n] *= hw;                                                                                    \
      f[n] = hw * (F)w[n] * es_scalar_##SFX((F)z[n], beta, c, (F)ns / 2);                            \
      F th = (F)2 * (F)3.14159265358979329 * (F)(nf / 2 - z[n]) / (F)nf;                             \
      /* exp(complex(0*v/nf, theta)) as libstdc++ std::exp(complex) -> cexp */                       \
      a[n] = CEXP((F)0.0 + th * I);                                                                  \
    }                                                                                                \
    int nout = nf / 2 + 1;                                                                           \
    int nt = nout < num_threads ? nout : num_threads;                                                \
    if (nt < 1) nt = 1;                                                                              \
    for (int t = 0; t < nt; ++t) {                                                                   \
      int b0 = (int)(0.5 + nout * t / (double)nt), b1 = (int)(0.5 + nout * (t + 1) / (double)nt);    \
      for (int n = 0; n < q; ++n) {                                                                  \
        /* std::pow(complex<T>, T) of libstdc++: polar(exp(y*log|x|), y*arg x) via clog */          \
        F y = (F)b0;                                                                                 \
        if (cimag(a[n]) == 0 && creal(a[n]) > 0) {                                                   \
          aj[n] = (F)pow(creal(a[n]), y);                                                            \
        } else {                                                                                     \
          CPLX lg = CLOG(a[n]);                                                                      \
          F rho = (F)EXPF(y * (F)creal(lg));                                                         \
          F th = y * (F)cimag(lg);                                                                   \
          aj[n] = rho * (F)COSF(th) + rho * (F)SINF(th) * I;                                         \
        }                                                                                            \
      }                                                                                              \
      for (int j = b0; j < b1; ++j) {                                                                \
        F x = (F)0.0;                                                                                \
        for (int n = 0; n < q; ++n) {                                                                \
          x += f[n] * 2 * (F)creal(aj[n]);                                                           \
          F ar = (F)creal(aj[n]), ai = (F)cimag(aj[n]), br = (F)creal(a[n]), bi = (F)cimag(a[n]);    \
          volatile F t1 = ar * br, t2 = ai * bi, t3 = ar * bi, t4 = ai * br;                         \
          aj[n] = (F)(t1 - t2) + (F)(t3 + t4) * I;                                                   \
        }                                                                                            \
        out[j] = x;                                                                                  \
      }                                                                                              \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* calculate_scale_factor, nufft_util.cc:43-62. */                                                 \
  double oracle_scale_factor_##SFX(int rank, double beta_, double c_) {                              \
    F beta = (F)beta_, c = (F)c_;                                                                    \
    int n = 100;                                                                                     \
    F h = (F)(2.0 / n), x = (F)-1.0, sum = (F)0.0;                                                   \
    for (int i = 1; i < n; i++) { x += h; sum = (F)(sum + exp(beta * sqrt(1.0 - x * x))); }          \
    sum = (F)(sum + 1.0);                                                                            \
    sum *= h;                                                                                        \
    sum = (F)(sum * sqrt(1.0 / c));                                                                  \
    F scale = sum;                                                                                   \
    if (rank > 1) scale *= sum;                                                                      \
    if (rank > 2) scale *= sum;                                                                      \
    return (double)(F)(1.0 / scale);                                                                 \
  }                                                                                                  \
                                                                                                    \
  static inline long wrapl_##SFX(long g, long n) { g %= n; return g < 0 ? g + n : g; }               \
                                                                                                    \
  /* Spreader: spread_subproblem_{1,2,3}d + add_wrapped_subgrid, nufft_plan.cc:1463-1682, folded     \
   * into one loop over the sorted points with per-cell wrap (same sums, point order = idx).       \
   * pts [rank][M] folded; c [M] complex interleaved; fw [nf0*nf1*nf2] complex, zeroed here. */      \
  void oracle_spread_##SFX(int rank, long M, const F* pts, const int* idx, const F* c, const int* nf, \
                           int ns, double beta_, double c_par, F* fw) {                              \
    F beta = (F)beta_, cc = (F)c_par, hw = (F)ns / 2;                                                \
    long n1 = nf[0], n2 = rank > 1 ? nf[1] : 1, n3 = rank > 2 ? nf[2] : 1;                           \
    memset(fw, 0, sizeof(F) * 2 * n1 * n2 * n3);                                                     \
    F k1[ORACLE_MAX_W], k2[ORACLE_MAX_W], k3[ORACLE_MAX_W], kv[2 * ORACLE_MAX_W];                    \
    k2[0] = k3[0] = (F)1;                                                                            \
    for (long jj = 0; jj < M; ++jj) {                                                                \
      long j = idx ? idx[jj] : jj;                                                                   \
      F re0 = c[2 * j], im0 = c[2 * j + 1];                                                          \
      long i1 = (long)CEILF(pts[j] - hw), i2 = 0, i3 = 0;                                            \
      es_vector_##SFX(k1, (F)i1 - pts[j], ns, beta, cc, hw);                                         \
      if (rank > 1) { F y = pts[M + j]; i2 = (long)CEILF(y - hw); es_vector_##SFX(k2, (F)i2 - y, ns, beta, cc, hw); } \
      if (rank > 2) { F z = pts[2 * M + j]; i3 = (long)CEILF(z - hw); es_vector_##SFX(k3, (F)i3 - z, ns, beta, cc, hw); } \
      for (int i = 0; i < ns; ++i) { kv[2 * i] = re0 * k1[i]; kv[2 * i + 1] = im0 * k1[i]; }         \
      for (int dz = 0; dz < (rank > 2 ? ns : 1); ++dz) {                                             \
        long oz = rank > 2 ? wrapl_##SFX(i3 + dz, n3) * n1 * n2 : 0;                                 \
        for (int dy = 0; dy < (rank > 1 ? ns : 1); ++dy) {                                           \
          long oy = oz + (rank > 1 ? wrapl_##SFX(i2 + dy, n2) * n1 : 0);                             \
          F kerval = rank > 2 ? k2[dy] * k3[dz] : (rank > 1 ? k2[dy] : (F)1);                        \
          for (int dx = 0; dx < ns; ++dx) {                                                          \
            long o = oy + wrapl_##SFX(i1 + dx, n1);                                                  \
            if (rank == 1) { fw[2 * o] += kv[2 * dx]; fw[2 * o + 1] += kv[2 * dx + 1]; }             \
            else { fw[2 * o] += kerval * kv[2 * dx]; fw[2 * o + 1] += kerval * kv[2 * dx + 1]; }     \
          }                                                                                          \
        }                                                                                            \
      }                                                                                              \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* Interpolator: interpSorted + interp_line/square/cube, nufft_plan.cc:1136-1461. */               \
  void oracle_interp_##SFX(int rank, long M, const F* pts, const F* fw, const int* nf, int ns,       \
                           double beta_, double c_par, F* c) {                                       \
    F beta = (F)beta_, cc = (F)c_par, hw = (F)ns / 2;                                                \
    long n1 = nf[0], n2 = rank > 1 ? nf[1] : 1, n3 = rank > 2 ? nf[2] : 1;                           \
    _Pragma("omp parallel for schedule(static)")                                                     \
    for (long j = 0; j < M; ++j) {                                                                   \
      F k1[ORACLE_MAX_W], k2[ORACLE_MAX_W], k3[ORACLE_MAX_W];                                        \
      long i1 = (long)CEILF(pts[j] - hw), i2 = 0, i3 = 0;                                            \
      es_vector_##SFX(k1, (F)i1 - pts[j], ns, beta, cc, hw);                                         \
      if (rank > 1) { F y = pts[M + j]; i2 = (long)CEILF(y - hw); es_vector_##SFX(k2, (F)i2 - y, ns, beta, cc, hw); } \
      if (rank > 2) { F z = pts[2 * M + j]; i3 = (long)CEILF(z - hw); es_vector_##SFX(k3, (F)i3 - z, ns, beta, cc, hw); } \
      F o0 = 0, o1 = 0;                                                                              \
      for (int dz = 0; dz < (rank > 2 ? ns : 1); ++dz) {                                             \
        long oz = rank > 2 ? wrapl_##SFX(i3 + dz, n3) * n1 * n2 : 0;                                 \
        for (int dy = 0; dy < (rank > 1 ? ns : 1); ++dy) {                                           \
          long oy = oz + (rank > 1 ? wrapl_##SFX(i2 + dy, n2) * n1 : 0);                             \
          F k23 = rank > 2 ? k2[dy] * k3[dz] : (rank > 1 ? k2[dy] : (F)1);                           \
          for (int dx = 0; dx < ns; ++dx) {                                                          \
            long o = oy + wrapl_##SFX(i1 + dx, n1);                                                  \
            F k = rank > 1 ? k1[dx] * k23 : k1[dx];                                                  \
            o0 += fw[2 * o] * k;                                                                     \
            o1 += fw[2 * o + 1] * k;                                                                 \
          }                                                                                          \
        }                                                                                            \
      }                                                                                              \
      c[2 * j] = o0;                                                                                 \
      c[2 * j + 1] = o1;                                                                             \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* deconvolve_{1,2,3}d, nufft_plan.cc:729-881, CMCL mode order. dir 1: fk <- fw (type 1);          \
   * dir 2: fw <- fk with zero padding (type 2). n = modes per dim, nf = fine dims. */               \
  void oracle_deconvolve_##SFX(int rank, int dir, const int* n, const int* nf, const F* p1,          \
                               const F* p2, const F* p3, F* fk, F* fw) {                             \
    long n1 = n[0], n2 = rank > 1 ? n[1] : 1, n3 = rank > 2 ? n[2] : 1;                              \
    long f1 = nf[0], f2 = rank > 1 ? nf[1] : 1, f3 = rank > 2 ? nf[2] : 1;                           \
    if (dir == 2) memset(fw, 0, sizeof(F) * 2 * f1 * f2 * f3);                                       \
    for (long i3 = 0; i3 < n3; ++i3) {                                                               \
      long k3 = i3 - n3 / 2, w3 = k3 >= 0 ? k3 : f3 + k3;                                            \
      F pre3 = rank > 2 ? (F)1.0 / p3[labs(k3)] : (F)1.0;                                            \
      for (long i2 = 0; i2 < n2; ++i2) {                                                             \
        long k2 = i2 - n2 / 2, w2 = k2 >= 0 ? k2 : f2 + k2;                                          \
        F pre = rank > 1 ? pre3 / p2[labs(k2)] : pre3;                                               \
        for (long i1 = 0; i1 < n1; ++i1) {                                                           \
          long k1 = i1 - n1 / 2, w1 = k1 >= 0 ? k1 : f1 + k1;                                        \
          long o = i1 + n1 * (i2 + n2 * i3), w = w1 + f1 * (w2 + f2 * (rank > 2 ? w3 : 0));          \
          F d = p1[labs(k1)];                                                                        \
          if (dir == 1) { fk[2 * o] = (pre * fw[2 * w]) / d; fk[2 * o + 1] = (pre * fw[2 * w + 1]) / d; } \
          else { fw[2 * w] = (pre * fk[2 * o]) / d; fw[2 * w + 1] = (pre * fk[2 * o + 1]) / d; }     \
        }                                                                                            \
      }                                                                                              \
    }                                                                                                \
  }                                                                                                  \
                                                                                                    \
  /* Whole transform: Plan::initialize + set_points + execute, nufft_plan.cc:166-351, with the GPU  \
   * plan's parameter choices (nufft_plan.cu.cc:1849-1857, 3121-3204). grid_dims x-fastest; points   \
   * [rank][M] raw radians; c [T][M], f [T][N] complex interleaved. Returns 0, or 1 on bad args.    \
   * info_out (optional, 8 ints): ns, nf0, nf1, nf2. */                                              \
  int oracle_nufft_##SFX(int type, int rank, const int* grid_dims, int fft_sign, int T, double tol,  \
                         int points_range, int num_threads, long M, const F* points, F* c, F* f,     \
                         int* info_out) {                                                            \
    if (rank < 1 || rank > 3 || T < 1) return 1;                                                     \
    double beta, cpar;                                                                               \
    int ns = oracle_kernel_params_##SFX(tol, 2.0, &beta, &cpar);                                     \
    int nf[3] = {1, 1, 1}, n[3] = {1, 1, 1}, bins[3] = {1, 1, 1};                                    \
    long nftot = 1, ntot = 1;                                                                        \
    for (int d = 0; d < rank; ++d) {                                                                 \
      n[d] = grid_dims[d];                                                                           \
      int g = (int)(2.0 * n[d]);                                                                     \
      if (g < 2 * ns) g = 2 * ns;                                                                    \
      nf[d] = oracle_next_smooth_int(g);                                                             \
      nftot *= nf[d];                                                                                \
      ntot *= n[d];                                                                                  \
    }                                                                                                \
    if (rank == 1) bins[0] = 1024;                                                                   \
    else if (rank == 2) { bins[0] = 32; bins[1] = 32; }                                              \
    else { bins[0] = 16; bins[1] = 16; bins[2] = 2; }                                                \
    if (info_out) { info_out[0] = ns; info_out[1] = nf[0]; info_out[2] = nf[1]; info_out[3] = nf[2]; } \
    F* ph[3] = {NULL, NULL, NULL};                                                                   \
    for (int d = 0; d < rank; ++d) {                                                                 \
      ph[d] = (F*)malloc(sizeof(F) * (nf[d] / 2 + 1));                                               \
      oracle_kernel_fseries_##SFX(nf[d], ns, beta, cpar, num_threads, ph[d]);                        \
    }                                                                                                \
    F* folded = (F*)malloc(sizeof(F) * rank * (M > 0 ? M : 1));                                      \
    for (int d = 0; d < rank; ++d)                                                                   \
      oracle_fold_rescale_##SFX(M, points + (long)d * M, folded + (long)d * M, points_range, nf[d]); \
    int nbtot = 1;                                                                                   \
    for (int d = 0; d < rank; ++d) nbtot *= (nf[d] + bins[d] - 1) / bins[d];                         \
    int* idx = (int*)malloc(sizeof(int) * (M > 0 ? M : 1));                                          \
    int* bs = (int*)malloc(sizeof(int) * nbtot);                                                     \
    int* bz = (int*)malloc(sizeof(int) * nbtot);                                                     \
    oracle_binsort_##SFX(rank, M, folded, nf, bins, 0, idx, bs, bz);                                 \
    int fdims[3];                                                                                    \
    for (int d = 0; d < rank; ++d) fdims[d] = nf[rank - 1 - d];                                      \
    _Pragma("omp parallel for schedule(dynamic,1) if (T > 1)")                                       \
    for (int t = 0; t < T; ++t) {                                                                    \
      F* fw = (F*)malloc(sizeof(F) * 2 * nftot);                                                     \
      F* ct = c + 2 * (long)t * M;                                                                   \
      F* ft = f + 2 * (long)t * ntot;                                                                \
      if (type == 1) {                                                                               \
        oracle_spread_##SFX(rank, M, folded, idx, ct, nf, ns, beta, cpar, fw);                       \
        FFT(fw, rank, fdims, 1, nftot, fft_sign, T > 1 ? 1 : num_threads);                           \
        oracle_deconvolve_##SFX(rank, 1, n, nf, ph[0], ph[1], ph[2], ft, fw);                        \
      } else {                                                                                       \
        oracle_deconvolve_##SFX(rank, 2, n, nf, ph[0], ph[1], ph[2], ft, fw);                        \
        FFT(fw, rank, fdims, 1, nftot, fft_sign, T > 1 ? 1 : num_threads);                           \
        oracle_interp_##SFX(rank, M, folded, fw, nf, ns, beta, cpar, ct);                            \
      }                                                                                              \
      free(fw);                                                                                      \
    }                                                                                                \
    free(idx); free(bs); free(bz); free(folded);                                                     \
    for (int d = 0; d < rank; ++d) free(ph[d]);                                                      \
    return 0;                                                                                        \
  }

DEFINE_ORACLE(float, f32, float complex, cexpf, clogf, expf, cosf, sinf, log10f, logf, fmodf, ceilf, floorf, fabsf, 6e-08f, fft235_f32)
DEFINE_ORACLE(double, f64, double complex, cexp, clog, exp, cos, sin, log10, log, fmod, ceil, floor, fabs, 1.1e-16, fft235_f64)
