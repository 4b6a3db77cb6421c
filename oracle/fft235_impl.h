/* fft235_impl.h -- body of oracle/fft235.c, included once per precision with
 *   T      real type,  SUF  name suffix,  VL  lines transformed together (one 64-byte vector).
 * TEST INFRASTRUCTURE ONLY (see fft235.c).
 *
 * Scheme: VL lines are transformed at once in a split layout re[k * VL + l], im[k * VL + l]
 * (k = element, l = line), so every butterfly of the Stockham autosort passes is a plain loop over
 * l that the compiler turns into full-width SIMD, whatever the pass stride. Radix 4, 2, 3 and 5
 * butterflies are written out; any other prime factor takes a generic O(p^2) butterfly. The
 * arithmetic order does not depend on the instruction set the loops are compiled for. */

typedef struct { T re, im; } CPLX(SUF);

typedef T FN(vec, SUF) __attribute__((vector_size(VL * sizeof(T)), aligned(64)));
#define VT FN(vec, SUF)
#define LD(base_, k_) (*(const VT*)((base_) + (size_t)(k_) * VL))
#define ST(base_, k_) (*(VT*)((base_) + (size_t)(k_) * VL))

/* One Stockham pass of radix r from (xr, xi) to (yr, yi): n_cur = current sub-length, s = stride.
 * Element k of the working array is the VL-lane row starting at k * VL (one 64-byte vector; GCC
 * vector extensions, lowered to whatever the clone's instruction set offers).
 * w[j] = exp(sign*2*pi*i*j/n). */
static inline __attribute__((always_inline)) void FN(pass, SUF)(
    int n, int n_cur, int s, int r, int sign, const T* xr, const T* xi, T* yr, T* yi, const CPLX(SUF)* w) {
  const int m = n_cur / r;
  const int wstep = n / n_cur;
  const T sg = (T)sign;
  if (r == 4) {
    for (int p = 0; p < m; ++p) {
      const CPLX(SUF) w1 = w[p * wstep], w2 = w[2 * p * wstep], w3 = w[3 * p * wstep];
      for (int q = 0; q < s; ++q) {
        const int i0 = q + s * p, i1 = i0 + s * m, i2 = i1 + s * m, i3 = i2 + s * m;
        const int o0 = q + s * 4 * p;
        const VT a0r = LD(xr, i0), a0i = LD(xi, i0), a1r = LD(xr, i1), a1i = LD(xi, i1);
        const VT a2r = LD(xr, i2), a2i = LD(xi, i2), a3r = LD(xr, i3), a3i = LD(xi, i3);
        const VT b0r = a0r + a2r, b0i = a0i + a2i;
        const VT b1r = a0r - a2r, b1i = a0i - a2i;
        const VT b2r = a1r + a3r, b2i = a1i + a3i;
        const VT dr = a1r - a3r, di = a1i - a3i;       /* (a1 - a3) * (sign * i) */
        const VT b3r = -sg * di, b3i = sg * dr;
        const VT c1r = b1r + b3r, c1i = b1i + b3i;
        const VT c2r = b0r - b2r, c2i = b0i - b2i;
        const VT c3r = b1r - b3r, c3i = b1i - b3i;
        ST(yr, o0) = b0r + b2r;                      ST(yi, o0) = b0i + b2i;
        if (p == 0) {
          ST(yr, o0 + s) = c1r;     ST(yi, o0 + s) = c1i;
          ST(yr, o0 + 2 * s) = c2r; ST(yi, o0 + 2 * s) = c2i;
          ST(yr, o0 + 3 * s) = c3r; ST(yi, o0 + 3 * s) = c3i;
        } else {
          ST(yr, o0 + s) = c1r * w1.re - c1i * w1.im;     ST(yi, o0 + s) = c1r * w1.im + c1i * w1.re;
          ST(yr, o0 + 2 * s) = c2r * w2.re - c2i * w2.im; ST(yi, o0 + 2 * s) = c2r * w2.im + c2i * w2.re;
          ST(yr, o0 + 3 * s) = c3r * w3.re - c3i * w3.im; ST(yi, o0 + 3 * s) = c3r * w3.im + c3i * w3.re;
        }
      }
    }
  } else if (r == 2) {
    for (int p = 0; p < m; ++p) {
      const CPLX(SUF) w1 = w[p * wstep];
      for (int q = 0; q < s; ++q) {
        const int i0 = q + s * p, i1 = i0 + s * m, o0 = q + s * 2 * p;
        const VT a0r = LD(xr, i0), a0i = LD(xi, i0), a1r = LD(xr, i1), a1i = LD(xi, i1);
        const VT dr = a0r - a1r, di = a0i - a1i;
        ST(yr, o0) = a0r + a1r; ST(yi, o0) = a0i + a1i;
        ST(yr, o0 + s) = dr * w1.re - di * w1.im; ST(yi, o0 + s) = dr * w1.im + di * w1.re;
      }
    }
  } else if (r == 3) {
    const T hs = sg * (T)0.86602540378443864676;   /* sign * sin(2 pi / 3) */
    for (int p = 0; p < m; ++p) {
      const CPLX(SUF) w1 = w[p * wstep], w2 = w[2 * p * wstep];
      for (int q = 0; q < s; ++q) {
        const int i0 = q + s * p, i1 = i0 + s * m, i2 = i1 + s * m, o0 = q + s * 3 * p;
        const VT a0r = LD(xr, i0), a0i = LD(xi, i0), a1r = LD(xr, i1), a1i = LD(xi, i1);
        const VT a2r = LD(xr, i2), a2i = LD(xi, i2);
        const VT t1r = a1r + a2r, t1i = a1i + a2i;
        const VT t2r = a0r - (T)0.5 * t1r, t2i = a0i - (T)0.5 * t1i;
        const VT t3r = hs * (a1r - a2r), t3i = hs * (a1i - a2i);
        const VT c1r = t2r - t3i, c1i = t2i + t3r;     /* X1 = t2 + i t3, X2 = t2 - i t3 */
        const VT c2r = t2r + t3i, c2i = t2i - t3r;
        ST(yr, o0) = a0r + t1r; ST(yi, o0) = a0i + t1i;
        ST(yr, o0 + s) = c1r * w1.re - c1i * w1.im;     ST(yi, o0 + s) = c1r * w1.im + c1i * w1.re;
        ST(yr, o0 + 2 * s) = c2r * w2.re - c2i * w2.im; ST(yi, o0 + 2 * s) = c2r * w2.im + c2i * w2.re;
      }
    }
  } else if (r == 5) {
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
    const T s1 = sg * (T)0.95105651629515357212, s2 = sg * (T)0.58778525229247312917;
    for (int p = 0; p < m; ++p) {
      const CPLX(SUF) w1 = w[p * wstep], w2 = w[2 * p * wstep], w3 = w[3 * p * wstep], w4 = w[4 * p * wstep];
      for (int q = 0; q < s; ++q) {
        const int i0 = q + s * p, sm = s * m, o0 = q + s * 5 * p;
        const VT a0r = LD(xr, i0), a0i = LD(xi, i0);
        const VT a1r = LD(xr, i0 + sm), a1i = LD(xi, i0 + sm), a4r = LD(xr, i0 + 4 * sm), a4i = LD(xi, i0 + 4 * sm);
        const VT a2r = LD(xr, i0 + 2 * sm), a2i = LD(xi, i0 + 2 * sm), a3r = LD(xr, i0 + 3 * sm), a3i = LD(xi, i0 + 3 * sm);
        const VT p14r = a1r + a4r, p14i = a1i + a4i, m14r = a1r - a4r, m14i = a1i - a4i;
        const VT p23r = a2r + a3r, p23i = a2i + a3i, m23r = a2r - a3r, m23i = a2i - a3i;
        const VT e1r = a0r + c1 * p14r + c2 * p23r, e1i = a0i + c1 * p14i + c2 * p23i;
        const VT e2r = a0r + c2 * p14r + c1 * p23r, e2i = a0i + c2 * p14i + c1 * p23i;
        /* o1 = s1 m14 + s2 m23, o2 = s2 m14 - s1 m23; X1 = e1 + i o1, X4 = e1 - i o1, X2 = e2 + i o2, X3 = e2 - i o2 */
        const VT o1r = s1 * m14r + s2 * m23r, o1i = s1 * m14i + s2 * m23i;
        const VT o2r = s2 * m14r - s1 * m23r, o2i = s2 * m14i - s1 * m23i;
        const VT x1r = e1r - o1i, x1i = e1i + o1r, x4r = e1r + o1i, x4i = e1i - o1r;
        const VT x2r = e2r - o2i, x2i = e2i + o2r, x3r = e2r + o2i, x3i = e2i - o2r;
        ST(yr, o0) = a0r + p14r + p23r; ST(yi, o0) = a0i + p14i + p23i;
        ST(yr, o0 + s) = x1r * w1.re - x1i * w1.im;     ST(yi, o0 + s) = x1r * w1.im + x1i * w1.re;
        ST(yr, o0 + 2 * s) = x2r * w2.re - x2i * w2.im; ST(yi, o0 + 2 * s) = x2r * w2.im + x2i * w2.re;
        ST(yr, o0 + 3 * s) = x3r * w3.re - x3i * w3.im; ST(yi, o0 + 3 * s) = x3r * w3.im + x3i * w3.re;
        ST(yr, o0 + 4 * s) = x4r * w4.re - x4i * w4.im; ST(yi, o0 + 4 * s) = x4r * w4.im + x4i * w4.re;
      }
    }
  } else {
    const int rstep = n / r;
    for (int p = 0; p < m; ++p) {
      for (int q = 0; q < s; ++q) {
        for (int u = 0; u < r; ++u) {
          VT br = {0}, bi = {0};
          for (int t = 0; t < r; ++t) {
            const CPLX(SUF) o = w[((long)t * u % r) * rstep];
            const VT atr = LD(xr, q + s * (p + t * m)), ati = LD(xi, q + s * (p + t * m));
            br += atr * o.re - ati * o.im;
            bi += atr * o.im + ati * o.re;
          }
          const CPLX(SUF) tw = w[((long)p * u % n_cur) * wstep];
          ST(yr, q + s * (r * p + u)) = br * tw.re - bi * tw.im;
          ST(yi, q + s * (r * p + u)) = br * tw.im + bi * tw.re;
        }
      }
    }
  }
}

/* Transforms NG groups of VL adjacent lines. Line (g, l), element k lives at
 * base[k * st + (g * VL + l) * lst]  (st = element stride, lst = line stride; one of them is 1).
 * nl = number of valid lines (<= NG * VL); missing lines are zero-filled and not stored.
 * wk = 4 * n * VL reals of scratch. */
FFT235_CLONES
static void FN(block, SUF)(CPLX(SUF)* restrict base, long st, long lst, int n, int nl, int ng, int sign,
                           const CPLX(SUF)* restrict w, const int* fac, int nfac, T* restrict wk) {
  T* restrict b0r = wk;
  T* restrict b0i = wk + (size_t)n * VL;
  T* restrict b1r = wk + (size_t)2 * n * VL;
  T* restrict b1i = wk + (size_t)3 * n * VL;
  for (int g = 0; g < ng; ++g) {
    const int l0 = g * VL;
    const int cnt = nl - l0 < VL ? nl - l0 : VL;
    if (cnt <= 0) break;
    /* gather */
    if (lst == 1) {
      if (cnt == VL) {
        for (int k = 0; k < n; ++k) {
          const CPLX(SUF)* src = base + (size_t)k * st + l0;
          for (int l = 0; l < VL; ++l) { b0r[(size_t)k * VL + l] = src[l].re; b0i[(size_t)k * VL + l] = src[l].im; }
        }
      } else {
        for (int k = 0; k < n; ++k) {
          const CPLX(SUF)* src = base + (size_t)k * st + l0;
          for (int l = 0; l < VL; ++l) {
            b0r[(size_t)k * VL + l] = l < cnt ? src[l].re : (T)0;
            b0i[(size_t)k * VL + l] = l < cnt ? src[l].im : (T)0;
          }
        }
      }
    } else {   /* contiguous lines (st == 1), lines lst apart: transpose in */
      for (int l = 0; l < VL; ++l) {
        if (l < cnt) {
          const CPLX(SUF)* src = base + (size_t)(l0 + l) * lst;
          for (int k = 0; k < n; ++k) { b0r[(size_t)k * VL + l] = src[k].re; b0i[(size_t)k * VL + l] = src[k].im; }
        } else {
          for (int k = 0; k < n; ++k) { b0r[(size_t)k * VL + l] = 0; b0i[(size_t)k * VL + l] = 0; }
        }
      }
    }
    /* passes */
    T *xr = b0r, *xi = b0i, *yr = b1r, *yi = b1i;
    int n_cur = n, s = 1;
    for (int i = 0; i < nfac; ++i) {
      FN(pass, SUF)(n, n_cur, s, fac[i], sign, xr, xi, yr, yi, w);
      n_cur /= fac[i];
      s *= fac[i];
      T* t = xr; xr = yr; yr = t;
      t = xi; xi = yi; yi = t;
    }
    /* scatter */
    if (lst == 1) {
      if (cnt == VL) {
        for (int k = 0; k < n; ++k) {
          CPLX(SUF)* dst = base + (size_t)k * st + l0;
          for (int l = 0; l < VL; ++l) { dst[l].re = xr[(size_t)k * VL + l]; dst[l].im = xi[(size_t)k * VL + l]; }
        }
      } else {
        for (int k = 0; k < n; ++k) {
          CPLX(SUF)* dst = base + (size_t)k * st + l0;
          for (int l = 0; l < cnt; ++l) { dst[l].re = xr[(size_t)k * VL + l]; dst[l].im = xi[(size_t)k * VL + l]; }
        }
      }
    } else {
      for (int l = 0; l < cnt; ++l) {
        CPLX(SUF)* dst = base + (size_t)(l0 + l) * lst;
        for (int k = 0; k < n; ++k) { dst[k].re = xr[(size_t)k * VL + l]; dst[k].im = xi[(size_t)k * VL + l]; }
      }
    }
  }
}

void FN(fft235, SUF)(T* data_, int rank, const int* dims, int howmany, long dist, int sign, int nthreads) {
  CPLX(SUF)* data = (CPLX(SUF)*)data_;
  int n3[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) n3[3 - rank + i] = dims[i];
  const long tot = (long)n3[0] * n3[1] * n3[2];
  const long stride3[3] = {(long)n3[1] * n3[2], (long)n3[2], 1};
  if (nthreads < 1) nthreads = 1;
  for (int ax = 2; ax >= 0; --ax) {
    const int n = n3[ax];
    if (n == 1) continue;
    const long st = stride3[ax];
    int fac[64];
    const int nfac = fft235_factorize(n, fac);
    CPLX(SUF)* w = (CPLX(SUF)*)malloc(sizeof(CPLX(SUF)) * n);
    for (int k = 0; k < n; ++k) {
      const double ang = sign * 2.0 * M_PI * (double)k / (double)n;
      w[k].re = (T)cos(ang);
      w[k].im = (T)sin(ang);
    }
    /* Lines of this axis: (outer, inner), element k at outer*st*n + inner + k*st, inner in [0, st).
     * st == 1: lines are contiguous, consecutive `outer` are n apart -> blocks of NG*VL lines over
     * outer (all batches and outer indices form one run when dist == tot).
     * st  > 1: lines with adjacent `inner` are adjacent in memory -> blocks of NG*VL over inner. */
    const int NG = 4;
    const int BW = NG * VL;
    long nblocks, per_outer = 1, n_outer = tot / (st * n);
    long hm = howmany;
    if (st == 1) {
      if (dist == tot) { n_outer *= howmany; hm = 1; }   /* back-to-back batches: one run of lines */
      per_outer = (n_outer + BW - 1) / BW;      /* blocks per batch entry */
      nblocks = hm * per_outer;
    } else {
      per_outer = (st + BW - 1) / BW;           /* blocks per (batch, outer) */
      nblocks = (long)howmany * n_outer * per_outer;
    }
#pragma omp parallel num_threads(nthreads)
    {
      T* wk = (T*)aligned_alloc(64, sizeof(T) * 4 * (size_t)n * VL);
#pragma omp for schedule(static)
      for (long blk = 0; blk < nblocks; ++blk) {
        if (st == 1) {
          const long bi = blk % per_outer, batch = blk / per_outer;
          const long o0 = bi * BW;
          const int nl = (int)(n_outer - o0 < BW ? n_outer - o0 : BW);
          FN(block, SUF)(data + batch * dist + o0 * n, 1, n, n, nl, NG, sign, w, fac, nfac, wk);
        } else {
          const long bi = blk % per_outer;
          const long rest = blk / per_outer;
          const long outer = rest % n_outer, batch = rest / n_outer;
          const long i0 = bi * BW;
          const int nl = (int)(st - i0 < BW ? st - i0 : BW);
          FN(block, SUF)(data + batch * dist + outer * st * n + i0, st, 1, n, nl, NG, sign, w, fac, nfac, wk);
        }
      }
      free(wk);
    }
    free(w);
  }
}
#undef VT
#undef LD
#undef ST
