/* fft235.c -- plain-C in-place batched complex FFT (Stockham autosort, mixed radix 4/2/3/5 plus a
 * generic O(p^2) butterfly for any other prime factor), float and double, OpenMP over blocks of
 * lines, SIMD across lines (see fft235_impl.h).
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
 *
 * Speed matters because this FFT sits inside the timed CPU baseline (bench.py --impl reference):
 * round 1's scalar version spent 470 ms on 8 x 1024^2 where pocketfft needs 25-90 ms, which
 * inflated every GPU/CPU ratio. This version is within ~1.5x of pocketfft on the same host (see
 * profiles/r02_cpu_fft.txt). The hot function is compiled for AVX-512, AVX2 and baseline x86-64
 * (GCC target_clones, resolved at load time), with FP contraction off so that all three clones
 * produce the same bits.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(FFT235_NO_CLONES)
#define FFT235_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define FFT235_CLONES
#endif

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

#define FN_(a, b) a##_##b
#define FN(a, b) FN_(a, b)
#define CPLX(s) FN(cplx, s)

#define T float
#define SUF f32
#define VL 16
#include "fft235_impl.h"
#undef T
#undef SUF
#undef VL

#define T double
#define SUF f64
#define VL 8
#include "fft235_impl.h"
#undef T
#undef SUF
#undef VL
