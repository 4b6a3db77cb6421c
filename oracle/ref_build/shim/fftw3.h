// Declarations of the FFTW3 entry points named by the reference's fftw_api.h. FFTW itself is not
// in this image; fft_shim.cc implements them with a plain mixed-radix (2,3,4,5) Stockham FFT.
// Test infrastructure only. NOTE: FFT results are therefore not FFTW's bits.
#pragma once
#include <cstddef>
typedef float fftwf_complex[2];
typedef double fftw_complex[2];
typedef struct fftwf_plan_s* fftwf_plan;
typedef struct fftw_plan_s* fftw_plan;
#define FFTW_MEASURE 0u
#define FFTW_ESTIMATE 64u
#define FFTW_PATIENT 32u
#define FFTW_EXHAUSTIVE 8u
extern "C" {
int fftwf_init_threads(); int fftw_init_threads();
void fftwf_cleanup_threads(); void fftw_cleanup_threads();
void fftwf_plan_with_nthreads(int); void fftw_plan_with_nthreads(int);
void fftwf_make_planner_thread_safe(); void fftw_make_planner_thread_safe();
float* fftwf_alloc_real(size_t); double* fftw_alloc_real(size_t);
fftwf_complex* fftwf_alloc_complex(size_t); fftw_complex* fftw_alloc_complex(size_t);
void fftwf_free(void*); void fftw_free(void*);
fftwf_plan fftwf_plan_many_dft(int, const int*, int, fftwf_complex*, const int*, int, int,
                               fftwf_complex*, const int*, int, int, int, unsigned);
fftw_plan fftw_plan_many_dft(int, const int*, int, fftw_complex*, const int*, int, int,
                             fftw_complex*, const int*, int, int, int, unsigned);
void fftwf_execute(fftwf_plan); void fftw_execute(fftw_plan);
void fftwf_destroy_plan(fftwf_plan); void fftw_destroy_plan(fftw_plan);
}
