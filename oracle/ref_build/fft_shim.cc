// fft_shim.cc -- the FFTW3 entry points the reference names (fftw_api.h:27-205), implemented on
// oracle/fft235.c. TEST INFRASTRUCTURE ONLY: lets the unmodified reference CPU plan run in an
// image without FFTW. In-place, contiguous plan_many_dft only (all the reference uses,
// nufft_plan.cc:413-426).
#include <fftw3.h>
#include <chrono>
#include <cstdlib>
extern "C" {
void fft235_f32(float* data, int rank, const int* dims, int howmany, long dist, int sign, int nthreads);
void fft235_f64(double* data, int rank, const int* dims, int howmany, long dist, int sign, int nthreads);
}
namespace {
int g_threads = 1;
double g_fft_seconds = 0.0;   // wall time spent inside fftw[f]_execute since the last reset
struct FftTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~FftTimer() { g_fft_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
struct PlanRec { int rank; int n[3]; int howmany; void* data; long dist; int sign; };
PlanRec* make(int r, const int* n, int h, void* in, int dist, int sign) {
  PlanRec* p = new PlanRec{r, {1, 1, 1}, h, in, dist, sign};
  for (int i = 0; i < r; ++i) p->n[i] = n[i];
  return p;
}
}  // namespace
extern "C" {
int fftwf_init_threads() { return 1; }
int fftw_init_threads() { return 1; }
void fftwf_cleanup_threads() {}
void fftw_cleanup_threads() {}
void fftwf_plan_with_nthreads(int n) { g_threads = n > 0 ? n : 1; }
void fftw_plan_with_nthreads(int n) { g_threads = n > 0 ? n : 1; }
void fftwf_make_planner_thread_safe() {}
void fftw_make_planner_thread_safe() {}
float* fftwf_alloc_real(size_t n) { return static_cast<float*>(std::malloc(sizeof(float) * n)); }
double* fftw_alloc_real(size_t n) { return static_cast<double*>(std::malloc(sizeof(double) * n)); }
fftwf_complex* fftwf_alloc_complex(size_t n) { return static_cast<fftwf_complex*>(std::malloc(sizeof(fftwf_complex) * n)); }
fftw_complex* fftw_alloc_complex(size_t n) { return static_cast<fftw_complex*>(std::malloc(sizeof(fftw_complex) * n)); }
void fftwf_free(void* p) { std::free(p); }
void fftw_free(void* p) { std::free(p); }
fftwf_plan fftwf_plan_many_dft(int r, const int* n, int h, fftwf_complex* in, const int*, int, int dist,
                               fftwf_complex*, const int*, int, int, int sign, unsigned) {
  return reinterpret_cast<fftwf_plan>(make(r, n, h, in, dist, sign));
}
fftw_plan fftw_plan_many_dft(int r, const int* n, int h, fftw_complex* in, const int*, int, int dist,
                             fftw_complex*, const int*, int, int, int sign, unsigned) {
  return reinterpret_cast<fftw_plan>(make(r, n, h, in, dist, sign));
}
void fftwf_execute(fftwf_plan pl) {
  PlanRec* p = reinterpret_cast<PlanRec*>(pl);
  FftTimer timer;
  fft235_f32(static_cast<float*>(p->data), p->rank, p->n, p->howmany, p->dist, p->sign, g_threads);
}
void fftw_execute(fftw_plan pl) {
  PlanRec* p = reinterpret_cast<PlanRec*>(pl);
  FftTimer timer;
  fft235_f64(static_cast<double*>(p->data), p->rank, p->n, p->howmany, p->dist, p->sign, g_threads);
}
// Stage timing for bench.py's cpu_baseline.stages_ms: seconds inside the FFT since the last call
// with reset != 0 (the reference plan itself has no stage timers).
double ref_fft_seconds(int reset) {
  const double v = g_fft_seconds;
  if (reset) g_fft_seconds = 0.0;
  return v;
}
void fftwf_destroy_plan(fftwf_plan p) { delete reinterpret_cast<PlanRec*>(p); }
void fftw_destroy_plan(fftw_plan p) { delete reinterpret_cast<PlanRec*>(p); }
}
