// host_params.h -- host-side plan parameters of the B200 NUFFT engine: kernel width / beta from
// the tolerance, fine-grid sizes, Gauss-Legendre nodes and the deconvolution factors.
// Behavioural references (re-stated, not copied): nufft_plan.cu.cc:3040-3099 (setup_spreader),
// :3166-3204 (set_grid_size), nufft_util.cc:43-133, nufft_plan.h:739-780.
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

constexpr int kMaxKernelWidth = 16;   // nufft_plan.h:68

struct KernelParams {
  int ns = 0;             // kernel width in fine-grid cells
  double beta = 0;        // ES exponent scale, value held exactly as FloatType
  double c = 0;           // 4/ns^2, value held exactly as FloatType
  double half_width = 0;  // ns/2 as FloatType
  double sigma = 2.0;     // upsampling factor
};

// Width from tolerance, in FloatType arithmetic with the host libm (glibc log10f/log10).
template <typename F> int kernel_width_from_tol(F tol, double sigma);
// Fills beta / c / half_width in FloatType arithmetic.
template <typename F> KernelParams kernel_params_from_width(int ns, double sigma);
template <typename F> KernelParams make_kernel_params(F tol, double sigma);

int next_smooth_int(int n);
// Fine-grid size along one dim; returns false if spread_only and the size is not admissible.
bool fine_grid_size(int64_t n_modes, double sigma, int ns, bool spread_only, int* nf);

// n-point Gauss-Legendre nodes (ascending) and weights on [-1, 1]; weights normalised to sum 2.
void gauss_legendre(int n, double* x, double* w);

// ES kernel at x in FloatType-with-double-intermediates arithmetic (nufft_util.cc:64-69).
template <typename F> F es_kernel_host(F x, const KernelParams& kp);

// Deconvolution factors phi_hat[0..nf/2]. mode 0: reference-compatible (FloatType phase winding,
// restarted at each of min(nout, num_threads) chunk boundaries); mode 1: double precision.
template <typename F>
void kernel_fseries(int nf, const KernelParams& kp, int mode, int num_threads, F* out);

// Spread-only output scale, 1 / (integral of phi)^rank by a 100-point trapezoid rule.
template <typename F> F kernel_scale_factor(int rank, const KernelParams& kp);

}  // namespace b200
