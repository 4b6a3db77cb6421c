// host_params.cc -- see host_params.h. Compiled WITHOUT FMA contraction and without
// -ffast-math: in reference-compatible mode the float results must carry the same roundings as
// the reference library (built with -O3 -march=x86-64, /root/reference/Makefile:38).
#include "host_params.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace b200 {

namespace {
template <typename F> constexpr F kEps();
template <> constexpr float kEps<float>() { return 6e-08f; }      // nufft_plan.h:87
template <> constexpr double kEps<double>() { return 1.1e-16; }   // nufft_plan.h:89
template <typename F> constexpr F kPiF() { return F(3.14159265358979329); }
}  // namespace

template <typename F>
int kernel_width_from_tol(F tol, double sigma) {
  if (tol < kEps<F>()) tol = kEps<F>();
  int ns;
  if (sigma == 2.0) {
    // one digit per power of ten; evaluated in FloatType (float: glibc log10f gives 7 at 1e-6)
    ns = static_cast<int>(std::ceil(-std::log10(tol / F(10.0))));
  } else {
    ns = static_cast<int>(std::ceil(-std::log(tol) / (kPiF<F>() * std::sqrt(1.0 - 1.0 / sigma))));
  }
  ns = std::max(2, ns);
  ns = std::min(ns, kMaxKernelWidth);
  return ns;
}

template <typename F>
KernelParams make_kernel_params(F tol, double sigma) {
  return kernel_params_from_width<F>(kernel_width_from_tol<F>(tol, sigma), sigma);
}

template <typename F>
KernelParams kernel_params_from_width(int ns, double sigma) {
  KernelParams kp;
  kp.sigma = sigma;
  kp.ns = ns;
  F half = static_cast<F>(ns) / 2;
  F c = static_cast<F>(4.0 / static_cast<F>(ns * ns));
  F beta_over_ns = F(2.30);
  if (ns == 2) beta_over_ns = F(2.20);
  if (ns == 3) beta_over_ns = F(2.26);
  if (ns == 4) beta_over_ns = F(2.38);
  if (sigma != 2.0) {
    F gamma = F(0.97);
    beta_over_ns = static_cast<F>(gamma * kPiF<F>() * (1 - 1 / (2 * sigma)));
  }
  F beta = beta_over_ns * static_cast<F>(ns);
  kp.half_width = half;
  kp.c = c;
  kp.beta = beta;
  return kp;
}

int next_smooth_int(int n) {
  if (n <= 2) return 2;
  if (n % 2 == 1) n += 1;
  for (int p = n;; p += 2) {
    int d = p;
    while (d % 2 == 0) d /= 2;
    while (d % 3 == 0) d /= 3;
    while (d % 5 == 0) d /= 5;
    if (d == 1) return p;
  }
}

bool fine_grid_size(int64_t n_modes, double sigma, int ns, bool spread_only, int* nf) {
  int g = spread_only ? static_cast<int>(n_modes) : static_cast<int>(sigma * n_modes);
  if (g < 2 * ns) g = 2 * ns;
  g = next_smooth_int(g);
  *nf = g;
  if (spread_only && g != n_modes) return false;
  return true;
}

void gauss_legendre(int n, double* x, double* w) {
  // Newton iteration on P_n via the three-term recurrence, Chebyshev-like initial guesses.
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double t = std::cos(pi * (i + 0.75) / (n + 0.5));  // approximates the i-th largest root
    double dp = 0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; ++k) {
        double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
      double dt = p1 / dp;
      t -= dt;
      if (std::fabs(dt) < 1e-16 * std::max(1.0, std::fabs(t))) break;
    }
    // re-evaluate the derivative at the converged root
    {
      double p0 = 1.0, p1 = t;
      for (int k = 2; k <= n; ++k) {
        double pk = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
    }
    x[n - 1 - i] = t;
    x[i] = -t;
    double wi = 2.0 / ((1.0 - t) * (1.0 + t) * dp * dp);
    w[n - 1 - i] = wi;
    w[i] = wi;
  }
  if (n % 2 == 1) x[n / 2] = 0.0;
  double s = 0;
  for (int i = 0; i < n; ++i) s += w[i];
  for (int i = 0; i < n; ++i) w[i] = 2.0 * w[i] / s;
}

template <typename F>
F es_kernel_host(F x, const KernelParams& kp) {
  F hw = static_cast<F>(kp.half_width);
  F c = static_cast<F>(kp.c);
  F beta = static_cast<F>(kp.beta);
  if (std::fabs(x) >= hw) return F(0);
  // c*x*x in FloatType, the rest through double, rounded once at the end.
  F cxx = c * x * x;
  return static_cast<F>(std::exp(beta * std::sqrt(1.0 - cxx)));
}

template <typename F>
void kernel_fseries(int nf, const KernelParams& kp, int mode, int num_threads, F* out) {
  const int ns = kp.ns;
  const int nout = nf / 2 + 1;
  if (mode == 1) {
    // Accurate mode: everything in double, closed-form cosines (no phase winding).
    double hw = ns / 2.0;
    int q = static_cast<int>(2 + 3.0 * hw);
    std::vector<double> z(2 * q), w(2 * q), f(q);
    gauss_legendre(2 * q, z.data(), w.data());
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < q; ++n) {
      z[n] *= hw;
      double arg = 1.0 - kp.c * z[n] * z[n];
      f[n] = hw * w[n] * (arg > 0 ? std::exp(kp.beta * std::sqrt(arg)) : 0.0);
    }
    for (int j = 0; j < nout; ++j) {
      double acc = 0;
      for (int n = 0; n < q; ++n)
        acc += f[n] * 2.0 * std::cos(2.0 * pi * j * (nf / 2 - z[n]) / nf);
      out[j] = static_cast<F>(acc);
    }
    return;
  }
  // Reference-compatible mode.
  F hw = static_cast<F>(ns / 2.0);
  int q = static_cast<int>(2 + 3.0 * hw);
  std::vector<double> z(2 * q), w(2 * q);
  gauss_legendre(2 * q, z.data(), w.data());
  std::vector<F> f(q);
  std::vector<std::complex<F>> a(q);
  const std::complex<F> iu(F(0), F(1));
  for (int n = 0; n < q; ++n) {
    z[n] *= hw;
    f[n] = hw * static_cast<F>(w[n]) * es_kernel_host<F>(static_cast<F>(z[n]), kp);
    a[n] = std::exp(F(2) * kPiF<F>() * iu * static_cast<F>(nf / 2 - z[n]) / static_cast<F>(nf));
  }
  int nt = std::min(nout, std::max(1, num_threads));
  std::vector<int> brk(nt + 1);
  for (int t = 0; t <= nt; ++t) brk[t] = static_cast<int>(0.5 + nout * t / static_cast<double>(nt));
  std::vector<std::complex<F>> aj(q);
  for (int t = 0; t < nt; ++t) {
    for (int n = 0; n < q; ++n) aj[n] = std::pow(a[n], static_cast<F>(brk[t]));
    for (int j = brk[t]; j < brk[t + 1]; ++j) {
      F x = F(0);
      for (int n = 0; n < q; ++n) {
        x += f[n] * 2 * aj[n].real();
        // plain (limited-range) complex product, no FMA
        F ar = aj[n].real(), ai = aj[n].imag(), br = a[n].real(), bi = a[n].imag();
        volatile F t1 = ar * br, t2 = ai * bi, t3 = ar * bi, t4 = ai * br;
        aj[n] = std::complex<F>(t1 - t2, t3 + t4);
      }
      out[j] = x;
    }
  }
}

template <typename F>
F kernel_scale_factor(int rank, const KernelParams& kp) {
  F beta = static_cast<F>(kp.beta);
  F c = static_cast<F>(kp.c);
  int n = 100;
  F h = static_cast<F>(2.0 / n);
  F x = F(-1.0);
  F sum = F(0.0);
  for (int i = 1; i < n; i++) {
    x += h;
    sum = static_cast<F>(sum + std::exp(beta * std::sqrt(1.0 - x * x)));
  }
  sum = static_cast<F>(sum + 1.0);
  sum *= h;
  sum = static_cast<F>(sum * std::sqrt(1.0 / c));
  F scale = sum;
  if (rank > 1) scale *= sum;
  if (rank > 2) scale *= sum;
  return static_cast<F>(1.0 / scale);
}

template int kernel_width_from_tol<float>(float, double);
template int kernel_width_from_tol<double>(double, double);
template KernelParams kernel_params_from_width<float>(int, double);
template KernelParams kernel_params_from_width<double>(int, double);
template KernelParams make_kernel_params<float>(float, double);
template KernelParams make_kernel_params<double>(double, double);
template float es_kernel_host<float>(float, const KernelParams&);
template double es_kernel_host<double>(double, const KernelParams&);
template void kernel_fseries<float>(int, const KernelParams&, int, int, float*);
template void kernel_fseries<double>(int, const KernelParams&, int, int, double*);
template float kernel_scale_factor<float>(int, const KernelParams&);
template double kernel_scale_factor<double>(int, const KernelParams&);

}  // namespace b200
