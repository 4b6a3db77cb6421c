// ref_driver.cc -- flat extern "C" driver around the reference's own CPU plan,
// tensorflow::nufft::Plan<CPUDevice, F> (nufft_plan.h:367-508, nufft_plan.cc), compiled from
// /root/reference where it lies. TEST INFRASTRUCTURE ONLY (oracle/_ref/libref.so): the checker
// for the CUDA engine and the timed CPU baseline; never on the product path.
//
// The driver plays the role of NUFFTBaseOp::Execute (nufft_kernels.cc:381-542): it fills
// InternalOptions the same way (:448-465) and calls initialize / set_points / execute.
// Two modes matter:
//   ref_cpu_auto      upsampfac=0, kerevalmeth=0 (AUTO -> Horner, auto sigma): what tfft.nufft does on /cpu:0
//   ref_cpu_gpuparams upsampfac=2, kerevalmeth=1 (DIRECT): the CPU code driven with the parameter
//                     choices of Plan<GPUDevice> (nufft_plan.cu.cc:1849-1857) -- parity target.
#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
// binsort_indices_ and do_binsort_ are private (nufft_plan.h:486-491); this TU only reads them.
#define private public
#define protected public
#include "tensorflow_nufft/cc/kernels/nufft_plan.h"
#undef private
#undef protected
#include "tensorflow_nufft/cc/kernels/nufft_util.h"

using namespace tensorflow;
using namespace tensorflow::nufft;

namespace {
struct Handle {
  int is_double;
  OpKernelContext ctx;
  std::unique_ptr<Plan<CPUDevice, float>> pf;
  std::unique_ptr<Plan<CPUDevice, double>> pd;
};
int fail(const Status& s, char* err, int errlen) {
  if (err && errlen > 0) { std::snprintf(err, errlen, "%s", s.message().c_str()); }
  return 1;
}
template <typename F>
Status init(Plan<CPUDevice, F>* p, int type, int rank, const int* grid_dims, int fft_sign, int ntransf,
            double tol, int points_range, int check_range, int max_batch, double upsampfac,
            int kerevalmeth, int num_threads, int spread_only) {
  InternalOptions opt;
  opt.mutable_debugging()->set_check_points_range(check_range != 0);
  opt.set_max_batch_size(max_batch);
  opt.set_points_range(static_cast<PointsRange>(points_range));
  opt.upsampling_factor = upsampfac;
  opt.kernel_evaluation_method = static_cast<KernelEvaluationMethod>(kerevalmeth);
  opt.num_threads = num_threads;
  if (spread_only) { opt.spread_only = true; opt.upsampling_factor = 2.0; }  // nufft_kernels.cc:457-460
  int dims[3] = {1, 1, 1};
  for (int d = 0; d < rank; ++d) dims[d] = grid_dims[d];
  return p->initialize(type == 1 ? TransformType::TYPE_1 : TransformType::TYPE_2, rank, dims,
                       fft_sign < 0 ? FftDirection::FORWARD : FftDirection::BACKWARD, ntransf,
                       static_cast<F>(tol), opt);
}
}  // namespace

extern "C" {

void* ref_plan_create(int is_double, int type, int rank, const int* grid_dims, int fft_sign, int ntransf,
                      double tol, int points_range, int check_range, int max_batch, double upsampfac,
                      int kerevalmeth, int num_threads, int spread_only, char* err, int errlen) {
  Handle* h = new Handle();
  h->is_double = is_double;
  Status s;
  if (is_double) {
    h->pd.reset(new Plan<CPUDevice, double>(&h->ctx));
    s = init(h->pd.get(), type, rank, grid_dims, fft_sign, ntransf, tol, points_range, check_range,
             max_batch, upsampfac, kerevalmeth, num_threads, spread_only);
  } else {
    h->pf.reset(new Plan<CPUDevice, float>(&h->ctx));
    s = init(h->pf.get(), type, rank, grid_dims, fft_sign, ntransf, tol, points_range, check_range,
             max_batch, upsampfac, kerevalmeth, num_threads, spread_only);
  }
  if (!s.ok()) { fail(s, err, errlen); delete h; return nullptr; }
  return h;
}

// x/y/z: M reals each, MUTATED in place (folded+rescaled), exactly as the reference does
// (nufft_plan.h:237-239); they must stay alive until the plan is destroyed.
int ref_set_points(void* hv, int M, void* x, void* y, void* z, char* err, int errlen) {
  Handle* h = static_cast<Handle*>(hv);
  Status s = h->is_double
      ? h->pd->set_points(M, static_cast<double*>(x), static_cast<double*>(y), static_cast<double*>(z))
      : h->pf->set_points(M, static_cast<float*>(x), static_cast<float*>(y), static_cast<float*>(z));
  return s.ok() ? 0 : fail(s, err, errlen);
}

// op: 0 execute, 1 interp, 2 spread.
int ref_run(void* hv, int op, void* c, void* f, char* err, int errlen) {
  Handle* h = static_cast<Handle*>(hv);
  Status s;
  if (h->is_double) {
    auto* cc = static_cast<std::complex<double>*>(c); auto* ff = static_cast<std::complex<double>*>(f);
    s = op == 0 ? h->pd->execute(cc, ff) : (op == 1 ? h->pd->interp(cc, ff) : h->pd->spread(cc, ff));
  } else {
    auto* cc = static_cast<std::complex<float>*>(c); auto* ff = static_cast<std::complex<float>*>(f);
    s = op == 0 ? h->pf->execute(cc, ff) : (op == 1 ? h->pf->interp(cc, ff) : h->pf->spread(cc, ff));
  }
  return s.ok() ? 0 : fail(s, err, errlen);
}

// out[0]=kernel width, [1..3]=fine dims, [4]=batch size, [5]=num_threads, [6]=kerevalmeth(0 direct,1 horner)
// dout[0]=beta, [1]=c, [2]=upsampling factor, [3]=kernel_scale (spread-only), [4]=half width
void ref_get_params(void* hv, int* out, double* dout) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->is_double) {
    auto* p = h->pd.get();
    out[0] = p->spread_params_.kernel_width; out[1] = p->fine_dims_[0]; out[2] = p->rank_ > 1 ? p->fine_dims_[1] : 1;
    out[3] = p->rank_ > 2 ? p->fine_dims_[2] : 1; out[4] = p->batch_size_; out[5] = p->options_.num_threads;
    out[6] = p->spread_params_.kerevalmeth;
    dout[0] = p->spread_params_.kernel_beta; dout[1] = p->spread_params_.kernel_c;
    dout[2] = p->options_.upsampling_factor; dout[3] = p->options_.spread_only ? p->spread_params_.kernel_scale : 0.0;
    dout[4] = p->spread_params_.kernel_half_width;
  } else {
    auto* p = h->pf.get();
    out[0] = p->spread_params_.kernel_width; out[1] = p->fine_dims_[0]; out[2] = p->rank_ > 1 ? p->fine_dims_[1] : 1;
    out[3] = p->rank_ > 2 ? p->fine_dims_[2] : 1; out[4] = p->batch_size_; out[5] = p->options_.num_threads;
    out[6] = p->spread_params_.kerevalmeth;
    dout[0] = p->spread_params_.kernel_beta; dout[1] = p->spread_params_.kernel_c;
    dout[2] = p->options_.upsampling_factor; dout[3] = p->options_.spread_only ? p->spread_params_.kernel_scale : 0.0;
    dout[4] = p->spread_params_.kernel_half_width;
  }
}

// Copies the deconvolution factors of dimension `dim` (fine_dims[dim]/2+1 reals of the plan's type).
int ref_get_fseries(void* hv, int dim, void* out) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->is_double) {
    auto* p = h->pd.get(); if (!p->fseries_data_[dim]) return 1;
    std::memcpy(out, p->fseries_data_[dim], sizeof(double) * (p->fine_dims_[dim] / 2 + 1));
  } else {
    auto* p = h->pf.get(); if (!p->fseries_data_[dim]) return 1;
    std::memcpy(out, p->fseries_data_[dim], sizeof(float) * (p->fine_dims_[dim] / 2 + 1));
  }
  return 0;
}

// Copies the CPU bin-sort permutation (binsort_indices_, nufft_plan.h:491): M int32.
int ref_get_sort(void* hv, int M, int* out) {
  Handle* h = static_cast<Handle*>(hv);
  int* src = h->is_double ? h->pd->binsort_indices_.flat<int>().data() : h->pf->binsort_indices_.flat<int>().data();
  std::memcpy(out, src, sizeof(int) * M);
  return h->is_double ? h->pd->do_binsort_ : h->pf->do_binsort_;
}

void ref_plan_destroy(void* hv) { delete static_cast<Handle*>(hv); }

// Stand-alone access to the reference's host maths (nufft_util.cc) for known-answer tests.
void ref_kernel_fseries(int is_double, int nf, int ns, double beta, double c, int num_threads, void* out) {
  if (is_double) {
    SpreadParameters<double> sp; sp.kernel_width = ns; sp.kernel_beta = beta; sp.kernel_c = c;
    sp.kernel_half_width = ns / 2.0; sp.num_threads = num_threads;
    kernel_fseries_1d<double>(nf, sp, static_cast<double*>(out));
  } else {
    SpreadParameters<float> sp; sp.kernel_width = ns; sp.kernel_beta = (float)beta; sp.kernel_c = (float)c;
    sp.kernel_half_width = (float)ns / 2; sp.num_threads = num_threads;
    kernel_fseries_1d<float>(nf, sp, static_cast<float*>(out));
  }
}
int ref_next_smooth_int(int n) { return next_smooth_int<int>(n, 1); }
double ref_scale_factor(int is_double, int rank, int ns, double beta, double c) {
  if (is_double) {
    SpreadParameters<double> sp; sp.kernel_width = ns; sp.kernel_beta = beta; sp.kernel_c = c;
    return calculate_scale_factor<double>(rank, sp);
  }
  SpreadParameters<float> sp; sp.kernel_width = ns; sp.kernel_beta = (float)beta; sp.kernel_c = (float)c;
  return calculate_scale_factor<float>(rank, sp);
}
}  // extern "C"
