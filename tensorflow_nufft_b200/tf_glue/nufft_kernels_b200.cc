// nufft_kernels_b200.cc -- TensorFlow-side glue: the body a maintainer drops into
// NUFFTBaseOp<GPUDevice, FloatType>::Execute (tensorflow_nufft/cc/kernels/nufft_kernels.cc:381-542)
// so that the GPU kernels of the NUFFT / Interp / Spread ops call libb200nufft.so through the C ABI
// (include/b200nufft.h) instead of Plan<GPUDevice, FloatType> (nufft_plan.cu.cc).
//
// TensorFlow is not in this repository's image, so this file cannot be linked here; it IS parsed
// and type-checked on every build (tf_glue/glue_check.cc, run by __graft_entry__.build()) against
// the reference's own headers and the stand-in TF headers under oracle/ref_build/shim/. The same
// call sequence is exercised on the GPU by the torch-hosted mirror
// (tensorflow_nufft_b200/python/ops/nufft_ops.py) and by tests/test_gpu_boundary.py.
//
// Everything above Execute -- validation, batch bookkeeping, allocate_output, source/target
// transposes (nufft_kernels.cc:54-379) -- stays as is, except that the points reverse + transpose
// (:276-303) is no longer needed: the engine reads the op's own [..., M, rank] layout
// (b200nufft_set_points_interleaved).
//
// What changes against the stock Execute (:475-540), which builds and destroys a Plan (cuFFT plan,
// kernel factors, every buffer via allocate_temp) inside each op call:
//   * the plan comes from the library's process-level cache (b200nufft_plan_acquire / _release),
//     keyed by the create arguments: steady-state op calls create nothing;
//   * all device memory is TF's: the plan allocates through the op device's tensorflow::Allocator
//     (BFC) via the b200nufft_allocator callbacks -- long-lived raw allocations, so the cached plan
//     may keep them across op calls -- and never calls cudaMalloc;
//   * opts.reuse_points: a training / reconstruction loop that calls the op with an unchanged
//     trajectory (forward, then the gradient's adjoint and d/dpoints transforms) bin-sorts once.
#if defined(GOOGLE_CUDA) && defined(B200NUFFT_WITH_TENSORFLOW)

#include <type_traits>

#include "b200nufft.h"
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {
namespace nufft {

namespace {
Status FromB200(int rc, const char* msg) {
  switch (rc) {
    case B200NUFFT_OK: return OkStatus();
    case B200NUFFT_INVALID_ARGUMENT: return errors::InvalidArgument(msg);
    case B200NUFFT_UNIMPLEMENTED: return errors::Unimplemented(msg);
    case B200NUFFT_RESOURCE_EXHAUSTED: return errors::ResourceExhausted(msg);
    default: return errors::Internal(msg);
  }
}

// b200nufft_allocator thunks over the op device's allocator (the GPU BFC allocator).
void* TfAlloc(void* user, size_t bytes, int /*device*/) {
  return static_cast<Allocator*>(user)->AllocateRaw(256, bytes);
}
void TfFree(void* user, void* ptr, int /*device*/) {
  static_cast<Allocator*>(user)->DeallocateRaw(ptr);
}
}  // namespace

// `points` here is the op's reshaped input [calls, M, rank] (outer batch dims first), NOT the
// reversed/transposed copy the stock kernel builds.
template <typename FloatType>
Status ExecuteB200(OpKernelContext* ctx, TransformType type, int rank, FftDirection fft_direction,
                   int num_transforms, FloatType tol, OpType op_type, const Options& proto_options,
                   int64_t batch_rank, const int64_t* source_batch_dims, const int64_t* points_batch_dims,
                   const int64_t* grid_dims /* already x-fastest, nufft_kernels.cc:347-352 */,
                   int64_t num_points, const FloatType* points, void* source, void* target) {
  // The stream the reference's own kernels launch on (nufft_plan.cu.cc:2351).
  void* cu_stream = const_cast<void*>(static_cast<const void*>(ctx->eigen_device<GPUDevice>().stream()));
  const int device_ordinal = ctx->device()->tensorflow_accelerator_device_info()->gpu_id;

  b200nufft_opts opts;
  b200nufft_default_opts(&opts);
  opts.points_range = static_cast<int>(proto_options.points_range());
  opts.check_points_range = proto_options.debugging().check_points_range();
  opts.max_batch_size = proto_options.max_batch_size();
  opts.spread_only = op_type != OpType::NUFFT;
  opts.reuse_points = 1;
  // options.num_threads = TF intra-op pool size (nufft_kernels.cc:462-465): selects the chunking of
  // the reference-compatible float deconvolution factors.
  opts.num_threads_compat = ctx->device()->tensorflow_cpu_worker_threads()->num_threads;

  int64_t num_coeffs = 1;
  for (int d = 0; d < rank; ++d) num_coeffs *= grid_dims[d];
  int64_t num_calls = 1;
  for (int d = 0; d < batch_rank; ++d) num_calls *= points_batch_dims[d];

  b200nufft_allocator allocator;
  allocator.alloc = &TfAlloc;
  allocator.free = &TfFree;
  allocator.user = ctx->device()->GetAllocator(AllocatorAttributes());

  b200nufft_plan* plan = nullptr;
  int rc = b200nufft_plan_acquire(&plan, type == TransformType::TYPE_1 ? 1 : 2, rank, grid_dims,
                                  fft_direction == FftDirection::FORWARD ? -1 : 1, num_transforms,
                                  static_cast<double>(tol),
                                  std::is_same<FloatType, double>::value ? B200NUFFT_COMPLEX128 : B200NUFFT_COMPLEX64,
                                  &opts, device_ordinal, &allocator);
  if (rc != B200NUFFT_OK) return FromB200(rc, b200nufft_last_create_error());

  const size_t csize = 2 * sizeof(FloatType);
  Status status = OkStatus();
  for (int64_t call = 0; call < num_calls && status.ok(); ++call) {
    rc = b200nufft_set_points_interleaved(plan, num_points, points + call * num_points * rank, cu_stream);
    if (rc != B200NUFFT_OK) { status = FromB200(rc, b200nufft_last_error(plan)); break; }
    // source index with broadcasting over the outer dims (nufft_kernels.cc:512-523)
    int64_t source_index = 0, rem = call;
    for (int d = 0; d < batch_rank; ++d) {
      int64_t pf = 1, sf = 1;
      for (int j = d + 1; j < batch_rank; ++j) { pf *= points_batch_dims[j]; sf *= source_batch_dims[j]; }
      int64_t i_d = rem / pf;
      rem %= pf;
      if (source_batch_dims[d] == 1) i_d = 0;
      source_index += i_d * sf;
    }
    const int64_t c_elems = num_transforms * num_points, f_elems = num_transforms * num_coeffs;
    char* src = static_cast<char*>(source) + csize * source_index * (type == TransformType::TYPE_1 ? c_elems : f_elems);
    char* dst = static_cast<char*>(target) + csize * call * (type == TransformType::TYPE_1 ? f_elems : c_elems);
    void* c = type == TransformType::TYPE_1 ? src : dst;
    void* f = type == TransformType::TYPE_1 ? dst : src;
    switch (op_type) {
      case OpType::NUFFT: rc = b200nufft_execute(plan, c, f, cu_stream); break;
      case OpType::INTERP: rc = b200nufft_interp(plan, c, f, cu_stream); break;
      case OpType::SPREAD: rc = b200nufft_spread(plan, c, f, cu_stream); break;
    }
    if (rc != B200NUFFT_OK) status = FromB200(rc, b200nufft_last_error(plan));
  }
  b200nufft_plan_release(plan);   // back to the cache: cuFFT plan, buffers and the bin-sort survive the op call
  return status;
}

// Alternative for hosts that insist on per-call temporaries (OpKernelContext::allocate_temp, as the
// reference does, nufft_plan.cu.cc:1981-2013): create the plan with opts.external_workspace = 1 and
//   Tensor ws;  ctx->allocate_temp(DT_INT8, TensorShape({(int64_t)b200nufft_workspace_bytes(plan, M)}), &ws);
//   b200nufft_bind_workspace(plan, ws.flat<int8>().data(), bytes, M);  ... set_points / execute ...
//   b200nufft_unbind_workspace(plan);   // before `ws` dies at the end of Compute
// (no point-set reuse across calls in that mode: the sorted points live in the temporary).

template Status ExecuteB200<float>(OpKernelContext*, TransformType, int, FftDirection, int, float, OpType,
                                   const Options&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                                   int64_t, const float*, void*, void*);
template Status ExecuteB200<double>(OpKernelContext*, TransformType, int, FftDirection, int, double, OpType,
                                    const Options&, int64_t, const int64_t*, const int64_t*, const int64_t*,
                                    int64_t, const double*, void*, void*);

}  // namespace nufft
}  // namespace tensorflow

#endif  // GOOGLE_CUDA && B200NUFFT_WITH_TENSORFLOW
