// nufft_kernels_b200.cc -- TensorFlow-side glue: the body a maintainer drops into
// NUFFTBaseOp<GPUDevice, FloatType>::Execute (tensorflow_nufft/cc/kernels/nufft_kernels.cc:381-542)
// so that the GPU kernels of the NUFFT / Interp / Spread ops call libb200nufft.so through the C ABI
// (include/b200nufft.h) instead of Plan<GPUDevice, FloatType> (nufft_plan.cu.cc).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: TensorFlow headers are absent here. It is kept as
// reviewed source; the same call sequence is exercised (and tested) by the torch-hosted mirror in
// tensorflow_nufft_b200/python/ops/nufft_ops.py. Everything above Execute -- validation, batch
// bookkeeping, allocate_output, source/target transposes (nufft_kernels.cc:54-379) -- stays as is,
// except that the points reverse + transpose (:276-303) is no longer needed: the engine reads the
// op's own [..., M, rank] layout (b200nufft_set_points_interleaved).
#if defined(GOOGLE_CUDA) && defined(B200NUFFT_WITH_TENSORFLOW)

#include "b200nufft.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/platform/stream_executor.h"

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
}  // namespace

// `points` here is the op's reshaped input [calls, M, rank] (outer batch dims first), NOT the
// reversed/transposed copy the stock kernel builds.
template <typename FloatType>
Status ExecuteB200(OpKernelContext* ctx, TransformType type, int rank, FftDirection fft_direction,
                   int num_transforms, FloatType tol, OpType op_type, const Options& proto_options,
                   int64_t batch_rank, const int64_t* source_batch_dims, const int64_t* points_batch_dims,
                   const int64_t* grid_dims /* already x-fastest, nufft_kernels.cc:347-352 */,
                   int64_t num_points, const FloatType* points, void* source, void* target) {
  auto* stream = ctx->op_device_context()->stream();
  if (!stream) return errors::Internal("No GPU stream available.");
  void* cu_stream = *reinterpret_cast<void**>(stream->platform_specific_handle().stream);  // cudaStream_t

  b200nufft_opts opts;
  b200nufft_default_opts(&opts);
  opts.points_range = static_cast<int>(proto_options.points_range());
  opts.check_points_range = proto_options.debugging().check_points_range();
  opts.max_batch_size = proto_options.max_batch_size();
  opts.spread_only = op_type != OpType::NUFFT;
  // options.num_threads = TF intra-op pool size (nufft_kernels.cc:462-465): selects the chunking of
  // the reference-compatible float deconvolution factors.
  opts.num_threads_compat = ctx->device()->tensorflow_cpu_worker_threads()->num_threads;

  int64_t num_coeffs = 1;
  for (int d = 0; d < rank; ++d) num_coeffs *= grid_dims[d];
  int64_t num_calls = 1;
  for (int d = 0; d < batch_rank; ++d) num_calls *= points_batch_dims[d];

  b200nufft_plan* plan = nullptr;
  int rc = b200nufft_plan_create(&plan, type == TransformType::TYPE_1 ? 1 : 2, rank, grid_dims,
                                 static_cast<int>(fft_direction), num_transforms, static_cast<double>(tol),
                                 std::is_same<FloatType, double>::value ? B200NUFFT_COMPLEX128 : B200NUFFT_COMPLEX64,
                                 &opts, ctx->eigen_gpu_device().stream() ? /*device ordinal*/ stream->parent()->device_ordinal() : 0);
  if (rc != B200NUFFT_OK) return FromB200(rc, b200nufft_last_create_error());
  // A production build keeps `plan` in a process-level LRU keyed by the create arguments (the cuFFT
  // plan and all buffers then cost nothing per op call); destroyed here for brevity.
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
  b200nufft_plan_destroy(plan);
  return status;
}

}  // namespace nufft
}  // namespace tensorflow

#endif  // GOOGLE_CUDA && B200NUFFT_WITH_TENSORFLOW
