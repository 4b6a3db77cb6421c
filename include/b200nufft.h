/* b200nufft.h -- C ABI of the B200-native NUFFT engine (libb200nufft.so).
 *
 * This is the drop-in boundary for the `tfft.nufft` hot path of mrphys/tensorflow-nufft. It
 * replaces the C++ virtual interface PlanBase<GPUDevice, FloatType>
 *   (reference: tensorflow_nufft/cc/kernels/nufft_plan.h:205-362; GPU implementation
 *    tensorflow_nufft/cc/kernels/nufft_plan.cu.cc:1808-3032),
 * which is instantiated and driven only from NUFFTBaseOp::Execute
 *   (tensorflow_nufft/cc/kernels/nufft_kernels.cc:475-540).
 * Plain pointers and sizes only; no TensorFlow, torch or C++ types. See INTEGRATION.md for the
 * OpKernel-side binding.
 *
 * Layout contract (identical to PlanBase's):
 *   - grid_dims are given fastest-varying axis first ("x-fastest"; the OpKernel reverses TF's
 *     row-major grid shape, nufft_kernels.cc:347-352);
 *   - points are `rank` contiguous device arrays of M reals, coordinate 0 = fastest grid axis,
 *     in radians/sample; they are NOT mutated (the reference mutates them, nufft_plan.h:237-239);
 *   - c is [num_transforms][M] complex, f is [num_transforms][N] complex with modes in CMCL order
 *     (index i <-> frequency i - N/2) and x fastest; both are device pointers;
 *   - all work is enqueued on the caller's stream; no device-wide synchronisation, except that
 *     set_points with check_points_range=1 reads one flag back (as the reference does,
 *     nufft_plan.h:880-898).
 * Memory: by default a plan allocates its buffers with cudaMalloc, growing geometrically when a
 * larger point set arrives. Three ways to take that off the hot path / hand ownership to the host
 * framework (TF's BFC allocator, nufft_plan.cu.cc:1981-2013 uses allocate_temp):
 *   - b200nufft_reserve(plan, M): size every per-point buffer once; set_points with <= M points and
 *     execute then never allocate or free;
 *   - b200nufft_plan_create_ex with a b200nufft_allocator: every device allocation goes through the
 *     caller's callbacks (long-lived raw allocations, e.g. tensorflow::Allocator::AllocateRaw);
 *   - opts.external_workspace = 1 + b200nufft_workspace_bytes / b200nufft_bind_workspace: the caller
 *     allocates ONE block (allocate_temp / torch.empty) per op call and the plan carves its fine
 *     grid and per-point buffers out of it.
 * CUDA graphs: after b200nufft_reserve (or with a bound workspace) and with check_points_range = 0,
 * set_points + execute only enqueue kernels, cuFFT executions and memsets on the caller's stream, so
 * the pair can be captured into a CUDA graph and replayed (tests/test_gpu_boundary.py; the profile
 * option records events and is not capturable).
 * Every entry point saves and restores the calling thread's current CUDA device. A handle may be
 * used from several threads / streams: calls on one handle are serialised by a per-plan lock, and
 * work enqueued on a new stream waits (cudaStreamWaitEvent) for the plan's previous work.
 */
#ifndef B200NUFFT_H_
#define B200NUFFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200nufft_plan b200nufft_plan;

/* Return codes; they map onto the tensorflow::errors the reference returns. */
enum {
  B200NUFFT_OK = 0,
  B200NUFFT_INVALID_ARGUMENT = 1,   /* errors::InvalidArgument  */
  B200NUFFT_UNIMPLEMENTED = 2,      /* errors::Unimplemented    */
  B200NUFFT_RESOURCE_EXHAUSTED = 3, /* errors::ResourceExhausted */
  B200NUFFT_INTERNAL = 4            /* errors::Internal         */
};

enum { B200NUFFT_COMPLEX64 = 0, B200NUFFT_COMPLEX128 = 1 };

/* points_range: proto enum PointsRange (tensorflow_nufft/proto/nufft_options.proto:13-17). */
enum { B200NUFFT_RANGE_STRICT = 0, B200NUFFT_RANGE_EXTENDED = 1, B200NUFFT_RANGE_INFINITE = 2 };

/* Decoded subset of InternalOptions (tensorflow_nufft/cc/kernels/nufft_options.h:92-162) plus
 * engine knobs. Zero-initialise, or call b200nufft_default_opts. */
typedef struct b200nufft_opts {
  int points_range;        /* B200NUFFT_RANGE_*; tfft.nufft sends EXTENDED (nufft_options.py:251) */
  int check_points_range;  /* options.debugging.check_points_range                               */
  int max_batch_size;      /* options.max_batch_size; 0 = min(num_transforms, 32) with the fine-grid
                              batch capped at 4 GiB (the reference: min(num_transforms, 8),
                              nufft_plan.cu.cc:1923-1928); always clamped to 65535                */
  int spread_only;         /* Interp/Spread ops: no oversampling, no FFT, no deconvolution
                              (nufft_kernels.cc:457-460)                                          */
  int fseries_mode;        /* 0 = reference-compatible deconvolution factors: FloatType arithmetic
                              with the reference's per-thread chunked phase winding
                              (nufft_util.cc:71-117); 1 = accurate (double) factors              */
  int num_threads_compat;  /* chunk count the reference would use for those factors =
                              options.num_threads (TF intra-op pool size, nufft_kernels.cc:462-465);
                              0 = std::thread::hardware_concurrency()                             */
  int bin_dims[3];         /* engine bin geometry (cf. InternalOptions::gpu_bin_size); 0 = auto   */
  int max_subproblem_size; /* points per subproblem (cf. gpu_max_subproblem_size = 1024); 0 = auto */
  int spread_method;       /* 0 auto, 1 global-atomic point-driven, 2 shared-memory tiles,
                              3 window-sorted register runs, 4 same with even-row windows, 6 / 7 window
                              swept along x with rotating row accumulators in 2D / 3D (7 streams the
                              tile through an 8-plane ring). 4, 6, 7: type-1 NUFFT plans only; they
                              fall back to 3 / 2 elsewhere                                        */
  int interp_method;       /* 0 auto, 1 point-driven from L2, 2 shared-memory tiles (TMA staged),
                              lanes over one point's stencil, 3 shared-memory tiles, quarter warp
                              per point, 7 (3D type-2 NUFFT plans, opt-in) quarter-warp gather from
                              an 8-plane ring of the tile streamed along z                        */
  int profile;             /* 1: record CUDA events around the stages (b200nufft_get_timings)     */
  int upsampling;          /* fine-grid oversampling sigma: 0 = 2.0 (what Plan<GPUDevice> always uses,
                              nufft_plan.cu.cc:1855-1857); 1 = 1.25 (low-upsampling mode, width and
                              beta per nufft_plan.h:769-771 / nufft_plan.cu.cc:3089-3092); 2 = the
                              reference CPU plan's automatic choice (nufft_plan.h:739-752)         */
  int reuse_points;        /* 1: set_points fingerprints the raw coordinates on the device and, when
                              they equal the previous call's, every set_points kernel exits at once
                              (bin-sort and stencil records are kept). No host synchronisation.  */
  int external_workspace;  /* 1: the plan allocates no fine grid / per-point buffers itself; the
                              caller binds a block with b200nufft_bind_workspace                  */
  int reserved[8];         /* engine A/B switches used by the tests and probes (0 = default):
                              [0] 1: stage interpolator tiles with cp.async instead of TMA
                              [1] coils per CTA of the 2D spreader / interpolator (1, 2, 4, 8, 16)
                              [2] 1: no pre-clear of the fine grid on the plan's internal stream
                              [3] 1: sweep spreaders use scalar FFMA instead of packed FFMA2
                              [4] FFT stage: 0 the engine's own pruned passes when eligible (complex64,
                                  2D / 3D, power-of-two fine sizes 64..1024, x modes a multiple of 32),
                                  else cuFFT; 1: one full cuFFT plan; 2: cuFFT only (3D: the pruned
                                  three-plan scheme) + amplify / deconvolve kernels
                              [5] 1: flush spreader tiles with REDG instead of TMA reduce-add
                              [6] 1: 3D tiles move all their z-planes (no per-subproblem z range)
                              [7] 1: 2D sweep spreader gathers coil-major strengths (no point-major
                                  pre-pass); 2: 3D sweep spreader of single-transform plans evaluates
                                  the stencil weights itself (no stencil records; measured slower) */
} b200nufft_opts;

typedef struct b200nufft_info {
  int kernel_width;        /* ns                                                          */
  double kernel_beta;      /* ES kernel beta, as FloatType                                 */
  double kernel_c;         /* 4/ns^2, as FloatType                                         */
  double upsampling_factor;
  double kernel_scale;     /* spread-only output scale (nufft_util.cc:43-62), else 0        */
  int fine_dims[3];        /* nf per dim (1 for unused)                                    */
  int bin_dims[3];         /* engine bin geometry                                          */
  int num_bins[3];
  int batch_size;
  int num_threads_compat;
  int64_t num_points;
  int64_t subproblem_bound; /* upper bound on subproblem count used for the launch grid     */
  int spread_method;       /* kernel family actually selected (values of opts.spread_method; 5 = row-lane tiles) */
  int interp_method;       /* ... (values of opts.interp_method; 5 = row-lane tiles)       */
  int fft_method;          /* 0 none (spread-only), 1 one cuFFT plan, 2 cuFFT three-plan pruned scheme (3D),
                              3 the engine's own pruned passes with amplify / deconvolve fused */
} b200nufft_info;

void b200nufft_default_opts(b200nufft_opts* opts);

/* Device-memory callbacks (e.g. thunks around tensorflow::Allocator::AllocateRaw/DeallocateRaw or
 * torch's caching allocator). alloc returns a device pointer aligned to >= 256 bytes, or NULL. */
typedef void* (*b200nufft_alloc_fn)(void* user, size_t bytes, int device);
typedef void (*b200nufft_free_fn)(void* user, void* ptr, int device);
typedef struct b200nufft_allocator {
  b200nufft_alloc_fn alloc;
  b200nufft_free_fn free;
  void* user;
} b200nufft_allocator;

/* Replaces Plan<GPUDevice,F>::initialize (nufft_plan.cu.cc:1809-2030).
 * type 1|2; rank 1..3; fft_sign -1 forward / +1 backward (FftDirection, nufft_plan.h:126-129);
 * tol is the value the plan sees, i.e. static_cast<FloatType>(float attr) (nufft_kernels.cc:361);
 * dtype B200NUFFT_COMPLEX64|128; device = CUDA ordinal. On failure *out = NULL and
 * b200nufft_last_create_error() holds the message. */
int b200nufft_plan_create(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims,
                          int fft_sign, int num_transforms, double tol, int dtype,
                          const b200nufft_opts* opts, int device);

/* Same, with every device allocation routed through `allocator` (NULL = cudaMalloc/cudaFree).
 * Replaces the allocate_temp calls of Plan<GPUDevice,F>::initialize (nufft_plan.cu.cc:1981-2013). */
int b200nufft_plan_create_ex(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims,
                             int fft_sign, int num_transforms, double tol, int dtype,
                             const b200nufft_opts* opts, int device,
                             const b200nufft_allocator* allocator);

/* Replaces ~Plan (nufft_plan.cu.cc:2032-2052). */
void b200nufft_plan_destroy(b200nufft_plan* plan);

/* Process-level plan cache (LRU; capacity from the environment variable B200NUFFT_PLAN_CACHE,
 * default 8 idle plans), keyed by every create argument including opts and the allocator. The
 * reference builds and tears down a plan (cuFFT plan, buffers, kernel factors) inside every op call
 * (nufft_kernels.cc:475); with acquire / release the OpKernel keeps them across calls, and with
 * opts.reuse_points = 1 also the bin-sort of an unchanged trajectory. acquire hands out an idle
 * cached plan or creates one; a plan is never handed to two callers at once. release returns it
 * (evicting, i.e. destroying, the least recently used idle plan beyond the capacity). */
int b200nufft_plan_acquire(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims,
                           int fft_sign, int num_transforms, double tol, int dtype,
                           const b200nufft_opts* opts, int device,
                           const b200nufft_allocator* allocator);
void b200nufft_plan_release(b200nufft_plan* plan);
void b200nufft_plan_cache_clear(void);
/* out[0] = acquire calls served from the cache, out[1] = acquire calls that created a plan,
 * out[2] = idle plans held now. */
void b200nufft_plan_cache_stats(int64_t out[3]);

/* Bytes of ONE caller-owned block that holds the fine-grid batch and every per-point buffer of this
 * plan for point sets of up to num_points points (PlanBase's allocate_temp calls,
 * nufft_plan.cu.cc:1981-2013, 2897-3032). */
size_t b200nufft_workspace_bytes(const b200nufft_plan* plan, int64_t num_points);
/* Carves the plan's buffers out of `workspace` (device memory, >= workspace_bytes(plan,
 * num_points), 256-byte aligned). Until unbind, set_points (<= num_points points) and execute
 * neither allocate nor free. Binding invalidates the current point set. */
int b200nufft_bind_workspace(b200nufft_plan* plan, void* workspace, size_t bytes, int64_t num_points);
int b200nufft_unbind_workspace(b200nufft_plan* plan);
/* Sizes the plan's own per-point buffers for up to num_points points now, so that later
 * set_points / execute calls never call the allocator. */
int b200nufft_reserve(b200nufft_plan* plan, int64_t num_points);
/* Number of device allocations / frees issued by this library so far in this process (cudaMalloc or
 * allocator callbacks); lets a test assert that a code path does not allocate. */
void b200nufft_debug_alloc_counts(int64_t* allocs, int64_t* frees);
/* Test hook: out[0] = set_points calls on this plan that were skipped by the device-side
 * fingerprint match (opts.reuse_points), out[1] = calls that did the full work. Synchronises. */
int b200nufft_get_reuse_stats(b200nufft_plan* plan, int64_t out[2]);

/* Replaces Plan<GPUDevice,F>::set_points (nufft_plan.cu.cc:2054-2111): range check (optional),
 * fold+rescale (nufft_plan.h:676-734, 901-948), bin-sort (:2897-2991), subproblem setup
 * (:2993-3032). x, y, z: device arrays of M reals (y/z ignored when rank < 2/3). */
int b200nufft_set_points(b200nufft_plan* plan, int64_t num_points, const void* x, const void* y,
                         const void* z, void* stream);

/* Fused "points prep": same as set_points but reads the op's own layout, points[M][rank]
 * (row-major, TF's last grid axis last), doing the reverse + transpose of
 * nufft_kernels.cc:276-303 on the fly. */
int b200nufft_set_points_interleaved(b200nufft_plan* plan, int64_t num_points, const void* points,
                                     void* stream);

/* Replaces Plan<GPUDevice,F>::execute (nufft_plan.cu.cc:2113-2168).
 * type 1: c [T][M] in, f [T][N] out; type 2: f in, c out. */
int b200nufft_execute(b200nufft_plan* plan, void* c, void* f, void* stream);

/* Replace Plan<GPUDevice,F>::interp / ::spread (nufft_plan.cu.cc:2170-2225); plan must have been
 * created with spread_only = 1; f is then the fine grid itself, [T][prod(grid_dims)]. */
int b200nufft_interp(b200nufft_plan* plan, void* c, const void* f, void* stream);
int b200nufft_spread(b200nufft_plan* plan, const void* c, void* f, void* stream);

/* Parity hooks (device pointers owned by the plan, valid until the next set_points/destroy):
 * the engine's bin-sort in the engine's own geometry. idx[M] = point ids grouped by bin, stable
 * within a bin; bin_start[bins] = exclusive scan of bin_sizes[bins]. */
int b200nufft_get_sort(const b200nufft_plan* plan, const int32_t** idx, const int32_t** bin_start,
                       const int32_t** bin_sizes, int32_t* bin_count);

/* Stand-alone bin-sort of already folded+rescaled coordinates with an explicit geometry, for
 * bit-exact comparison with the reference's sorts. rounding 0: GPU rule, bin = floor(x/bin_dim),
 * clamp into [0, nbins) with nbins = ceil(nf/bin_dim) (CalcBinSizeNoGhost*, nufft_plan.cu.cc:160-231);
 * rounding 1: CPU rule, bin = int(x/bin_dim) with nbins = nf/bin_dim + 1 (binsort_singlethread,
 * nufft_plan.cc:475-531). All pointers are device pointers; idx_out[M], bin_start_out/bin_sizes_out
 * [prod(nbins)]. is_double selects the coordinate type. */
int b200nufft_binsort(int is_double, int rank, int64_t num_points, const void* x, const void* y,
                      const void* z, const int* fine_dims, const int* bin_dims, int rounding,
                      int32_t* idx_out, int32_t* bin_start_out, int32_t* bin_sizes_out,
                      void* stream);

/* Stand-alone fold+rescale (FoldAndRescale functors, nufft_plan.h:676-734): out[i] for one
 * coordinate array; bit-exact FloatType arithmetic. Device pointers. */
int b200nufft_fold_rescale(int is_double, int points_range, int64_t num_points, const void* in,
                           void* out, int fine_dim, void* stream);

/* Test helper for the parity hooks: synchronous device -> host copy of plan-owned arrays
 * (cudaMemcpy). Not used on the hot path. */
int b200nufft_copy_to_host(void* dst_host, const void* src_device, size_t bytes);

int b200nufft_get_info(const b200nufft_plan* plan, b200nufft_info* info);

/* Copies the deconvolution factors of dimension `dim` (fine_dims[dim]/2+1 reals, FloatType) to
 * host memory (kernel_fseries_1d, nufft_util.cc:71-117). */
int b200nufft_get_fseries(const b200nufft_plan* plan, int dim, void* host_out);

/* After an execute with opts.profile = 1: milliseconds of the last execute's stages,
 * out[0] = spread|interp, out[1] = FFT, out[2] = deconvolve|amplify, out[3] = last set_points.
 * Synchronises on the recorded events. */
int b200nufft_get_timings(b200nufft_plan* plan, float out[4]);

/* Number of kernel launches (own kernels + cuFFT execs) issued by this plan so far. */
int64_t b200nufft_launch_count(const b200nufft_plan* plan);

const char* b200nufft_last_error(const b200nufft_plan* plan);
const char* b200nufft_last_create_error(void);

/* Host-side parameter maths, exported for CPU-only known-answer tests (no GPU needed). */
int b200nufft_host_kernel_width(int is_double, double tol, double upsampling_factor);
int b200nufft_host_next_smooth_int(int n);
/* out[fine_dim/2+1]; FloatType = float|double per is_double; mode/num_threads as in opts. */
int b200nufft_host_fseries(int is_double, int fine_dim, int kernel_width, int mode, int num_threads,
                           void* out);
double b200nufft_host_scale_factor(int is_double, int rank, int kernel_width);
/* n-point Gauss-Legendre rule on [-1,1], nodes ascending. */
int b200nufft_host_gauss_legendre(int n, double* nodes, double* weights);

const char* b200nufft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H_ */
