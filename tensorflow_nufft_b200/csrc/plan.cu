// plan.cu -- the plan object and the C ABI (include/b200nufft.h) of the B200 NUFFT engine.
// Orchestration follows the stages of Plan<GPUDevice,F> (nufft_plan.cu.cc:1808-2168) but not its
// structure: everything is stream-ordered, the cuFFT plan and all buffers live in the plan (so a
// cached plan costs nothing per call), no host synchronisation in set_points / execute, the
// stencil records are precomputed once per point set and reused by every transform and execute.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200nufft.h"
#include "deconv.cuh"
#include "fft_pruned.cuh"
#include "dev_common.cuh"
#include "host_params.h"
#include "interp.cuh"
#include "interp_qw.cuh"
#include "interp_ring.cuh"
#include "rowlane.cuh"
#include "points.cuh"
#include "scan_sort.cuh"
#include "spread.cuh"
#include "spread_ws.cuh"
#include "spread_ws2.cuh"
#include "spread_sweep.cuh"

using namespace b200;

namespace {

thread_local std::string g_create_error;

std::atomic<int64_t> g_allocs{0}, g_frees{0};

// Saves / restores the calling thread's current device around an entry point (a plan on another
// device must not leave the caller on that device).
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) { cudaSetDevice(dev); changed = true; }
  }
  ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
};

// Where a plan's device memory comes from: cudaMalloc/cudaFree, or the caller's callbacks.
struct MemCtx {
  b200nufft_allocator a{nullptr, nullptr, nullptr};
  int device = 0;
  cudaError_t alloc(void** p, size_t bytes) {
    g_allocs++;
    if (a.alloc) {
      *p = a.alloc(a.user, bytes, device);
      return *p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
    return cudaMalloc(p, bytes);
  }
  void free(void* p) {
    if (!p) return;
    g_frees++;
    if (a.free) a.free(a.user, p, device);
    else cudaFree(p);
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool external = false;   // carved out of a caller-owned workspace: never freed here
  // headroom: grow by 1/8 more than asked, so that slowly growing point sets do not reallocate
  cudaError_t reserve(MemCtx& m, size_t bytes, bool headroom = true) {
    if (bytes <= cap) return cudaSuccess;
    if (external) return cudaErrorMemoryAllocation;   // a bound workspace is never outgrown silently
    release(m);
    size_t want = headroom ? bytes + bytes / 8 + 256 : bytes;
    want = (want + 255) & ~static_cast<size_t>(255);
    cudaError_t e = m.alloc(&p, want);
    if (e == cudaSuccess) cap = want;
    else p = nullptr;
    return e;
  }
  void release(MemCtx& m) {
    if (p && !external) m.free(p);
    p = nullptr;
    cap = 0;
    external = false;
  }
  void bind(void* ptr, size_t bytes) {
    p = ptr;
    cap = bytes;
    external = true;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct b200nufft_plan {
  int type = 0, rank = 0, fft_sign = -1, ntransf = 1, dtype = 0, device = 0;
  bool is_double = false;
  b200nufft_opts opts{};
  KernelParams kp;
  double tol = 0;
  double kernel_scale = 0;
  int64_t n_modes[3] = {1, 1, 1};
  int64_t n_modes_tot = 1;
  int nf[3] = {1, 1, 1};
  int64_t nftot = 1;
  int batch = 1;
  int bin[3] = {1, 1, 1};
  int nbins[3] = {1, 1, 1};
  int nbtot = 1;
  int msub = 1024;
  int PX = 8, PY = 8, R = 8;
  int num_threads_compat = 1;
  int spread_method = 1, interp_method = 1;
  // TMA descriptors of a grid batch with a box of one tile: `in` feeds the interpolators' tile
  // loads, `out` the spreaders' reduce-add tile flush. Re-encoded when the pointer / box changes.
  struct TileMap {
    CUtensorMap map;
    const void* ptr = nullptr;
    int batch = 0, box_x = 0, box_y = 0, box_z = 0, coils = 0;
    bool ok = false;
  };
  TileMap tmap_in, tmap_out;
  RowLaneGeom rl{};        // row-lane 2D tile kernels (complex128; complex64 with ns > 7)
  int rl_pxt = 0, rl_lp = 0;
  bool adaptive_bin_z = false; // 3D type-2: bin depth chosen per point set (set_points)
  bool adaptive_bin_x = false; // 3D sweep spreader: bin width chosen per point set (set_points)
  bool zrange_valid = false;   // sub_desc.w holds the z extent of every subproblem (3D interp plans)
  bool ws = false;         // window-sorted keys (type-1 register-accumulating spreader)
  bool ws2 = false;        // ... with even-row windows (spread_ws2.cuh): records carry a y shift
  bool ws3 = false;        // 3D sweep spreader: even-aligned windows in x, y and z, (bin, wz, wy, wx) keys
  bool otf = false;        // ... single transform: weights evaluated inside the spreader, no stencil records
  size_t tile_smem = 0;

  // device state
  DevBuf fine;            // [batch][nftot] complex
  DevBuf fser[3];         // deconvolution factors
  std::vector<char> fser_host[3];
  cufftHandle fft = 0, fft_rem = 0;
  bool has_fft = false, has_fft_rem = false;
  int fft_rem_batch = 0;
  // Pruned 3D FFT (3 cuFFT plans on the fine grid): the slowest axis holds modes only in its first
  // zlo and last zhi planes (zero padding on input for type 2, cropped output for type 1), so the
  // 2D (x, y) transforms run on those planes only and a strided 1D transform does the z axis.
  cufftHandle fft_xy_lo = 0, fft_xy_hi = 0, fft_z = 0;
  bool pruned_fft = false;
  // own pruned + fused FFT passes (fft_pruned.cuh): complex64, power-of-two fine sizes
  bool own_fft = false;
  DevBuf fft_tw[3];        // per-axis twiddle tables (fft_fill_twiddles)
  DevBuf fft_rfac[3];      // reciprocals of the deconvolution factors, rounded once from double
  int zlo = 0, zhi = 0;

  int64_t M = 0;
  bool points_set = false;
  DevBuf folded, keys0, pairs0, pairs1, idxbuf, hist, start, wrec;   // sort: keys, two (key, index) pair buffers, sorted indices
  DevBuf bin_sizes, bin_start, num_sub, sub_start, sub_desc, misc;  // misc: scan tmp[1024] + sub_total + range flag
  DevBuf reuse;            // ReuseState (opts.reuse_points)
  DevBuf ct;               // point-major strengths [M][batch] of the running batch (2D sweep spreader)
  MemCtx mem;
  int nb_max = 1;          // largest bin count over the geometries set_points may choose
  bool ws_bound = false;   // external workspace bound (opts.external_workspace)
  int64_t ws_points = 0;   // ... for this many points
  // point-set reuse: host-side knowledge of what the device-side fingerprint describes
  bool fp_valid = false;
  int64_t fp_M = -1;
  int fp_layout = -1;
  // cross-stream / cross-thread use of one handle
  std::mutex mu;
  cudaEvent_t done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_work = false;
  // Type-1 pre-clear: after an execute the fine grid is zeroed again on an internal stream, so the
  // next execute finds it clean and the memset overlaps whatever the caller enqueues in between
  // (typically the next set_points). Off while capturing into a CUDA graph and with a bound workspace.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_cleared = nullptr;
  bool fine_cleared = false;
  const void* cleared_ptr = nullptr;
  size_t cleared_bytes = 0;
  bool call_capturing = false;
  // plan cache bookkeeping (b200nufft_plan_acquire / _release)
  bool from_cache = false;
  std::string cache_key;
  int* idx = nullptr;      // idxbuf once points are set
  int64_t sub_bound = 0;
  int* h_flag = nullptr;   // pinned

  std::vector<cudaEvent_t> ev_batch;  // 4 events per batch of the last execute (profile mode)
  int ev_batches = 0;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [4],[5]: set_points
  bool ev_exec = false, ev_setpts = false;
  int64_t launches = 0;
  char err[768] = {0};

  int* scan_tmp() const { return misc.as<int>(); }
  int* sub_total() const { return misc.as<int>() + kScanMaxBlocks; }
  // the range-check flag lives right behind the bin histogram so that one memset clears both
  int* range_flag() const { return bin_sizes.as<int>() + nbtot; }
  ReuseState* reuse_state() const { return reuse.as<ReuseState>(); }
  // skip flag read by every set_points kernel (nullptr: reuse off)
  const int* skip_flag() const { return opts.reuse_points ? &reuse.as<ReuseState>()->skip : nullptr; }
  // buffers that live in the caller's workspace when one is bound
  std::vector<DevBuf*> ws_bufs() {
    return {&fine, &bin_sizes, &bin_start, &num_sub, &sub_start, &folded, &keys0, &pairs0, &pairs1, &idxbuf,
            &hist, &start, &wrec, &sub_desc, &ct};
  }
};

namespace {

int set_err(b200nufft_plan* p, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(p->err, sizeof(p->err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(plan, expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return set_err(plan, _e == cudaErrorMemoryAllocation ? B200NUFFT_RESOURCE_EXHAUSTED      \
                                                           : B200NUFFT_INTERNAL,               \
                     "CUDA error %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__);    \
  } while (0)

#define LAUNCH_OK(plan)                                                                        \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return set_err(plan, B200NUFFT_INTERNAL, "kernel launch failed: %s at %s:%d",            \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                              \
  } while (0)

int grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(kNumSMsB200) * per_sm;
  return static_cast<int>(std::max<int64_t>(1, std::min(blocks, cap)));
}

int ilog2_ceil(int64_t n) {
  int b = 0;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

GridGeom grid_geom(const b200nufft_plan* p) {
  GridGeom g;
  g.rank = p->rank;
  for (int d = 0; d < 3; ++d) { g.nf[d] = p->nf[d]; g.bin[d] = p->bin[d]; g.nbins[d] = p->nbins[d]; }
  g.nftot = p->nftot;
  return g;
}

ModeGeom mode_geom(const b200nufft_plan* p) {
  ModeGeom m;
  m.rank = p->rank;
  for (int d = 0; d < 3; ++d) { m.n[d] = static_cast<int>(p->n_modes[d]); m.nf[d] = p->nf[d]; }
  m.ntot = p->n_modes_tot;
  m.nftot = p->nftot;
  return m;
}

template <typename F>
void points_bounds(const b200nufft_plan* p, F* lo, F* hi) {
  // PlanBase::points_upper_bound (nufft_plan.h:957-996), RADIANS_PER_SAMPLE
  F ub = MathConst<F>::pi;
  if (p->opts.points_range == B200NUFFT_RANGE_EXTENDED) ub *= F(3.0);
  if (p->opts.points_range == B200NUFFT_RANGE_INFINITE) ub = INFINITY;
  *hi = ub;
  *lo = -ub;
}

constexpr int kNumWsBufs = 15;
// Bytes of every workspace-class buffer for point sets of up to M points, in ws_bufs() order.
template <typename F>
void ws_sizes(const b200nufft_plan* p, int64_t M, size_t out[kNumWsBufs]) {
  const int64_t m = std::max<int64_t>(M, 0);
  const size_t bins = sizeof(int) * (static_cast<size_t>(p->nb_max) + 1);
  const int msub_min = p->opts.max_subproblem_size > 0 ? p->opts.max_subproblem_size : 64;
  const int64_t sub_bound = std::min<int64_t>(p->nb_max, m) + m / msub_min + 2;
  out[0] = p->opts.spread_only ? 0 : sizeof(Cplx<F>) * static_cast<size_t>(p->nftot) * p->batch;
  out[1] = out[2] = out[3] = out[4] = bins;
  out[5] = sizeof(F) * 4 * m;
  out[6] = sizeof(uint32_t) * m;
  out[7] = out[8] = sizeof(uint2) * m;
  out[9] = sizeof(int) * m;
  out[10] = sizeof(int) * radix_hist_ints(m);
  out[11] = p->otf ? 0 : sizeof(int4) * m;
  out[12] = p->otf ? 0 : sizeof(F) * m * p->R;
  out[13] = sizeof(int4) * sub_bound;
  out[14] = p->spread_method == 6 ? sizeof(Cplx<F>) * m * std::min(p->batch, p->ntransf) : 0;
}
void ws_sizes_any(const b200nufft_plan* p, int64_t M, size_t out[kNumWsBufs]) {
  if (p->is_double) ws_sizes<double>(p, M, out); else ws_sizes<float>(p, M, out);
}
inline size_t align256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

// Serialises the calls on one handle and orders its work across streams: a call on a new stream
// first waits for the event recorded after the plan's previous call.
struct PlanCall {
  b200nufft_plan* p;
  std::lock_guard<std::mutex> lock;
  DeviceGuard dev;
  cudaStream_t st;
  bool capturing = false;   // the stream is being captured into a CUDA graph: no cross-stream events
  PlanCall(b200nufft_plan* plan, cudaStream_t stream) : p(plan), lock(plan->mu), dev(plan->device), st(stream) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess) capturing = cs != cudaStreamCaptureStatusNone;
    p->call_capturing = capturing;
    if (!capturing && p->has_work && p->done && st != p->last_stream) cudaStreamWaitEvent(st, p->done, 0);
  }
  ~PlanCall() {
    if (!capturing && p->done && cudaEventRecord(p->done, st) == cudaSuccess) {
      p->has_work = true;
      p->last_stream = st;
    }
  }
};

// ------------------------------------------------------------------------------------------
// Tile-kernel dispatch on the kernel width.
// ------------------------------------------------------------------------------------------
constexpr int kInterpWarps = 4;
// Warps sharing one 3D tile (z-plane ownership). 1: every warp owns its tile and updates all 7
// planes of a point (7 independent load/FFMA/store chains per lane) and the per-point record is
// read once instead of once per warp; measured on cfg3 with the TMA flush: 2.39 ms (1 warp),
// 2.59 ms (2), 2.79 ms (4).
constexpr int kSpreadWarps3D = 1;

bool ensure_out_tensor_map(b200nufft_plan* p, const void* grid, int ntr, int box_z, int halo_x = 8);

template <int RANK, int WPT>
cudaError_t launch_spread_tile(const b200nufft_plan* p, int ntr, const float2* c, float2* fw, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr);
  const size_t smem = spread_tile_smem_bytes<RANK, WPT>(p->bin);
  // 3D: one-plane boxes (the kernel sends the planes its subproblem touched); needs 128-byte planes
  const bool plane_ok = ((p->bin[0] + 8) * (p->bin[1] + 8) * sizeof(float2)) % 128 == 0;
  const int use_tma = (p->opts.reserved[5] == 0 && (RANK == 2 || plane_ok) &&
                       ensure_out_tensor_map(const_cast<b200nufft_plan*>(p), fw, ntr, RANK == 3 ? 1 : 0)) ? 1 : 0;
  const int zrange = (RANK == 3 && p->zrange_valid && p->opts.reserved[6] == 0) ? 1 : 0;
#define SPREAD_CASE(NS)                                                                          \
  case NS: {                                                                                     \
    auto k = spread_tile_f32_kernel<NS, RANK, WPT>;                                              \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, WPT * 32, smem, st>>>(p->M, g, p->sub_total(), p->sub_desc.as<int4>(), p->idx,     \
                                    p->start.as<int4>(), p->wrec.as<float4>(), c, fw, p->tmap_out.map, use_tma, zrange); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    SPREAD_CASE(2) SPREAD_CASE(3) SPREAD_CASE(4) SPREAD_CASE(5) SPREAD_CASE(6) SPREAD_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef SPREAD_CASE
  return cudaGetLastError();
}

template <int RANK, int TZ, int NC>
cudaError_t launch_spread_ws(const b200nufft_plan* p, int ntr, const float2* c, float2* fw, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr / NC);
  const size_t smem = spread_ws_smem_bytes<RANK, NC>(p->bin);
  const int use_tma = (p->opts.reserved[5] == 0 && ensure_out_tensor_map(const_cast<b200nufft_plan*>(p), fw, ntr, 0)) ? 1 : 0;
#define WS_CASE(NS)                                                                              \
  case NS: {                                                                                     \
    auto k = spread_ws_f32_kernel<NS, RANK, TZ, NC>;                                             \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, 32, smem, st>>>(p->M, g, p->sub_total(), p->sub_desc.as<int4>(), p->idx,           \
                              p->start.as<int4>(), p->wrec.as<float4>(), c, fw, p->tmap_out.map, use_tma); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    WS_CASE(2) WS_CASE(3) WS_CASE(4) WS_CASE(5) WS_CASE(6) WS_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef WS_CASE
  return cudaGetLastError();
}

template <int NC>
cudaError_t launch_spread_ws2(const b200nufft_plan* p, int ntr, const float2* c, float2* fw, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr / NC);
  const size_t smem = spread_ws_smem_bytes<2, NC>(p->bin);
  const int use_tma = (p->opts.reserved[5] == 0 && ensure_out_tensor_map(const_cast<b200nufft_plan*>(p), fw, ntr, 0)) ? 1 : 0;
#define WS2_CASE(NS)                                                                             \
  case NS: {                                                                                     \
    auto k = spread_ws2_f32_kernel<NS, NC>;                                                          \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, 32, smem, st>>>(p->M, g, p->sub_total(), p->sub_desc.as<int4>(), p->idx,           \
                              p->start.as<int4>(), p->wrec.as<float4>(), c, fw, p->tmap_out.map, use_tma); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    WS2_CASE(2) WS2_CASE(3) WS2_CASE(4) WS2_CASE(5) WS2_CASE(6) WS2_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef WS2_CASE
  return cudaGetLastError();
}

// pm != 0: c holds point-major strengths [M][ntr] (transpose_strengths_kernel), else coil-major [ntr][M]
template <int Y>
cudaError_t launch_spread_sweep2d(const b200nufft_plan* p, int ntr, const float2* c, float2* fw, cudaStream_t st, int pm) {
  GridGeom g = grid_geom(p);
  constexpr int NC = 4 * Y;
  const int ngroups = ntr / NC;
  const int64_t nblocks = p->sub_bound * ngroups;
  if (nblocks > 2147483647LL) return cudaErrorInvalidValue;
  const size_t smem = spread_sweep2d_smem_bytes<Y>(p->bin);
  // TMA tile flush needs 128-byte aligned coil tiles in shared memory
  const bool tile_ok = ((p->bin[0] + kSweepHaloX) * (p->bin[1] + 8) * sizeof(float2)) % 128 == 0;
  const int use_tma = (p->opts.reserved[5] == 0 && tile_ok &&
                       ensure_out_tensor_map(const_cast<b200nufft_plan*>(p), fw, ntr, 0, kSweepHaloX)) ? 1 : 0;
  const bool pack = p->opts.reserved[3] == 0;
#define SWEEP_CASE(NS)                                                                            \
  case NS: {                                                                                     \
    auto k = pm ? (pack ? spread_sweep2d_f32_kernel<NS, Y, 1, 1> : spread_sweep2d_f32_kernel<NS, Y, 0, 1>)   \
                : (pack ? spread_sweep2d_f32_kernel<NS, Y, 1, 0> : spread_sweep2d_f32_kernel<NS, Y, 0, 0>);  \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<static_cast<unsigned>(nblocks), 32, smem, st>>>(p->M, g, ngroups, p->sub_total(), p->sub_desc.as<int4>(), \
                              p->idx, p->start.as<int4>(), p->wrec.as<float4>(), c, fw, p->tmap_out.map, use_tma); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    SWEEP_CASE(2) SWEEP_CASE(3) SWEEP_CASE(4) SWEEP_CASE(5) SWEEP_CASE(6) SWEEP_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef SWEEP_CASE
  return cudaGetLastError();
}

cudaError_t launch_spread_sweep3d(const b200nufft_plan* p, int ntr, const float2* c, float2* fw, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  const int64_t nblocks = p->sub_bound * ntr;
  if (nblocks > 2147483647LL) return cudaErrorInvalidValue;
  const size_t smem = spread_sweep3d_smem_bytes(p->bin);
  // one-plane TMA boxes: the planes are (bin_x + 10) * (bin_y + 8) * 8 bytes = a multiple of 128
  const bool plane_ok = ((p->bin[0] + kSweepHaloX) * (p->bin[1] + 8) * sizeof(float2)) % 128 == 0;
  const int use_tma = (p->opts.reserved[5] == 0 && plane_ok &&
                       ensure_out_tensor_map(const_cast<b200nufft_plan*>(p), fw, ntr, 1, kSweepHaloX)) ? 1 : 0;
  const bool pack = p->opts.reserved[3] == 0;
#define SWEEP3_CASE(NS)                                                                          \
  case NS: {                                                                                     \
    auto k = p->otf ? spread_sweep3d_f32_kernel<NS, 1, 1>                                        \
                    : (pack ? spread_sweep3d_f32_kernel<NS, 1, 0> : spread_sweep3d_f32_kernel<NS, 0, 0>); \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<static_cast<unsigned>(nblocks), 32, smem, st>>>(p->M, g, ntr, p->sub_total(), p->sub_desc.as<int4>(), \
                              p->idx, p->start.as<int4>(), p->wrec.as<float4>(), p->folded.as<float4>(),           \
                              static_cast<float>(p->kp.beta), static_cast<float>(p->kp.c), static_cast<float>(p->kp.half_width), \
                              c, fw, p->tmap_out.map, use_tma); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    SWEEP3_CASE(2) SWEEP3_CASE(3) SWEEP3_CASE(4) SWEEP3_CASE(5) SWEEP3_CASE(6) SWEEP3_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef SWEEP3_CASE
  return cudaGetLastError();
}

// Builds (or reuses) the TMA tensor map of a fine-grid batch [ntr][nf2][nf1][2*nf0] float32 with a
// box of one tile. cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// Builds (or reuses) the tensor map of a grid batch [ntr][nf2][nf1][2*nf0] reals with a box of
// box_x x box_y (x bin_z + 8) cells x box_coils transforms.
bool ensure_tile_map(const b200nufft_plan* p, b200nufft_plan::TileMap* tm, const void* grid, int ntr,
                     int box_x, int box_y, int box_coils, int box_z = 0 /* 0: bin_z + 8 */) {
  if (tm->ok && tm->ptr == grid && tm->batch == ntr && tm->box_x == box_x && tm->box_y == box_y &&
      tm->coils == box_coils && tm->box_z == box_z)
    return true;
  // cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link), once per process
  static const EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(fn);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  tm->ok = false;
  if (!encode) return false;
  const int rank = p->rank;
  const size_t real_bytes = p->is_double ? 8 : 4;
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  dims[0] = 2ull * p->nf[0];
  box[0] = 2u * box_x;
  cuuint64_t row = static_cast<cuuint64_t>(p->nf[0]) * 2 * real_bytes;
  for (int d = 1; d < rank; ++d) {
    dims[d] = p->nf[d];
    box[d] = d == 1 ? box_y : (box_z > 0 ? box_z : p->bin[d] + 8);
    strides[d - 1] = row;
    row *= p->nf[d];
  }
  dims[rank] = ntr;
  box[rank] = box_coils;
  strides[rank - 1] = static_cast<cuuint64_t>(p->nftot) * 2 * real_bytes;
  for (int d = 0; d <= rank; ++d) if (box[d] > 256) return false;
  CUresult r = encode(&tm->map, p->is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      rank + 1, const_cast<void*>(grid), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  tm->ptr = grid;
  tm->batch = ntr;
  tm->box_x = box_x;
  tm->box_y = box_y;
  tm->coils = box_coils;
  tm->box_z = box_z;
  tm->ok = true;
  return true;
}

// The spreaders' output map: box = one (bin + 8)^rank tile of one transform.
bool ensure_out_tensor_map(b200nufft_plan* p, const void* grid, int ntr, int box_z, int halo_x) {
  return ensure_tile_map(p, &p->tmap_out, grid, ntr, p->bin[0] + halo_x, p->bin[1] + 8, 1, box_z);
}

template <int RANK>
cudaError_t launch_interp_tile(b200nufft_plan* p, int ntr, const float2* fw, float2* c, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  const int use_tma = (p->opts.reserved[0] == 0 && ensure_tile_map(p, &p->tmap_in, fw, ntr, p->bin[0] + 8, p->bin[1] + 8, 1)) ? 1 : 0;
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr);
  const size_t smem = interp_tile_smem_bytes<RANK, kInterpWarps>(p->bin);
#define INTERP_CASE(NS)                                                                          \
  case NS: {                                                                                     \
    auto k = interp_tile_f32_kernel<NS, RANK, kInterpWarps>;                                     \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, kInterpWarps * 32, smem, st>>>(p->M, g, p->sub_total(), p->sub_desc.as<int4>(),    \
                                             p->idx, p->start.as<int4>(), p->wrec.as<float4>(),  \
                                             fw, c, p->tmap_in.map, use_tma);                           \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    INTERP_CASE(2) INTERP_CASE(3) INTERP_CASE(4) INTERP_CASE(5) INTERP_CASE(6) INTERP_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef INTERP_CASE
  return cudaGetLastError();
}

constexpr int kQwWarps = 4;

template <int RANK, int NC, int PF>
cudaError_t launch_interp_qw(b200nufft_plan* p, int ntr, const float2* fw, float2* c, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  // 3D: box of ONE z-plane (the kernel loads the planes its subproblem needs); needs 128-byte planes
  const bool plane_ok = ((p->bin[0] + kQwHaloX) * (p->bin[1] + 8) * sizeof(float2)) % 128 == 0;
  const int box_z = (RANK == 3 && plane_ok) ? 1 : 0;
  const int use_tma = (p->opts.reserved[0] == 0 && (RANK == 2 || plane_ok) &&
                       ensure_tile_map(p, &p->tmap_in, fw, ntr, p->bin[0] + kQwHaloX, p->bin[1] + 8, NC, box_z)) ? 1 : 0;
  const int zrange = (RANK == 3 && p->zrange_valid && p->opts.reserved[6] == 0) ? 1 : 0;
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr / NC);
  const size_t smem = interp_qw_smem_bytes<RANK>(p->bin, NC);
#define QW_CASE(NS)                                                                              \
  case NS: {                                                                                     \
    auto k = interp_qw_f32_kernel<NS, RANK, NC, PF, kQwWarps>;                                       \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, kQwWarps * 32, smem, st>>>(p->M, g, p->sub_total(), p->sub_desc.as<int4>(),        \
                                         p->idx, p->start.as<int4>(), p->wrec.as<float4>(),      \
                                         fw, c, p->tmap_in.map, use_tma, zrange);                       \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    QW_CASE(2) QW_CASE(3) QW_CASE(4) QW_CASE(5) QW_CASE(6) QW_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef QW_CASE
  return cudaGetLastError();
}

cudaError_t launch_interp_ring3d(b200nufft_plan* p, int ntr, const float2* fw, float2* c, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  const int64_t nblocks = p->sub_bound * ntr;
  if (nblocks > 2147483647LL) return cudaErrorInvalidValue;
  // one-plane TMA boxes: planes of (bin_x + 10) * (bin_y + 8) * 8 bytes must be a multiple of 128
  const bool plane_ok = ((p->bin[0] + kQwHaloX) * (p->bin[1] + 8) * sizeof(float2)) % 128 == 0;
  const int use_tma = (p->opts.reserved[0] == 0 && plane_ok &&
                       ensure_tile_map(p, &p->tmap_in, fw, ntr, p->bin[0] + kQwHaloX, p->bin[1] + 8, 1, 1)) ? 1 : 0;
  const size_t smem = interp_ring_smem_bytes(p->bin);
#define RING_CASE(NS)                                                                            \
  case NS: {                                                                                     \
    auto k = interp_ring3d_f32_kernel<NS, kQwWarps>;                                             \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<static_cast<unsigned>(nblocks), kQwWarps * 32, smem, st>>>(p->M, g, ntr, p->sub_total(), \
        p->sub_desc.as<int4>(), p->idx, p->start.as<int4>(), p->wrec.as<float4>(), fw, c, p->tmap_in.map, use_tma, \
        (p->zrange_valid && p->opts.reserved[6] == 0) ? 1 : 0); \
    break;                                                                                       \
  }
  switch (p->kp.ns) {
    RING_CASE(2) RING_CASE(3) RING_CASE(4) RING_CASE(5) RING_CASE(6) RING_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef RING_CASE
  return cudaGetLastError();
}

template <typename F>
cudaError_t launch_interp_rowlane(b200nufft_plan* p, int ntr, const Cplx<F>* fw, Cplx<F>* c, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  const int use_tma = (p->rank == 2 && p->opts.reserved[0] == 0 &&
                       ensure_tile_map(p, &p->tmap_in, fw, ntr, p->rl.TX, p->rl.TY, 1)) ? 1 : 0;
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr);
  const size_t smem = rowlane_smem_bytes(p->rl, sizeof(Cplx<F>));
#define RL_CASE(PXT, LP)                                                                         \
  if (p->rl_pxt == PXT && p->rl_lp == LP) {                                                      \
    auto k = p->rank == 3 ? interp_rowlane_kernel<F, PXT, LP, 4, 3> : interp_rowlane_kernel<F, PXT, LP, 4, 2>; \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, 128, smem, st>>>(p->M, g, p->rl, p->sub_total(), p->sub_desc.as<int4>(), p->idx,   \
                               p->start.as<int4>(), p->wrec.as<F>(), fw, c, p->tmap_in.map, use_tma); \
    return cudaGetLastError();                                                                   \
  }
  RL_CASE(8, 8) RL_CASE(12, 8) RL_CASE(12, 16) RL_CASE(16, 16)
#undef RL_CASE
  return cudaErrorInvalidValue;
}

constexpr int kRowLaneWarps3D = 4;   // warps sharing one 3D row-lane spreader tile (z-plane ownership)

template <typename F>
cudaError_t launch_spread_rowlane(b200nufft_plan* p, int ntr, const Cplx<F>* c, Cplx<F>* fw, cudaStream_t st) {
  GridGeom g = grid_geom(p);
  dim3 grid(static_cast<unsigned>(p->sub_bound), ntr);
  const size_t smem = rowlane_smem_bytes(p->rl, sizeof(Cplx<F>));
#define RL_CASE(PXT, LP)                                                                         \
  if (p->rl_pxt == PXT && p->rl_lp == LP) {                                                      \
    auto k = p->rank == 3 ? spread_rowlane_kernel<F, PXT, LP, 3, kRowLaneWarps3D> : spread_rowlane_kernel<F, PXT, LP, 2, 1>; \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k<<<grid, p->rank == 3 ? 32 * kRowLaneWarps3D : 32, smem, st>>>(p->M, g, p->rl, p->sub_total(), p->sub_desc.as<int4>(), p->idx,    \
                              p->start.as<int4>(), p->wrec.as<F>(), c, fw);                      \
    return cudaGetLastError();                                                                   \
  }
  RL_CASE(8, 8) RL_CASE(12, 8) RL_CASE(12, 16) RL_CASE(16, 16)
#undef RL_CASE
  return cudaErrorInvalidValue;
}

template <typename F>
int do_spread(b200nufft_plan* p, int ntr, const void* c, void* fw, cudaStream_t st) {
  if (p->M == 0) return B200NUFFT_OK;
  if (p->spread_method == 5) {
    cudaError_t e = launch_spread_rowlane<F>(p, ntr, static_cast<const Cplx<F>*>(c), static_cast<Cplx<F>*>(fw), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread rowlane launch: %s", cudaGetErrorString(e));
  } else if (p->spread_method == 7) {
    cudaError_t e = launch_spread_sweep3d(p, ntr, static_cast<const float2*>(c), static_cast<float2*>(fw), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread sweep3d launch: %s", cudaGetErrorString(e));
  } else if (p->spread_method == 6) {
    cudaError_t e;
    const float2* cc = static_cast<const float2*>(c);
    float2* ff = static_cast<float2*>(fw);
    const int nc_opt = p->opts.reserved[1];   // coils per CTA override (0 = auto)
    const int nc = nc_opt > 0 ? nc_opt : 8;
    // coil counts that are not a multiple of 4: groups of 4 through the sweep kernel, the rest
    // through the window-sorted kernel (same records, same sort)
    const int main_n = ntr & ~3;
    e = cudaSuccess;
    if (main_n > 0) {
      // Strengths to point-major order [M][main_n] (reserved[7] = 1: off): the spreader then fetches
      // the NC strengths of a point as ONE contiguous row piece (4 lanes x 16 bytes for 8 coils)
      // instead of NC scattered 8-byte gathers through the sort permutation -- one cache line per
      // point instead of NC.
      const int pm = p->opts.reserved[7] == 0 ? 1 : 0;
      const float2* src = cc;
      if (pm) {
        if (p->M % 2 == 0 && reinterpret_cast<uintptr_t>(cc) % 16 == 0)
          transpose_strengths_wide_kernel<<<dim3(static_cast<unsigned>((p->M + 63) / 64), (main_n + 31) / 32), 256, 0, st>>>(
              cc, p->ct.as<float2>(), p->M, main_n);
        else
          transpose_strengths_kernel<<<dim3(static_cast<unsigned>((p->M + 31) / 32), (main_n + 31) / 32), dim3(32, 8), 0, st>>>(
              cc, p->ct.as<float2>(), p->M, main_n);
        p->launches++;
        src = p->ct.as<float2>();
      }
      if (nc >= 16 && main_n % 16 == 0) e = launch_spread_sweep2d<4>(p, main_n, src, ff, st, pm);
      else if (nc >= 8 && main_n % 8 == 0) e = launch_spread_sweep2d<2>(p, main_n, src, ff, st, pm);
      else e = launch_spread_sweep2d<1>(p, main_n, src, ff, st, pm);
    }
    for (int k = main_n; k < ntr && e == cudaSuccess; ++k)
      e = launch_spread_ws2<1>(p, 1, cc + static_cast<int64_t>(k) * p->M, ff + static_cast<int64_t>(k) * p->nftot, st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread sweep launch: %s", cudaGetErrorString(e));
  } else if (p->spread_method == 4) {
    cudaError_t e;
    const float2* cc = static_cast<const float2*>(c);
    float2* ff = static_cast<float2*>(fw);
    const int nc_opt = p->opts.reserved[1];   // coils per CTA override (0 = auto)
    const int nc = nc_opt > 0 ? nc_opt : 8;   // measured on cfg2: 8 coils 1.38 ms, 4 coils 1.45 ms per 32 x 2M
    if (nc >= 8 && ntr % 8 == 0) e = launch_spread_ws2<8>(p, ntr, cc, ff, st);
    else if (nc >= 4 && ntr % 4 == 0) e = launch_spread_ws2<4>(p, ntr, cc, ff, st);
    else if (nc >= 2 && ntr % 2 == 0) e = launch_spread_ws2<2>(p, ntr, cc, ff, st);
    else e = launch_spread_ws2<1>(p, ntr, cc, ff, st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread ws2 launch: %s", cudaGetErrorString(e));
  } else if (p->spread_method == 3) {
    cudaError_t e;
    const float2* cc = static_cast<const float2*>(c);
    float2* ff = static_cast<float2*>(fw);
    const int nc_opt = p->opts.reserved[1];   // coils per CTA override (0 = auto)
    if (p->rank == 2) {
      const int nc = nc_opt > 0 ? nc_opt : 4;
      if (nc >= 8 && ntr % 8 == 0) e = launch_spread_ws<2, 1, 8>(p, ntr, cc, ff, st);
      else if (nc >= 4 && ntr % 4 == 0) e = launch_spread_ws<2, 1, 4>(p, ntr, cc, ff, st);
      else if (nc >= 2 && ntr % 2 == 0) e = launch_spread_ws<2, 1, 2>(p, ntr, cc, ff, st);
      else e = launch_spread_ws<2, 1, 1>(p, ntr, cc, ff, st);
    }
    else if (p->bin[2] == 4) e = launch_spread_ws<3, 12, 1>(p, ntr, cc, ff, st);
    else if (p->bin[2] == 8) e = launch_spread_ws<3, 16, 1>(p, ntr, cc, ff, st);
    else e = cudaErrorInvalidValue;
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread ws launch: %s", cudaGetErrorString(e));
  } else if (p->spread_method == 2) {
    cudaError_t e = p->rank == 2
        ? launch_spread_tile<2, 1>(p, ntr, static_cast<const float2*>(c), static_cast<float2*>(fw), st)
        : launch_spread_tile<3, kSpreadWarps3D>(p, ntr, static_cast<const float2*>(c), static_cast<float2*>(fw), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "spread tile launch: %s", cudaGetErrorString(e));
  } else {
    const int rows = p->rank == 1 ? 1 : (p->rank == 2 ? p->kp.ns : p->kp.ns * p->kp.ns);
    spread_global_kernel<F><<<grid_for(p->M * rows, 256, 16), 256, 0, st>>>(
        p->M, ntr, grid_geom(p), p->kp.ns, p->R, p->PX, p->PY, p->idx, p->start.as<int4>(),
        p->wrec.as<F>(), static_cast<const Cplx<F>*>(c), static_cast<Cplx<F>*>(fw));
    LAUNCH_OK(p);
  }
  p->launches++;
  return B200NUFFT_OK;
}

template <typename F>
int do_interp(b200nufft_plan* p, int ntr, const void* fw, void* c, cudaStream_t st) {
  if (p->M == 0) return B200NUFFT_OK;
  if (p->interp_method == 5) {
    cudaError_t e = launch_interp_rowlane<F>(p, ntr, static_cast<const Cplx<F>*>(fw), static_cast<Cplx<F>*>(c), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "interp rowlane launch: %s", cudaGetErrorString(e));
  } else if (p->interp_method == 7) {
    cudaError_t e = launch_interp_ring3d(p, ntr, static_cast<const float2*>(fw), static_cast<float2*>(c), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "interp ring launch: %s", cudaGetErrorString(e));
  } else if (p->interp_method >= 3) {
    const float2* ff = static_cast<const float2*>(fw);
    float2* cc = static_cast<float2*>(c);
    // 2D: NC coils per CTA share the record loads (the L1 data pipe is the bound); 3D: one coil
    // (cfg2 mirrored to type 2, 32 coils: 1.20 ms at 8 coils per CTA, 1.29 at 4, 2.04 at 1)
    const int nc_opt = p->opts.reserved[1];
    const int nc = nc_opt > 0 ? nc_opt : 8;
    cudaError_t e;
    if (p->rank == 3) e = p->bin[2] >= 8 ? launch_interp_qw<3, 1, 2>(p, ntr, ff, cc, st)    // deep bins = sparse set
                                         : launch_interp_qw<3, 1, 1>(p, ntr, ff, cc, st);
    else if (nc >= 8 && ntr % 8 == 0) e = launch_interp_qw<2, 8, 1>(p, ntr, ff, cc, st);
    else if (nc >= 4 && ntr % 4 == 0) e = launch_interp_qw<2, 4, 1>(p, ntr, ff, cc, st);
    else if (nc >= 2 && ntr % 2 == 0) e = launch_interp_qw<2, 2, 1>(p, ntr, ff, cc, st);
    else e = launch_interp_qw<2, 1, 1>(p, ntr, ff, cc, st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "interp qw launch: %s", cudaGetErrorString(e));
  } else if (p->interp_method == 2) {
    cudaError_t e = p->rank == 2
        ? launch_interp_tile<2>(p, ntr, static_cast<const float2*>(fw), static_cast<float2*>(c), st)
        : launch_interp_tile<3>(p, ntr, static_cast<const float2*>(fw), static_cast<float2*>(c), st);
    if (e != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "interp tile launch: %s", cudaGetErrorString(e));
  } else {
    interp_global_kernel<F><<<grid_for(p->M, 128, 32), 128, 0, st>>>(
        p->M, ntr, grid_geom(p), p->kp.ns, p->R, p->PX, p->PY, p->idx, p->start.as<int4>(),
        p->wrec.as<F>(), static_cast<const Cplx<F>*>(fw), static_cast<Cplx<F>*>(c));
    LAUNCH_OK(p);
  }
  p->launches++;
  return B200NUFFT_OK;
}

// The plan's FFT stage as pruned one-axis passes with the amplify / deconvolve step fused into the
// pass that touches the mode array (fft_pruned.cuh). f: the batch's modes [ntr][N].
struct FftDeviceExec {
  b200nufft_plan* p;
  int ntr;
  float2* fw;
  float2* f;
  cudaStream_t st;
  cudaError_t err = cudaSuccess;

  template <int LOGN, int KIND>
  void col_t(int axis, const FftColGeom& g, const float* pa, const float* po) {
    auto k = fft_col_kernel<LOGN, KIND>;
    const size_t smem = fft_col_smem_bytes(LOGN);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(static_cast<unsigned>(g.N0 >> fft_logw(LOGN)), static_cast<unsigned>(g.outer_count), static_cast<unsigned>(ntr));
    k<<<grid, kFftColThreads, smem, st>>>(g, static_cast<float>(p->fft_sign), p->fft_tw[axis].as<float2>(), fw, f, pa, po,
                                          p->fft_rfac[0].as<float>());
    p->launches++;
  }
  template <int LOGN>
  void col_k(int axis, int kind, const FftColGeom& g, const float* pa, const float* po) {
    if (kind == kFftPlain) col_t<LOGN, kFftPlain>(axis, g, pa, po);
    else if (kind == kFftFromModes) col_t<LOGN, kFftFromModes>(axis, g, pa, po);
    else col_t<LOGN, kFftToModes>(axis, g, pa, po);
  }
  void col(int axis, int kind, const FftColGeom& g, int axis_a, int axis_o) {
    const float* pa = axis_a >= 0 ? p->fft_rfac[axis_a].as<float>() : nullptr;
    const float* po = axis_o >= 0 ? p->fft_rfac[axis_o].as<float>() : nullptr;
    switch (fft_log2(p->nf[axis])) {
      case 6: col_k<6>(axis, kind, g, pa, po); break;
      case 7: col_k<7>(axis, kind, g, pa, po); break;
      case 8: col_k<8>(axis, kind, g, pa, po); break;
      case 9: col_k<9>(axis, kind, g, pa, po); break;
      case 10: col_k<10>(axis, kind, g, pa, po); break;
      default: err = cudaErrorInvalidValue;
    }
  }
  template <int LOGN>
  void row_t(const FftRowGeom& g, long long rows) {
    auto k = fft_row_kernel<LOGN>;
    const size_t smem = fft_row_smem_bytes(LOGN);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(static_cast<unsigned>(rows / fft_rows_per_cta(LOGN)), static_cast<unsigned>(ntr));
    k<<<grid, kFftThreads, smem, st>>>(g, static_cast<float>(p->fft_sign), p->fft_tw[0].as<float2>(), fw);
    p->launches++;
  }
  void row(const FftRowGeom& g, long long rows) {
    switch (fft_log2(g.n0)) {
      case 6: row_t<6>(g, rows); break;
      case 7: row_t<7>(g, rows); break;
      case 8: row_t<8>(g, rows); break;
      case 9: row_t<9>(g, rows); break;
      case 10: row_t<10>(g, rows); break;
      default: err = cudaErrorInvalidValue;
    }
  }
};

int do_fft_own(b200nufft_plan* p, int ntr, float2* fw, float2* f, cudaStream_t st) {
  FftDeviceExec ex{p, ntr, fw, f, st};
  int N[3];
  for (int d = 0; d < 3; ++d) N[d] = static_cast<int>(p->n_modes[d]);
  fft_pruned_sequence(p->type, p->rank, p->nf, N, ex);
  if (ex.err != cudaSuccess) return set_err(p, B200NUFFT_INTERNAL, "own FFT: unsupported size");
  LAUNCH_OK(p);
  return B200NUFFT_OK;
}

int do_fft(b200nufft_plan* p, int ntr, cudaStream_t st) {
  const int dir = p->fft_sign < 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
  if (p->pruned_fft) {
    const size_t cs = p->is_double ? sizeof(double2) : sizeof(float2);
    const int64_t plane = static_cast<int64_t>(p->nf[0]) * p->nf[1];
    auto exec = [&](cufftHandle h, char* ptr) {
      cufftResult r = cufftSetStream(h, st);
      if (r != CUFFT_SUCCESS) return r;
      p->launches++;
      return p->is_double ? cufftExecZ2Z(h, reinterpret_cast<cufftDoubleComplex*>(ptr), reinterpret_cast<cufftDoubleComplex*>(ptr), dir)
                          : cufftExecC2C(h, reinterpret_cast<cufftComplex*>(ptr), reinterpret_cast<cufftComplex*>(ptr), dir);
    };
    for (int t = 0; t < ntr; ++t) {
      char* base = p->fine.as<char>() + cs * p->nftot * t;
      char* hi = base + cs * plane * (p->nf[2] - p->zhi);
      cufftResult r = CUFFT_SUCCESS;
      if (p->type == 2) {   // zero-padded input: (x, y) on the populated planes, then z everywhere
        r = exec(p->fft_xy_lo, base);
        if (r == CUFFT_SUCCESS && p->zhi > 0) r = exec(p->fft_xy_hi, hi);
        if (r == CUFFT_SUCCESS) r = exec(p->fft_z, base);
      } else {              // cropped output: z everywhere, then (x, y) on the planes that are kept
        r = exec(p->fft_z, base);
        if (r == CUFFT_SUCCESS) r = exec(p->fft_xy_lo, base);
        if (r == CUFFT_SUCCESS && p->zhi > 0) r = exec(p->fft_xy_hi, hi);
      }
      if (r != CUFFT_SUCCESS) return set_err(p, B200NUFFT_INTERNAL, "cufftExec (pruned) failed: %d", (int)r);
    }
    return B200NUFFT_OK;
  }
  cufftHandle h = p->fft;
  if (ntr != p->batch) {
    if (!p->has_fft_rem || p->fft_rem_batch != ntr) {
      if (p->has_fft_rem) cufftDestroy(p->fft_rem);
      int n[3];
      for (int d = 0; d < p->rank; ++d) n[d] = p->nf[p->rank - 1 - d];
      cufftResult r = cufftPlanMany(&p->fft_rem, p->rank, n, nullptr, 1, 0, nullptr, 1, 0,
                                    p->is_double ? CUFFT_Z2Z : CUFFT_C2C, ntr);
      if (r != CUFFT_SUCCESS) return set_err(p, B200NUFFT_INTERNAL, "cufftPlanMany (remainder) failed: %d", (int)r);
      p->has_fft_rem = true;
      p->fft_rem_batch = ntr;
    }
    h = p->fft_rem;
  }
  cufftResult r = cufftSetStream(h, st);
  if (r != CUFFT_SUCCESS) return set_err(p, B200NUFFT_INTERNAL, "cufftSetStream failed: %d", (int)r);
  if (p->is_double)
    r = cufftExecZ2Z(h, p->fine.as<cufftDoubleComplex>(), p->fine.as<cufftDoubleComplex>(), dir);
  else
    r = cufftExecC2C(h, p->fine.as<cufftComplex>(), p->fine.as<cufftComplex>(), dir);
  if (r != CUFFT_SUCCESS) return set_err(p, B200NUFFT_INTERNAL, "cufftExec failed: %d", (int)r);
  p->launches++;
  return B200NUFFT_OK;
}

template <typename F>
int execute_impl(b200nufft_plan* p, void* c_, void* f_, cudaStream_t st) {
  using C = Cplx<F>;
  C* c = static_cast<C*>(c_);
  C* f = static_cast<C*>(f_);
  const ModeGeom mg = mode_geom(p);
  const F* p1 = p->fser[0].as<F>();
  const F* p2 = p->fser[1].as<F>();
  const F* p3 = p->fser[2].as<F>();
  const bool prof = p->opts.profile != 0;
  int bi = 0;
  for (int b0 = 0; b0 < p->ntransf; b0 += p->batch, ++bi) {
    const int ntr = std::min(p->batch, p->ntransf - b0);
    cudaEvent_t* ev = nullptr;
    if (prof) {
      while (static_cast<int>(p->ev_batch.size()) < 4 * (bi + 1)) {
        cudaEvent_t e;
        CUDA_OK(p, cudaEventCreate(&e));
        p->ev_batch.push_back(e);
      }
      ev = p->ev_batch.data() + 4 * bi;
    }
    C* cb = c + static_cast<int64_t>(b0) * p->M;
    C* fb = f + static_cast<int64_t>(b0) * p->n_modes_tot;
    C* fw = p->fine.as<C>();
    if (p->type == 1) {
      if (prof) cudaEventRecord(ev[0], st);
      const size_t clear_bytes = sizeof(C) * p->nftot * ntr;
      if (b0 == 0 && p->fine_cleared && p->cleared_ptr == fw && p->cleared_bytes >= clear_bytes && !p->call_capturing) {
        CUDA_OK(p, cudaStreamWaitEvent(st, p->ev_cleared, 0));   // pre-cleared after the previous execute
      } else {
        CUDA_OK(p, cudaMemsetAsync(fw, 0, clear_bytes, st));
      }
      if (b0 == 0) p->fine_cleared = false;
      int rc = do_spread<F>(p, ntr, cb, fw, st);
      if (rc) return rc;
      if (prof) cudaEventRecord(ev[1], st);
      if (p->own_fft) {   // FFT passes, the last one divides by the factors and writes the modes
        rc = do_fft_own(p, ntr, reinterpret_cast<float2*>(fw), reinterpret_cast<float2*>(fb), st);
        if (rc) return rc;
        if (prof) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[3], st); }
      } else {
        rc = do_fft(p, ntr, st);
        if (rc) return rc;
        if (prof) cudaEventRecord(ev[2], st);
        dim3 grid(static_cast<unsigned>(p->n_modes_tot / p->n_modes[0]), ntr);
        deconvolve_kernel<F><<<grid, std::min<int>(256, std::max<int>(32, ((int)p->n_modes[0] + 31) / 32 * 32)), 0, st>>>(mg, p1, p2, p3, fw, fb);
        LAUNCH_OK(p);
        p->launches++;
        if (prof) cudaEventRecord(ev[3], st);
      }
    } else {
      if (prof) cudaEventRecord(ev[0], st);
      int rc = 0;
      if (p->own_fft) {   // the first FFT pass reads the modes and amplifies them: no fill of the fine grid
        if (prof) cudaEventRecord(ev[1], st);
        rc = do_fft_own(p, ntr, reinterpret_cast<float2*>(fw), reinterpret_cast<float2*>(fb), st);
      } else {
        dim3 grid(static_cast<unsigned>((p->nftot / p->nf[0] + kAmplifyRowsPerCta - 1) / kAmplifyRowsPerCta), ntr);
        amplify_kernel<F><<<grid, 32 * kAmplifyRowsPerCta, 0, st>>>(mg, p1, p2, p3, fb, fw);
        LAUNCH_OK(p);
        p->launches++;
        if (prof) cudaEventRecord(ev[1], st);
        rc = do_fft(p, ntr, st);
      }
      if (rc) return rc;
      if (prof) cudaEventRecord(ev[2], st);
      rc = do_interp<F>(p, ntr, fw, cb, st);
      if (rc) return rc;
      if (prof) cudaEventRecord(ev[3], st);
    }
  }
  p->ev_exec = prof;
  p->ev_batches = bi;
  if (p->type == 1 && p->side && !p->call_capturing && !p->ws_bound && p->opts.reserved[2] == 0) {
    // pre-clear the first batch's fine grids for the next execute, off the caller's stream
    const size_t bytes = sizeof(C) * p->nftot * std::min(p->batch, p->ntransf);
    if (cudaEventRecord(p->ev_fork, st) == cudaSuccess && cudaStreamWaitEvent(p->side, p->ev_fork, 0) == cudaSuccess &&
        cudaMemsetAsync(p->fine.p, 0, bytes, p->side) == cudaSuccess && cudaEventRecord(p->ev_cleared, p->side) == cudaSuccess) {
      p->fine_cleared = true;
      p->cleared_ptr = p->fine.p;
      p->cleared_bytes = bytes;
    }
  }
  return B200NUFFT_OK;
}

template <typename F>
int set_points_impl(b200nufft_plan* p, int64_t M, int layout, const void* x, const void* y, const void* z,
                    cudaStream_t st) {
  if (M < 0 || M > 2000000000LL) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "invalid num_points %lld", (long long)M);
  const bool prof = p->opts.profile != 0;
  if (prof) cudaEventRecord(p->ev[4], st);
  p->M = M;
  p->points_set = true;
  p->sub_bound = 0;
  if (M == 0) {
    if (prof) cudaEventRecord(p->ev[5], st);
    p->ev_setpts = prof;
    return B200NUFFT_OK;
  }
  const int rank = p->rank;
  if (p->adaptive_bin_z) {
    // 3D interpolation: deep bins (16 x 8 x 8) when the point set is sparse in the fine grid --
    // the per-CTA latency (descriptor, TMA round trip) is then amortised over 4x the points and
    // the z-range trimming keeps the tile traffic at what the points reach (cfg4: 1.97 vs
    // 2.11 ms) -- shallow bins (16 x 8 x 2) when it is dense (cfg3 as type 2: 1.18 vs 1.29 ms).
    const int bz = static_cast<double>(M) < 0.15 * static_cast<double>(p->nftot) ? 8 : 2;
    if (bz != p->bin[2]) {
      p->bin[2] = bz;
      p->nbins[2] = (p->nf[2] + bz - 1) / bz;
      p->nbtot = p->nbins[0] * p->nbins[1] * p->nbins[2];
    }
  }
  if (p->adaptive_bin_x && p->spread_method == 7) {
    // 3D sweep spreader: narrow bins (8 x 8 x 16: 18 x 16-cell planes, 9 one-warp CTAs per SM instead of
    // 7, shorter sweeps) when the point set is dense in the fine grid (cfg3, 0.48 points per cell: 1.19
    // vs 1.29 ms), wide bins (16 x 8 x 16) when it is sparse (cfg4's set as type 1: 1.46 vs 1.66 ms;
    // 800k points in 128^3: 0.46 vs 0.58 ms)
    const int bx = static_cast<double>(M) > 0.25 * static_cast<double>(p->nftot) ? 8 : 16;
    if (bx != p->bin[0]) {
      p->bin[0] = bx;
      p->nbins[0] = (p->nf[0] + bx - 1) / bx;
      p->nbtot = p->nbins[0] * p->nbins[1] * p->nbins[2];
    }
  }
  // Buffers: a bound workspace must already hold M points; otherwise grow the plan's own buffers
  // (geometric headroom; b200nufft_reserve sizes them ahead of time so that this never allocates).
  if (p->ws_bound) {
    if (M > p->ws_points)
      return set_err(p, B200NUFFT_RESOURCE_EXHAUSTED, "workspace was bound for %lld points, set_points got %lld",
                     (long long)p->ws_points, (long long)M);
  } else if (p->opts.external_workspace) {
    return set_err(p, B200NUFFT_INVALID_ARGUMENT, "external_workspace plan: call b200nufft_bind_workspace first");
  } else {
    size_t need[kNumWsBufs];
    ws_sizes<F>(p, M, need);
    auto bufs = p->ws_bufs();
    for (int i = 1; i < kNumWsBufs; ++i) {
      if (need[i] > bufs[i]->cap) p->fp_valid = false;   // contents are lost with the old buffer
      CUDA_OK(p, bufs[i]->reserve(p->mem, need[i]));
    }
  }

  // Point-set reuse: fingerprint the raw coordinates; on a match with the set the buffers hold,
  // the kernels below exit at once (skip flag), all on the stream, no host round trip.
  const int* skip = p->skip_flag();
  if (skip) {
    const int allow = (p->fp_valid && p->fp_M == M && p->fp_layout == layout) ? 1 : 0;
    const int64_t wpr = sizeof(F) / 4;   // 32-bit words per real
    const int64_t n0 = (layout == 1 ? M * rank : M) * wpr;
    const int64_t n1 = (layout == 0 && rank > 1) ? M * wpr : 0;
    const int64_t n2 = (layout == 0 && rank > 2) ? M * wpr : 0;
    fingerprint_kernel<<<grid_for(n0 + n1 + n2, 256 * 4, 4), 256, 0, st>>>(
        static_cast<const uint32_t*>(x), n0, static_cast<const uint32_t*>(y), n1,
        static_cast<const uint32_t*>(z), n2, allow, p->reuse_state());
    LAUNCH_OK(p);
    p->launches++;
    p->fp_valid = true;
    p->fp_M = M;
    p->fp_layout = layout;
  }

  BinGeom bg{};
  bg.rank = rank;
  bg.rounding = 0;
  for (int d = 0; d < 3; ++d) { bg.nf[d] = p->nf[d]; bg.bin[d] = p->bin[d]; bg.nbins[d] = p->nbins[d]; }
  bg.ws = p->ws ? 1 : 0;
  // Window counts per bin and dimension. Even-aligned stencil starts (relative to the tile origin
  // bin - 4) run from 0 to bin + 4 - ns/2, i.e. (bin + 4 - ns/2) / 2 + 1 window positions; unaligned
  // ones (first-generation window sort) use every row.
  const int wmax_x = (p->bin[0] + 4 - p->kp.ns / 2) / 2 + 1;
  const int wmax_y = (p->bin[1] + 4 - p->kp.ns / 2) / 2 + 1;
  const int wmax_z = (p->bin[2] + 4 - p->kp.ns / 2) / 2 + 1;
  bg.WX = (p->ws2 || p->ws3) ? wmax_x : p->bin[0] / 2 + 3;
  bg.WY = (p->ws2 || p->ws3) ? wmax_y : p->bin[1] + 7;
  bg.align_x = p->is_double ? 0 : 1;
  bg.align_y = (p->ws2 || p->ws3) ? 1 : 0;
  bg.align_z = p->ws3 ? 1 : 0;
  bg.WZ = p->ws3 ? wmax_z : 1;
  bg.zkey = (p->interp_method == 7 && p->type == 2 && rank == 3) ? 1 : 0;
  if (bg.zkey) bg.WZ = std::min(p->bin[2] + 9 - p->kp.ns, kRingMaxZ);   // stencil z starts per tile (interp_ring.cuh)
  const int64_t key_space = static_cast<int64_t>(p->nbtot) * (p->ws ? bg.WX * bg.WY * bg.WZ : (bg.zkey ? bg.WZ : 1));

  if (skip) {
    clear_ints_kernel<<<grid_for(p->nbtot + 1, 256), 256, 0, st>>>(p->bin_sizes.as<int>(), p->nbtot + 1, skip);
    p->launches++;
  } else {
    CUDA_OK(p, cudaMemsetAsync(p->bin_sizes.p, 0, sizeof(int) * (p->nbtot + 1), st));   // histogram + range flag
  }
  F lo, hi;
  points_bounds<F>(p, &lo, &hi);
  const int check = p->opts.check_points_range && p->opts.points_range != B200NUFFT_RANGE_INFINITE;
  fold_key_kernel<F><<<grid_for(M, 256, 8), 256, 0, st>>>(
      M, layout, static_cast<const F*>(x), static_cast<const F*>(y), static_cast<const F*>(z),
      p->opts.points_range, check, lo, hi, bg, static_cast<F>(p->kp.half_width), p->folded.as<F>(),
      p->keys0.as<uint32_t>(), p->bin_sizes.as<int>(),
      p->range_flag(), skip);
  LAUNCH_OK(p);
  p->launches++;

  p->launches += radix_sort_index(p->keys0.as<uint32_t>(), p->pairs0.as<uint2>(), p->pairs1.as<uint2>(),
                                  p->idxbuf.as<int>(), M, ilog2_ceil(key_space), p->hist.as<int>(),
                                  p->scan_tmp(), st, skip);
  LAUNCH_OK(p);
  p->idx = p->idxbuf.as<int>();

  // Points per subproblem: the reference caps at 1024 (gpu_max_subproblem_size, nufft_options.h:153).
  // Small point sets get smaller subproblems so that the launch still fills the 148 SMs.
  if (p->opts.max_subproblem_size > 0) {
    p->msub = p->opts.max_subproblem_size;
  } else {
    const int64_t per_item = M * std::min(p->ntransf, p->batch) / (static_cast<int64_t>(kNumSMsB200) * 16);
    int ms = 64;
    while (ms < 1024 && ms < per_item) ms *= 2;
    p->msub = ms;
  }
  // No host read of the subproblem count (the reference blocks on it, nufft_plan.cu.cc:3011):
  // launch the bound, surplus CTAs exit on the device-side count.
  p->sub_bound = std::min<int64_t>(p->nbtot, M) + M / p->msub;
  if (p->nbtot <= kScanSmallMax) {
    bins_small_kernel<<<1, 1024, 0, st>>>(p->bin_sizes.as<int>(), p->nbtot, p->msub, p->bin_start.as<int>(),
                                         p->sub_start.as<int>(), p->sub_total(), p->sub_desc.as<int4>(), skip);
    p->launches++;
  } else {
    p->launches += exclusive_scan_i32(p->bin_sizes.as<int>(), p->bin_start.as<int>(), p->nbtot,
                                      p->scan_tmp(), nullptr, st, skip);
    subproblem_count_kernel<<<ceil_div(p->nbtot, 256), 256, 0, st>>>(p->bin_sizes.as<int>(), p->nbtot, p->msub,
                                                                    p->num_sub.as<int>(), skip);
    p->launches++;
    p->launches += exclusive_scan_i32(p->num_sub.as<int>(), p->sub_start.as<int>(), p->nbtot, p->scan_tmp(),
                                      p->sub_total(), st, skip);
    subproblem_desc_kernel<<<ceil_div(p->nbtot, 256), 256, 0, st>>>(
        p->bin_sizes.as<int>(), p->bin_start.as<int>(), p->sub_start.as<int>(), p->nbtot, p->msub,
        p->sub_desc.as<int4>(), skip);
    p->launches++;
  }
  LAUNCH_OK(p);

  const int align_x = (!p->is_double) ? 1 : 0;
  const int align = align_x | ((p->ws2 || p->ws3) ? 2 : 0) | (p->ws3 ? 4 : 0);
  if (p->otf) {
    // single-transform 3D sweep plans evaluate the weights inside the spreader: no stencil records
  } else if (p->PX == 8 && p->PY == 8 && rank >= 2) {
    const F beta = static_cast<F>(p->kp.beta), cc = static_cast<F>(p->kp.c), hw = static_cast<F>(p->kp.half_width);
    if (rank == 2)
      stencil_record8_kernel<F, 2><<<grid_for(M * 2, 256, 16), 256, 0, st>>>(
          M, p->idx, p->folded.as<F>(), p->kp.ns, beta, cc, hw,
          align, p->start.as<int>(), p->wrec.as<F>(), skip);
    else
      stencil_record8_kernel<F, 3><<<grid_for(M * 3, 256, 16), 256, 0, st>>>(
          M, p->idx, p->folded.as<F>(), p->kp.ns, beta, cc, hw,
          p->ws3 ? align : align_x, p->start.as<int>(), p->wrec.as<F>(), skip);
  } else {
    stencil_record_kernel<F><<<grid_for(M, std::max(1, 256 / p->R), 16), dim3(p->R, std::max(1, 256 / p->R)), 0, st>>>(
        M, rank, p->idx, p->folded.as<F>(), p->kp.ns,
        static_cast<F>(p->kp.beta), static_cast<F>(p->kp.c), static_cast<F>(p->kp.half_width), align_x,
        p->R, p->PX, p->PY, p->start.as<int4>(), p->wrec.as<F>(), skip);
  }
  LAUNCH_OK(p);
  if (!p->otf) p->launches++;

  p->zrange_valid = false;
  const bool zr_interp = (p->interp_method == 3 || p->interp_method == 7) && (p->type == 2 || p->opts.spread_only);
  const bool zr_spread = p->spread_method == 2 && (p->type == 1 || p->opts.spread_only);
  if (rank == 3 && !p->is_double && (zr_interp || zr_spread) && p->sub_bound > 0) {
    subproblem_zrange_kernel<<<ceil_div(p->sub_bound, 8), 256, 0, st>>>(p->sub_total(), p->start.as<int4>(),
                                                                       p->sub_desc.as<int4>(), skip);
    LAUNCH_OK(p);
    p->launches++;
    p->zrange_valid = true;
  }

  if (prof) cudaEventRecord(p->ev[5], st);
  p->ev_setpts = prof;

  if (check) {
    CUDA_OK(p, cudaMemcpyAsync(p->h_flag, p->range_flag(), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(p, cudaStreamSynchronize(st));
    if (*p->h_flag) {
      int d = 0;
      while (d < 3 && !((*p->h_flag >> d) & 1)) ++d;
      p->points_set = false;
      return set_err(p, B200NUFFT_INVALID_ARGUMENT,
                     "Found points outside expected range for dimension %d. Valid range is [%g, %g]. "
                     "Check your points and/or set a less restrictive value for options.points_range.",
                     d, (double)lo, (double)hi);
    }
  }
  return B200NUFFT_OK;
}

template <typename F>
int create_impl(b200nufft_plan* p) {
  // ---- kernel + grid parameters (setup_spreader, set_grid_size) ----
  // Upsampling factor. Plan<GPUDevice> always uses 2.0 (nufft_plan.cu.cc:1855-1857): the default.
  // opts.upsampling = 1: sigma = 1.25 (smaller FFT, wider kernel); 2: the CPU plan's automatic choice
  // (PlanBase::set_default_options, nufft_plan.h:739-752): 1.25 for large grids at tol >= 1e-9.
  // Interp / Spread ops are pinned to 2.0 (nufft_kernels.cc:457-460).
  double sigma = 2.0;
  if (!p->opts.spread_only) {
    if (p->opts.upsampling == 1) sigma = 1.25;
    else if (p->opts.upsampling == 2) {
      if (static_cast<F>(p->tol) >= F(1e-9)) {
        const int64_t gs = p->n_modes_tot;
        if ((p->rank == 1 && gs > 10000000) || (p->rank == 2 && gs > 300000) || (p->rank == 3 && gs > 3000000)) sigma = 1.25;
      }
    } else if (p->opts.upsampling != 0) {
      return set_err(p, B200NUFFT_INVALID_ARGUMENT, "opts.upsampling must be 0 (2.0), 1 (1.25) or 2 (automatic), got %d",
                     p->opts.upsampling);
    }
  }
  p->kp = make_kernel_params<F>(static_cast<F>(p->tol), sigma);
  p->nftot = 1;
  for (int d = 0; d < p->rank; ++d) {
    if (!fine_grid_size(p->n_modes[d], sigma, p->kp.ns, p->opts.spread_only != 0, &p->nf[d])) {
      return set_err(p, B200NUFFT_INVALID_ARGUMENT,
                     "Invalid grid size: %lld. Value should be even, larger than the kernel (%d) and have no "
                     "prime factors larger than 5.", (long long)p->n_modes[d], 2 * p->kp.ns);
    }
    p->nftot *= p->nf[d];
  }
  // The reference batches min(T, 8) transforms (cuFINUFFT heuristic, nufft_plan.cu.cc:1923-1928).
  // With 180 GB of HBM a larger batch costs nothing and saves launches / tail effects (cfg2:
  // 2.52 -> 2.39 ms per 32 coils): min(T, 32), capped so that the fine-grid batch stays <= 4 GiB.
  {
    const int64_t grid_bytes = p->nftot * static_cast<int64_t>(sizeof(Cplx<F>));
    const int cap = static_cast<int>(std::max<int64_t>(1, (int64_t(4) << 30) / std::max<int64_t>(1, grid_bytes)));
    p->batch = p->opts.max_batch_size > 0 ? std::min(p->opts.max_batch_size, p->ntransf)
                                          : std::min(std::min(p->ntransf, 32), cap);
    p->batch = std::max(1, std::min(p->batch, 65535));   // the batch is a grid dimension of every launch
  }
  if (p->nftot * p->batch > 2000000000LL) {
    // shrink the batch rather than fail (the reference errors out above kMaxArraySize elements)
    p->batch = static_cast<int>(std::max<int64_t>(1, 2000000000LL / p->nftot));
  }
  if (p->opts.spread_only) p->kernel_scale = kernel_scale_factor<F>(p->rank, p->kp);

  // ---- weight record layout ----
  const int ns = p->kp.ns;
  p->PX = ns <= 7 ? 8 : ((ns + 1 + 3) / 4) * 4;
  p->PY = ns <= 7 ? 8 : ((ns + 3) / 4) * 4;
  p->R = p->PX + (p->rank > 1 ? p->PY : 0) + (p->rank > 2 ? p->PY : 0);

  // ---- method + bin geometry ----
  const bool tile_ok = !p->is_double && ns <= 7 && p->rank >= 2;
  // auto: 2D -> window-sorted register-accumulating spreader (4 = even-row windows for NUFFT plans,
  // 3 for spread-only plans); 3D -> plane-owner tile kernel (2)
  // (measured on B200, cfg2 per 32 coils: 1.38 ms (4) vs 1.70 ms (3) vs 4.5 ms (2); cfg3: 3.2 ms
  //  tile vs >= 4.0 ms window-sorted)
  // (round 2: sweep spreaders 6 / 7 -- cfg2 0.96 ms vs 1.29 (4); cfg3 1.32 ms vs 2.36 (2); cfg4's point set as
  //  type 1: 1.63 ms vs 1.97 (2))
  p->spread_method = (p->opts.spread_method == 0) ? (tile_ok ? (p->rank == 2 ? 6 : 7) : 1) : p->opts.spread_method;
  // interpolator: 3 = quarter-warp gather (measured: cfg2-type2 0.56 vs 0.92 ms per 8 coils, cfg3-type2
  // 1.18 vs 1.61 ms, cfg4 2.0 vs 3.8 ms per 2 coils against the lanes-over-stencil tile kernel 2)
  p->interp_method = (p->opts.interp_method == 0) ? (tile_ok ? 3 : 1)
                                                  : (p->opts.interp_method == 7 ? 7 : std::min(p->opts.interp_method, 3));
  // 7: 3D ring interpolator (z-slab streaming), type-2 NUFFT plans only (its sort key carries the z start)
  if (p->interp_method == 7 && (p->rank != 3 || p->type != 2 || p->opts.spread_only || !tile_ok)) p->interp_method = tile_ok ? 3 : 1;
  if (!tile_ok) { p->spread_method = 1; p->interp_method = 1; }
  // complex128, and complex64 with widths the float tile kernels do not cover (ns 8..15; e.g.
  // sigma = 1.25 at tol 1e-6 -> ns = 10), 2D and 3D: row-lane tile kernels (rowlane.cuh) unless the
  // generic kernels are asked for
  const bool rl_ok = p->rank >= 2 && ns <= 15 && (p->is_double || ns > 7);
  if (rl_ok) {
    if (p->opts.spread_method != 1) p->spread_method = 5;
    if (p->opts.interp_method != 1) p->interp_method = 5;
    p->rl_pxt = p->PX;
    p->rl_lp = p->PY <= 8 ? 8 : 16;
  }
  // ws2 is 2D, type-1 NUFFT plans only (its records are not usable by the other kernels)
  // ws2 (4) and the sweep spreader (6) are 2D, type-1 NUFFT plans only (their records carry a row shift)
  if ((p->spread_method == 4 || p->spread_method == 6) && (p->rank != 2 || p->type != 1 || p->opts.spread_only))
    p->spread_method = 3;
  // 7: 3D sweep spreader (type-1 NUFFT plans; its records carry y and z shifts)
  if (p->spread_method == 7 && (p->rank != 3 || p->type != 1 || p->opts.spread_only || !tile_ok)) p->spread_method = tile_ok ? 2 : 1;
  const bool ws_any = p->spread_method == 3 || p->spread_method == 4 || p->spread_method == 6 || p->spread_method == 7;
  int def_bin[3] = {1, 1, 1};
  if (p->rank == 1) { def_bin[0] = 1024; }
  else if (p->rank == 2) {
    // window-sorted spreader: small tiles (24 x 16 cells) so that 4 coils' tiles + the stage fit
    // ~8 CTAs per SM (measured best on cfg2: 0.48 ms per 8 coils vs 0.79 ms at 32 x 32)
    def_bin[0] = (p->type == 1 && ws_any) ? 16 : 32;
    def_bin[1] = (p->type == 1 && ws_any) ? 8 : 32;
    if (p->type == 2 && p->interp_method == 3) { def_bin[0] = 16; def_bin[1] = 16; }   // cfg2-type2 0.559 vs 0.565 ms, cfg1 12 vs 14 us
    if (rl_ok) { def_bin[0] = 16; def_bin[1] = 16; }
  }
  else {
    def_bin[0] = 16;
    def_bin[1] = 16;
    def_bin[2] = (p->type == 2) ? 2 : (p->spread_method == 3 ? 8 : 2);
    if (p->type == 1 && p->spread_method == 3) def_bin[1] = 8;
    if (p->type == 2 && p->interp_method == 3) def_bin[1] = 8;   // cfg3-type2 1.18 vs 1.35 ms at 16 x 16 x 2
    if (p->type == 1 && p->spread_method == 2) def_bin[1] = 8;   // cfg3 2.79 vs 3.01 ms at 16 x 16 x 2 (TMA flush)
    if (p->type == 1 && p->spread_method == 7) { def_bin[1] = 8; def_bin[2] = 16; }   // ring of 8 planes: depth is free
    if (p->type == 2 && p->interp_method == 7) { def_bin[1] = 8; def_bin[2] = 16; }   // ring of 8 planes: depth is free
    if (rl_ok) { def_bin[0] = p->is_double ? 8 : 16; def_bin[1] = 8; def_bin[2] = 4; }   // tile = bin + ns + 1 per dim
  }
  p->nbtot = 1;
  for (int d = 0; d < 3; ++d) {
    p->bin[d] = d < p->rank ? (p->opts.bin_dims[d] > 0 ? p->opts.bin_dims[d] : def_bin[d]) : 1;
    p->nbins[d] = d < p->rank ? (p->nf[d] + p->bin[d] - 1) / p->bin[d] : 1;
    p->nbtot *= p->nbins[d];
  }
  p->adaptive_bin_z = p->rank == 3 && p->type == 2 && !p->opts.spread_only && tile_ok && p->interp_method == 3 &&
                      p->opts.bin_dims[2] == 0 && p->bin[0] == 16 && p->bin[1] == 8;
  p->adaptive_bin_x = p->rank == 3 && p->type == 1 && !p->opts.spread_only && p->spread_method == 7 &&
                      p->opts.bin_dims[0] == 0 && p->bin[0] == 16;
  p->nb_max = p->nbtot;
  if (p->adaptive_bin_z) {   // set_points picks a bin depth of 2 or 8
    for (int bz : {2, 8}) p->nb_max = std::max(p->nb_max, p->nbins[0] * p->nbins[1] * ((p->nf[2] + bz - 1) / bz));
  }
  if (p->adaptive_bin_x) p->nb_max = std::max(p->nb_max, ((p->nf[0] + 7) / 8) * p->nbins[1] * p->nbins[2]);
  p->msub = p->opts.max_subproblem_size > 0 ? p->opts.max_subproblem_size : 1024;  // refined per set_points
  const bool uses_tile = (p->type == 1 || p->opts.spread_only) ? p->spread_method >= 2 : false;
  p->ws = uses_tile && p->type == 1 && ws_any;
  p->ws2 = p->ws && (p->spread_method == 4 || p->spread_method == 6);
  p->ws3 = p->ws && p->spread_method == 7;
  // On-the-fly weights are OFF by default (reserved[7] = 2 turns them on): measured on cfg3 the
  // spreader goes from 1.29 to 4.85 ms while set_points only drops from 1.18 to 0.71 ms -- one lane
  // evaluating 24 kernel values per point serialises ~1400 instructions per batch in a one-warp CTA.
  p->otf = p->ws3 && p->ntransf == 1 && p->opts.reserved[7] == 2;
  if (p->ws3 && ((p->bin[1] & 1) || (p->bin[2] & 1)))
    return set_err(p, B200NUFFT_INVALID_ARGUMENT, "3D sweep spreader needs even bin_dims[1] and bin_dims[2]");
  if (p->ws2 && (p->bin[1] & 1))
    return set_err(p, B200NUFFT_INVALID_ARGUMENT, "even-row window spreader needs an even bin_dims[1]");
  if (p->ws && !p->ws3 && p->rank == 3 && p->bin[2] != 4 && p->bin[2] != 8)
    return set_err(p, B200NUFFT_INVALID_ARGUMENT, "window-sorted 3D spreader needs bin_dims[2] of 4 or 8");
  if (p->ws && static_cast<int64_t>(p->nbtot) * (p->bin[0] / 2 + 3) * (p->bin[1] + 7) * (p->ws3 ? p->bin[2] / 2 + 3 : 1) >= (int64_t(1) << 31)) {
    p->ws = false;
    p->ws2 = false;
    p->ws3 = false;
    p->spread_method = 2;
  }
  const bool uses_tile_i = (p->type == 2 || p->opts.spread_only) ? p->interp_method >= 2 : false;
  if (rl_ok && (p->spread_method == 5 || p->interp_method == 5)) {
    p->rl = rowlane_geom(p->bin, p->rank, ns, p->rl_pxt, p->rl_lp, p->R, p->PX, p->PY, p->is_double ? 0 : 1);
    p->tile_smem = rowlane_smem_bytes(p->rl, sizeof(Cplx<F>));
    if (p->tile_smem > 227 * 1024) {   // user-chosen bins too large: fall back to the generic kernels
      p->spread_method = 1;
      p->interp_method = 1;
    }
  } else if (uses_tile || uses_tile_i) {
    // tile pitch: bin_x + 8 = 8 (mod 16) cells for the round-1 tile kernels; the sweep spreaders need an
    // ODD number of 16-byte cell pairs per row ((bin_x + 10) / 2), i.e. any multiple of 8
    const bool sweep = uses_tile && (p->spread_method == 6 || p->spread_method == 7);
    if (sweep ? (p->bin[0] % 8) != 0 : (p->bin[0] % 16) != 0)
      return set_err(p, B200NUFFT_INVALID_ARGUMENT, "bin_dims[0] must be a multiple of 16 for the tile kernels (of 8 for the sweep spreaders)");
    size_t need = 0;
    if (uses_tile && p->spread_method == 6) need = std::max(need, spread_sweep2d_smem_bytes<4>(p->bin));
    else if (uses_tile && p->spread_method == 7) need = std::max(need, spread_sweep3d_smem_bytes(p->bin));
    else if (uses_tile && ws_any) need = std::max(need, p->rank == 2 ? spread_ws_smem_bytes<2, 8>(p->bin) : spread_ws_smem_bytes<3, 1>(p->bin));
    else if (uses_tile) need = std::max(need, p->rank == 2 ? spread_tile_smem_bytes<2, 1>(p->bin) : spread_tile_smem_bytes<3, kSpreadWarps3D>(p->bin));
    if (uses_tile_i && p->interp_method == 7) need = std::max(need, interp_ring_smem_bytes(p->bin));
    else if (uses_tile_i && p->interp_method >= 3) need = std::max(need, p->rank == 2 ? interp_qw_smem_bytes<2>(p->bin, 8) : interp_qw_smem_bytes<3>(p->bin));
    else if (uses_tile_i) need = std::max(need, p->rank == 2 ? interp_tile_smem_bytes<2, kInterpWarps>(p->bin) : interp_tile_smem_bytes<3, kInterpWarps>(p->bin));
    p->tile_smem = need;
    if (p->tile_smem > 227 * 1024)
      return set_err(p, B200NUFFT_RESOURCE_EXHAUSTED, "tile of %zu bytes exceeds shared memory", p->tile_smem);
  }

  // ---- deconvolution factors (host, then H2D; as the reference does, nufft_plan.cu.cc:1988-2021) ----
  p->num_threads_compat = p->opts.num_threads_compat > 0
      ? p->opts.num_threads_compat
      : std::max(1u, std::thread::hardware_concurrency());
  if (!p->opts.spread_only) {
    for (int d = 0; d < p->rank; ++d) {
      const int nc = p->nf[d] / 2 + 1;
      p->fser_host[d].resize(sizeof(F) * nc);
      kernel_fseries<F>(p->nf[d], p->kp, p->opts.fseries_mode, p->num_threads_compat,
                        reinterpret_cast<F*>(p->fser_host[d].data()));
      CUDA_OK(p, p->fser[d].reserve(p->mem, sizeof(F) * nc, false));
      CUDA_OK(p, cudaMemcpy(p->fser[d].p, p->fser_host[d].data(), sizeof(F) * nc, cudaMemcpyHostToDevice));
    }
    if (!p->opts.external_workspace)
      CUDA_OK(p, p->fine.reserve(p->mem, sizeof(Cplx<F>) * p->nftot * p->batch, false));
    const cufftType ftype = p->is_double ? CUFFT_Z2Z : CUFFT_C2C;
    // reserved[4]: 0 = own pruned passes when eligible, else cuFFT (3D: pruned three-plan scheme);
    // 1 = one full cuFFT plan; 2 = cuFFT only (3D: the three-plan scheme)
    p->own_fft = !p->is_double && p->opts.reserved[4] == 0 && fft_pruned_ok(p->rank, p->nf, p->n_modes);
    if (p->own_fft) {
      for (int d = 0; d < p->rank; ++d) {
        const int logn = fft_log2(p->nf[d]);
        std::vector<float2> tw(fft_tw_count(logn) + 1);
        fft_fill_twiddles(logn, p->fft_sign < 0 ? -1 : 1, tw.data());
        CUDA_OK(p, p->fft_tw[d].reserve(p->mem, sizeof(float2) * tw.size(), false));
        CUDA_OK(p, cudaMemcpy(p->fft_tw[d].p, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
        const int nc = p->nf[d] / 2 + 1;
        std::vector<float> rf(nc);
        const float* fs = reinterpret_cast<const float*>(p->fser_host[d].data());
        for (int k = 0; k < nc; ++k) rf[k] = static_cast<float>(1.0 / static_cast<double>(fs[k]));
        CUDA_OK(p, p->fft_rfac[d].reserve(p->mem, sizeof(float) * nc, false));
        CUDA_OK(p, cudaMemcpy(p->fft_rfac[d].p, rf.data(), sizeof(float) * nc, cudaMemcpyHostToDevice));
      }
    }
    if (!p->own_fft && p->rank == 3 && p->opts.reserved[4] != 1) {
      // modes k = -(n/2) .. (n-1)/2 live in fine planes [0, zlo) and [nf - zhi, nf)
      p->zlo = static_cast<int>((p->n_modes[2] - 1) / 2 + 1);
      p->zhi = static_cast<int>(p->n_modes[2] / 2);
      if (p->zlo + p->zhi < p->nf[2]) {
        int nxy[2] = {p->nf[1], p->nf[0]};
        int nz[1] = {p->nf[2]};
        const int plane = p->nf[0] * p->nf[1];
        bool ok = cufftPlanMany(&p->fft_xy_lo, 2, nxy, nullptr, 1, 0, nullptr, 1, 0, ftype, p->zlo) == CUFFT_SUCCESS;
        if (ok && p->zhi > 0) {
          ok = cufftPlanMany(&p->fft_xy_hi, 2, nxy, nullptr, 1, 0, nullptr, 1, 0, ftype, p->zhi) == CUFFT_SUCCESS;
          if (!ok) { cufftDestroy(p->fft_xy_lo); p->fft_xy_lo = 0; }
        }
        if (ok) {
          ok = cufftPlanMany(&p->fft_z, 1, nz, nz, plane, 1, nz, plane, 1, ftype, plane) == CUFFT_SUCCESS;
          if (!ok) {
            cufftDestroy(p->fft_xy_lo); p->fft_xy_lo = 0;
            if (p->fft_xy_hi) { cufftDestroy(p->fft_xy_hi); p->fft_xy_hi = 0; }
          }
        }
        p->pruned_fft = ok;
      }
    }
    if (!p->pruned_fft && !p->own_fft) {
      int n[3];
      for (int d = 0; d < p->rank; ++d) n[d] = p->nf[p->rank - 1 - d];
      cufftResult r = cufftPlanMany(&p->fft, p->rank, n, nullptr, 1, 0, nullptr, 1, 0, ftype, p->batch);
      if (r != CUFFT_SUCCESS) return set_err(p, B200NUFFT_INTERNAL, "cufftPlanMany failed: %d", (int)r);
      p->has_fft = true;
    }
  }
  if (!p->opts.external_workspace) {
    for (DevBuf* b : {&p->bin_sizes, &p->bin_start, &p->num_sub, &p->sub_start})
      CUDA_OK(p, b->reserve(p->mem, sizeof(int) * (static_cast<size_t>(p->nb_max) + 1), false));
  }
  CUDA_OK(p, p->misc.reserve(p->mem, sizeof(int) * (kScanMaxBlocks + 8), false));
  if (p->opts.reuse_points) {
    CUDA_OK(p, p->reuse.reserve(p->mem, sizeof(ReuseState), false));
    CUDA_OK(p, cudaMemset(p->reuse.p, 0, sizeof(ReuseState)));
  }
  CUDA_OK(p, cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
  if (p->type == 1 && !p->opts.spread_only && !p->opts.external_workspace) {
    CUDA_OK(p, cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    CUDA_OK(p, cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    CUDA_OK(p, cudaEventCreateWithFlags(&p->ev_cleared, cudaEventDisableTiming));
  }
  CUDA_OK(p, cudaMallocHost(&p->h_flag, sizeof(int)));
  if (p->opts.profile) {
    CUDA_OK(p, cudaEventCreate(&p->ev[4]));
    CUDA_OK(p, cudaEventCreate(&p->ev[5]));
  }
  return B200NUFFT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

void b200nufft_default_opts(b200nufft_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->points_range = B200NUFFT_RANGE_EXTENDED;
}

namespace {
int plan_create_common(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims, int fft_sign,
                       int num_transforms, double tol, int dtype, const b200nufft_opts* opts, int device,
                       const b200nufft_allocator* allocator) {
  if (!out) return B200NUFFT_INVALID_ARGUMENT;
  *out = nullptr;
  auto fail = [&](int code, const std::string& m) { g_create_error = m; return code; };
  if (type != 1 && type != 2) return fail(B200NUFFT_UNIMPLEMENTED, "type-3 transforms are not implemented");
  if (rank < 1 || rank > 3) return fail(B200NUFFT_UNIMPLEMENTED, "rank must be 1, 2 or 3, but got: " + std::to_string(rank));
  if (num_transforms < 1) return fail(B200NUFFT_INVALID_ARGUMENT, "num_transforms must be >= 1");
  if (dtype != B200NUFFT_COMPLEX64 && dtype != B200NUFFT_COMPLEX128) return fail(B200NUFFT_INVALID_ARGUMENT, "invalid dtype");
  if (fft_sign != 1 && fft_sign != -1) return fail(B200NUFFT_INVALID_ARGUMENT, "fft_sign must be -1 or +1");
  if (allocator && (!allocator->alloc || !allocator->free))
    return fail(B200NUFFT_INVALID_ARGUMENT, "allocator needs both alloc and free callbacks");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(B200NUFFT_INTERNAL, std::string("no CUDA device available: ") + cudaGetErrorString(ce) +
                                        " (this engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(B200NUFFT_INVALID_ARGUMENT, "invalid device ordinal");
  DeviceGuard guard(device);

  b200nufft_plan* p = new b200nufft_plan();
  p->type = type; p->rank = rank; p->fft_sign = fft_sign; p->ntransf = num_transforms;
  p->dtype = dtype; p->is_double = dtype == B200NUFFT_COMPLEX128; p->device = device; p->tol = tol;
  p->mem.device = device;
  if (allocator) p->mem.a = *allocator;
  if (opts) p->opts = *opts; else b200nufft_default_opts(&p->opts);
  p->n_modes_tot = 1;
  for (int d = 0; d < rank; ++d) {
    if (grid_dims[d] < 1 || grid_dims[d] > 2000000000LL) {
      delete p;
      return fail(B200NUFFT_INVALID_ARGUMENT, "invalid grid dimension");
    }
    p->n_modes[d] = grid_dims[d];
    p->n_modes_tot *= grid_dims[d];
  }
  int rc = p->is_double ? create_impl<double>(p) : create_impl<float>(p);
  if (rc != B200NUFFT_OK) {
    g_create_error = p->err;
    b200nufft_plan_destroy(p);
    return rc;
  }
  *out = p;
  return B200NUFFT_OK;
}

// ---- process-level plan cache -------------------------------------------------------------------
struct PlanCache {
  std::mutex mu;
  std::list<b200nufft_plan*> idle;   // most recently released first
  int64_t hits = 0, misses = 0;
  size_t capacity = 8;
  PlanCache() {
    if (const char* e = std::getenv("B200NUFFT_PLAN_CACHE")) capacity = static_cast<size_t>(std::max(0, std::atoi(e)));
  }
};
PlanCache& plan_cache() {
  static PlanCache* c = new PlanCache();   // never destroyed: plans must not outlive the CUDA context teardown order
  return *c;
}
std::string make_cache_key(int type, int rank, const int64_t* grid_dims, int fft_sign, int num_transforms, double tol,
                           int dtype, const b200nufft_opts* opts, int device, const b200nufft_allocator* allocator) {
  b200nufft_opts o;
  if (opts) o = *opts; else b200nufft_default_opts(&o);
  std::string k;
  auto put = [&k](const void* ptr, size_t n) { k.append(static_cast<const char*>(ptr), n); };
  int64_t dims[3] = {1, 1, 1};
  for (int d = 0; d < rank && d < 3; ++d) dims[d] = grid_dims[d];
  put(&type, sizeof type); put(&rank, sizeof rank); put(dims, sizeof dims); put(&fft_sign, sizeof fft_sign);
  put(&num_transforms, sizeof num_transforms); put(&tol, sizeof tol); put(&dtype, sizeof dtype);
  put(&o, sizeof o); put(&device, sizeof device);
  b200nufft_allocator a{nullptr, nullptr, nullptr};
  if (allocator) a = *allocator;
  put(&a.alloc, sizeof a.alloc); put(&a.free, sizeof a.free); put(&a.user, sizeof a.user);
  return k;
}
}  // namespace

int b200nufft_plan_create(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims, int fft_sign,
                          int num_transforms, double tol, int dtype, const b200nufft_opts* opts, int device) {
  return plan_create_common(out, type, rank, grid_dims, fft_sign, num_transforms, tol, dtype, opts, device, nullptr);
}

int b200nufft_plan_create_ex(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims, int fft_sign,
                             int num_transforms, double tol, int dtype, const b200nufft_opts* opts, int device,
                             const b200nufft_allocator* allocator) {
  return plan_create_common(out, type, rank, grid_dims, fft_sign, num_transforms, tol, dtype, opts, device, allocator);
}

void b200nufft_plan_destroy(b200nufft_plan* p) {
  if (!p) return;
  DeviceGuard guard(p->device);
  if (p->side) { cudaStreamSynchronize(p->side); cudaStreamDestroy(p->side); }   // a pre-clear may be in flight
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_cleared) cudaEventDestroy(p->ev_cleared);
  if (p->has_fft) cufftDestroy(p->fft);
  if (p->has_fft_rem) cufftDestroy(p->fft_rem);
  if (p->fft_xy_lo) cufftDestroy(p->fft_xy_lo);
  if (p->fft_xy_hi) cufftDestroy(p->fft_xy_hi);
  if (p->fft_z) cufftDestroy(p->fft_z);
  for (DevBuf* b : p->ws_bufs()) b->release(p->mem);
  for (int d = 0; d < 3; ++d) p->fser[d].release(p->mem);
  for (int d = 0; d < 3; ++d) { p->fft_tw[d].release(p->mem); p->fft_rfac[d].release(p->mem); }
  p->misc.release(p->mem);
  p->reuse.release(p->mem);
  if (p->h_flag) cudaFreeHost(p->h_flag);
  if (p->done) cudaEventDestroy(p->done);
  for (auto& e : p->ev) if (e) cudaEventDestroy(e);
  for (auto& e : p->ev_batch) cudaEventDestroy(e);
  delete p;
}

int b200nufft_plan_acquire(b200nufft_plan** out, int type, int rank, const int64_t* grid_dims, int fft_sign,
                           int num_transforms, double tol, int dtype, const b200nufft_opts* opts, int device,
                           const b200nufft_allocator* allocator) {
  if (!out) return B200NUFFT_INVALID_ARGUMENT;
  *out = nullptr;
  if (rank < 1 || rank > 3 || !grid_dims)   // let create produce the proper message
    return plan_create_common(out, type, rank, grid_dims, fft_sign, num_transforms, tol, dtype, opts, device, allocator);
  const std::string key = make_cache_key(type, rank, grid_dims, fft_sign, num_transforms, tol, dtype, opts, device, allocator);
  PlanCache& c = plan_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    for (auto it = c.idle.begin(); it != c.idle.end(); ++it) {
      if ((*it)->cache_key == key) {
        *out = *it;
        c.idle.erase(it);
        c.hits++;
        return B200NUFFT_OK;
      }
    }
    c.misses++;
  }
  int rc = plan_create_common(out, type, rank, grid_dims, fft_sign, num_transforms, tol, dtype, opts, device, allocator);
  if (rc == B200NUFFT_OK) {
    (*out)->from_cache = true;
    (*out)->cache_key = key;
  }
  return rc;
}

void b200nufft_plan_release(b200nufft_plan* p) {
  if (!p) return;
  if (!p->from_cache || p->ws_bound) {   // a plan still tied to a caller-owned block is not kept
    if (p->ws_bound) b200nufft_unbind_workspace(p);
    if (!p->from_cache) { b200nufft_plan_destroy(p); return; }
  }
  std::vector<b200nufft_plan*> evict;
  PlanCache& c = plan_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    c.idle.push_front(p);
    while (c.idle.size() > c.capacity) {
      evict.push_back(c.idle.back());
      c.idle.pop_back();
    }
  }
  for (b200nufft_plan* e : evict) b200nufft_plan_destroy(e);
}

void b200nufft_plan_cache_clear(void) {
  std::vector<b200nufft_plan*> all;
  PlanCache& c = plan_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    all.assign(c.idle.begin(), c.idle.end());
    c.idle.clear();
  }
  for (b200nufft_plan* e : all) b200nufft_plan_destroy(e);
}

void b200nufft_plan_cache_stats(int64_t out[3]) {
  PlanCache& c = plan_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  out[0] = c.hits;
  out[1] = c.misses;
  out[2] = static_cast<int64_t>(c.idle.size());
}

size_t b200nufft_workspace_bytes(const b200nufft_plan* p, int64_t M) {
  if (!p || M < 0) return 0;
  size_t need[kNumWsBufs], total = 0;
  ws_sizes_any(p, M, need);
  for (int i = 0; i < kNumWsBufs; ++i) total += align256(need[i]);
  return total + 256;
}

int b200nufft_bind_workspace(b200nufft_plan* p, void* workspace, size_t bytes, int64_t M) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(p->mu);
  if (!workspace || M < 0) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "bind_workspace: null workspace or negative num_points");
  const size_t want = b200nufft_workspace_bytes(p, M);
  if (bytes < want)
    return set_err(p, B200NUFFT_RESOURCE_EXHAUSTED, "workspace of %zu bytes is too small: %zu needed for %lld points",
                   bytes, want, (long long)M);
  DeviceGuard guard(p->device);
  size_t need[kNumWsBufs];
  ws_sizes_any(p, M, need);
  char* cur = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255));
  auto bufs = p->ws_bufs();
  for (int i = 0; i < kNumWsBufs; ++i) {
    bufs[i]->release(p->mem);   // plan-owned buffers of this class are given back
    bufs[i]->bind(cur, align256(need[i]));
    cur += align256(need[i]);
  }
  p->ws_bound = true;
  p->ws_points = M;
  p->points_set = false;
  p->fp_valid = false;
  p->tmap_in.ok = p->tmap_out.ok = false;
  return B200NUFFT_OK;
}

int b200nufft_unbind_workspace(b200nufft_plan* p) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(p->mu);
  if (!p->ws_bound) return B200NUFFT_OK;
  for (DevBuf* b : p->ws_bufs()) b->release(p->mem);   // external: pointers dropped, nothing freed
  p->ws_bound = false;
  p->ws_points = 0;
  p->points_set = false;
  p->fp_valid = false;
  p->idx = nullptr;
  p->tmap_in.ok = p->tmap_out.ok = false;
  return B200NUFFT_OK;
}

int b200nufft_reserve(b200nufft_plan* p, int64_t M) {
  if (!p || M < 0) return B200NUFFT_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(p->mu);
  if (p->ws_bound || p->opts.external_workspace)
    return set_err(p, B200NUFFT_INVALID_ARGUMENT, "reserve: this plan takes its buffers from a caller workspace");
  DeviceGuard guard(p->device);
  size_t need[kNumWsBufs];
  ws_sizes_any(p, M, need);
  auto bufs = p->ws_bufs();
  for (int i = 1; i < kNumWsBufs; ++i) {
    if (need[i] > bufs[i]->cap) { p->fp_valid = false; p->points_set = false; }
    CUDA_OK(p, bufs[i]->reserve(p->mem, need[i], false));
  }
  return B200NUFFT_OK;
}

void b200nufft_debug_alloc_counts(int64_t* allocs, int64_t* frees) {
  if (allocs) *allocs = g_allocs.load();
  if (frees) *frees = g_frees.load();
}

int b200nufft_get_reuse_stats(b200nufft_plan* p, int64_t out[2]) {
  if (!p || !out) return B200NUFFT_INVALID_ARGUMENT;
  out[0] = out[1] = 0;
  if (!p->opts.reuse_points) return B200NUFFT_OK;
  std::lock_guard<std::mutex> lock(p->mu);
  DeviceGuard guard(p->device);
  ReuseState h;
  CUDA_OK(p, cudaDeviceSynchronize());
  CUDA_OK(p, cudaMemcpy(&h, p->reuse.p, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = h.n_skipped;
  out[1] = h.n_full;
  return B200NUFFT_OK;
}

int b200nufft_set_points(b200nufft_plan* p, int64_t M, const void* x, const void* y, const void* z, void* stream) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanCall call(p, st);
  return p->is_double ? set_points_impl<double>(p, M, 0, x, y, z, st) : set_points_impl<float>(p, M, 0, x, y, z, st);
}

int b200nufft_set_points_interleaved(b200nufft_plan* p, int64_t M, const void* pts, void* stream) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanCall call(p, st);
  return p->is_double ? set_points_impl<double>(p, M, 1, pts, nullptr, nullptr, st)
                      : set_points_impl<float>(p, M, 1, pts, nullptr, nullptr, st);
}

int b200nufft_execute(b200nufft_plan* p, void* c, void* f, void* stream) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  if (p->opts.spread_only) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "execute called on a spread-only plan");
  if (!p->points_set) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "set_points must be called before execute");
  if (!p->fine.p) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "execute: no workspace bound (external_workspace plan)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanCall call(p, st);
  return p->is_double ? execute_impl<double>(p, c, f, st) : execute_impl<float>(p, c, f, st);
}

int b200nufft_interp(b200nufft_plan* p, void* c, const void* f, void* stream) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  if (!p->opts.spread_only) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "interp needs a spread-only plan");
  if (!p->points_set) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "set_points must be called before interp");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanCall call(p, st);
  const size_t cs = p->is_double ? sizeof(double2) : sizeof(float2);
  for (int b0 = 0; b0 < p->ntransf; b0 += p->batch) {
    const int ntr = std::min(p->batch, p->ntransf - b0);
    char* cb = static_cast<char*>(c) + cs * b0 * p->M;
    const char* fb = static_cast<const char*>(f) + cs * b0 * p->nftot;
    int rc = p->is_double ? do_interp<double>(p, ntr, fb, cb, st) : do_interp<float>(p, ntr, fb, cb, st);
    if (rc) return rc;
    const int64_t n = static_cast<int64_t>(ntr) * p->M;
    if (n > 0) {
      if (p->is_double) scale_kernel<double><<<grid_for(n, 256), 256, 0, st>>>(n, p->kernel_scale, reinterpret_cast<double2*>(cb));
      else scale_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(n, static_cast<float>(p->kernel_scale), reinterpret_cast<float2*>(cb));
      p->launches++;
    }
  }
  LAUNCH_OK(p);
  return B200NUFFT_OK;
}

int b200nufft_spread(b200nufft_plan* p, const void* c, void* f, void* stream) {
  if (!p) return B200NUFFT_INVALID_ARGUMENT;
  if (!p->opts.spread_only) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "spread needs a spread-only plan");
  if (!p->points_set) return set_err(p, B200NUFFT_INVALID_ARGUMENT, "set_points must be called before spread");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlanCall call(p, st);
  const size_t cs = p->is_double ? sizeof(double2) : sizeof(float2);
  for (int b0 = 0; b0 < p->ntransf; b0 += p->batch) {
    const int ntr = std::min(p->batch, p->ntransf - b0);
    const char* cb = static_cast<const char*>(c) + cs * b0 * p->M;
    char* fb = static_cast<char*>(f) + cs * b0 * p->nftot;
    CUDA_OK(p, cudaMemsetAsync(fb, 0, cs * p->nftot * ntr, st));
    int rc = p->is_double ? do_spread<double>(p, ntr, cb, fb, st) : do_spread<float>(p, ntr, cb, fb, st);
    if (rc) return rc;
    const int64_t n = static_cast<int64_t>(ntr) * p->nftot;
    if (p->is_double) scale_kernel<double><<<grid_for(n, 256), 256, 0, st>>>(n, p->kernel_scale, reinterpret_cast<double2*>(fb));
    else scale_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(n, static_cast<float>(p->kernel_scale), reinterpret_cast<float2*>(fb));
    p->launches++;
  }
  LAUNCH_OK(p);
  return B200NUFFT_OK;
}

int b200nufft_get_sort(const b200nufft_plan* p, const int32_t** idx, const int32_t** bin_start,
                       const int32_t** bin_sizes, int32_t* bin_count) {
  if (!p || !p->points_set) return B200NUFFT_INVALID_ARGUMENT;
  if (idx) *idx = p->idx;
  if (bin_start) *bin_start = p->bin_start.as<int32_t>();
  if (bin_sizes) *bin_sizes = p->bin_sizes.as<int32_t>();
  if (bin_count) *bin_count = p->nbtot;
  return B200NUFFT_OK;
}

int b200nufft_binsort(int is_double, int rank, int64_t M, const void* x, const void* y, const void* z,
                      const int* fine_dims, const int* bin_dims, int rounding, int32_t* idx_out,
                      int32_t* bin_start_out, int32_t* bin_sizes_out, void* stream) {
  if (rank < 1 || rank > 3 || M < 0) return B200NUFFT_INVALID_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BinGeom bg{};
  bg.rank = rank;
  bg.rounding = rounding;
  int nbtot = 1;
  for (int d = 0; d < 3; ++d) {
    bg.nf[d] = d < rank ? fine_dims[d] : 1;
    bg.bin[d] = d < rank ? bin_dims[d] : 1;
    bg.nbins[d] = d < rank ? (rounding == 0 ? (bg.nf[d] + bg.bin[d] - 1) / bg.bin[d] : bg.nf[d] / bg.bin[d] + 1) : 1;
    nbtot *= bg.nbins[d];
  }
  if (cudaMemsetAsync(bin_sizes_out, 0, sizeof(int) * nbtot, st) != cudaSuccess) return B200NUFFT_INTERNAL;
  DevBuf k0, pa, pb, v1, hist, tmp;
  MemCtx mem;
  cudaGetDevice(&mem.device);
  int rc = B200NUFFT_OK;
  if (M > 0) {
    if (k0.reserve(mem, 4 * M) || pa.reserve(mem, 8 * M) || pb.reserve(mem, 8 * M) || v1.reserve(mem, 4 * M) ||
        hist.reserve(mem, sizeof(int) * radix_hist_ints(M)) || tmp.reserve(mem, sizeof(int) * (kScanMaxBlocks + 8))) {
      rc = B200NUFFT_RESOURCE_EXHAUSTED;
    } else {
      if (is_double)
        key_only_kernel<double><<<grid_for(M, 256), 256, 0, st>>>(M, (const double*)x, (const double*)y, (const double*)z, bg,
                                                                 k0.as<uint32_t>(), bin_sizes_out);
      else
        key_only_kernel<float><<<grid_for(M, 256), 256, 0, st>>>(M, (const float*)x, (const float*)y, (const float*)z, bg,
                                                                k0.as<uint32_t>(), bin_sizes_out);
      radix_sort_index(k0.as<uint32_t>(), pa.as<uint2>(), pb.as<uint2>(), v1.as<int>(), M, ilog2_ceil(nbtot),
                       hist.as<int>(), tmp.as<int>(), st);
      cudaMemcpyAsync(idx_out, v1.as<int>(), sizeof(int) * M, cudaMemcpyDeviceToDevice, st);
    }
  } else {
    tmp.reserve(mem, sizeof(int) * (kScanMaxBlocks + 8));
  }
  if (rc == B200NUFFT_OK) {
    exclusive_scan_i32(bin_sizes_out, bin_start_out, nbtot, tmp.as<int>(), nullptr, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = B200NUFFT_INTERNAL;
  }
  k0.release(mem); pa.release(mem); pb.release(mem); v1.release(mem); hist.release(mem); tmp.release(mem);
  return rc;
}

int b200nufft_fold_rescale(int is_double, int points_range, int64_t M, const void* in, void* out, int fine_dim,
                           void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0) return B200NUFFT_OK;
  if (is_double)
    fold_only_kernel<double><<<grid_for(M, 256), 256, 0, st>>>(M, (const double*)in, (double*)out, points_range, fine_dim);
  else
    fold_only_kernel<float><<<grid_for(M, 256), 256, 0, st>>>(M, (const float*)in, (float*)out, points_range, fine_dim);
  return cudaGetLastError() == cudaSuccess ? B200NUFFT_OK : B200NUFFT_INTERNAL;
}

int b200nufft_copy_to_host(void* dst_host, const void* src_device, size_t bytes) {
  return cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? B200NUFFT_OK
                                                                                         : B200NUFFT_INTERNAL;
}

int b200nufft_get_info(const b200nufft_plan* p, b200nufft_info* info) {
  if (!p || !info) return B200NUFFT_INVALID_ARGUMENT;
  std::memset(info, 0, sizeof(*info));
  info->kernel_width = p->kp.ns;
  info->kernel_beta = p->kp.beta;
  info->kernel_c = p->kp.c;
  info->upsampling_factor = p->kp.sigma;
  info->kernel_scale = p->kernel_scale;
  for (int d = 0; d < 3; ++d) { info->fine_dims[d] = p->nf[d]; info->bin_dims[d] = p->bin[d]; info->num_bins[d] = p->nbins[d]; }
  info->batch_size = p->batch;
  info->num_threads_compat = p->num_threads_compat;
  info->num_points = p->M;
  info->subproblem_bound = p->sub_bound;
  info->spread_method = (p->type == 1 || p->opts.spread_only) ? p->spread_method : 0;
  info->interp_method = (p->type == 2 || p->opts.spread_only) ? p->interp_method : 0;
  info->fft_method = p->opts.spread_only ? 0 : (p->own_fft ? 3 : (p->pruned_fft ? 2 : 1));
  return B200NUFFT_OK;
}

int b200nufft_get_fseries(const b200nufft_plan* p, int dim, void* host_out) {
  if (!p || dim < 0 || dim >= p->rank || p->fser_host[dim].empty()) return B200NUFFT_INVALID_ARGUMENT;
  std::memcpy(host_out, p->fser_host[dim].data(), p->fser_host[dim].size());
  return B200NUFFT_OK;
}

int b200nufft_get_timings(b200nufft_plan* p, float out[4]) {
  if (!p || !p->opts.profile) return B200NUFFT_INVALID_ARGUMENT;
  for (int i = 0; i < 4; ++i) out[i] = 0.f;
  if (p->ev_exec) {
    // Sum over the batches of the last execute.
    for (int bi = 0; bi < p->ev_batches; ++bi) {
      cudaEvent_t* ev = p->ev_batch.data() + 4 * bi;
      if (cudaEventSynchronize(ev[3]) != cudaSuccess) return B200NUFFT_INTERNAL;
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, ev[0], ev[1]);
      cudaEventElapsedTime(&b, ev[1], ev[2]);
      cudaEventElapsedTime(&c, ev[2], ev[3]);
      if (p->type == 1) { out[0] += a; out[1] += b; out[2] += c; }
      else { out[2] += a; out[1] += b; out[0] += c; }
    }
  }
  if (p->ev_setpts) {
    if (cudaEventSynchronize(p->ev[5]) != cudaSuccess) return B200NUFFT_INTERNAL;
    cudaEventElapsedTime(&out[3], p->ev[4], p->ev[5]);
  }
  return B200NUFFT_OK;
}

int64_t b200nufft_launch_count(const b200nufft_plan* p) { return p ? p->launches : 0; }
const char* b200nufft_last_error(const b200nufft_plan* p) { return p ? p->err : "null plan"; }
const char* b200nufft_last_create_error(void) { return g_create_error.c_str(); }

int b200nufft_host_kernel_width(int is_double, double tol, double sigma) {
  return is_double ? kernel_width_from_tol<double>(tol, sigma) : kernel_width_from_tol<float>(static_cast<float>(tol), sigma);
}
int b200nufft_host_next_smooth_int(int n) { return next_smooth_int(n); }
int b200nufft_host_fseries(int is_double, int fine_dim, int kernel_width, int mode, int num_threads, void* out) {
  if (kernel_width < 2 || kernel_width > kMaxKernelWidth || fine_dim < 2) return B200NUFFT_INVALID_ARGUMENT;
  if (is_double)
    kernel_fseries<double>(fine_dim, kernel_params_from_width<double>(kernel_width, 2.0), mode, num_threads,
                           static_cast<double*>(out));
  else
    kernel_fseries<float>(fine_dim, kernel_params_from_width<float>(kernel_width, 2.0), mode, num_threads,
                          static_cast<float*>(out));
  return B200NUFFT_OK;
}
double b200nufft_host_scale_factor(int is_double, int rank, int kernel_width) {
  if (is_double) return kernel_scale_factor<double>(rank, kernel_params_from_width<double>(kernel_width, 2.0));
  return kernel_scale_factor<float>(rank, kernel_params_from_width<float>(kernel_width, 2.0));
}
int b200nufft_host_gauss_legendre(int n, double* nodes, double* weights) {
  if (n < 1) return B200NUFFT_INVALID_ARGUMENT;
  gauss_legendre(n, nodes, weights);
  return B200NUFFT_OK;
}
const char* b200nufft_version(void) { return "b200nufft 0.1.0 (sm_100a)"; }

}  // extern "C"
