// scan_sort.cuh -- hand-written device primitives for the bin-sort: an exclusive scan of int32
// arrays and a stable LSD radix sort of (key, value) pairs. Deterministic and stable: the bin-sort
// permutation must be bit-reproducible (the reference's is atomic-arrival ordered,
// nufft_plan.cu.cc:171,196,228; the stable order is the one it produces when atomics resolve
// in index order).
#pragma once
#include "dev_common.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Exclusive scan, three phases: per-block chunk sums -> scan of <=1024 block sums -> per-block
// sequential tile scan with carry. out may alias in. total (optional) receives the grand total.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // 1024
constexpr int kScanMaxBlocks = 1024;

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one int per thread (256 threads); returns exclusive prefix, and
// the block total through *total_out (same for all threads).
__device__ __forceinline__ int block_excl_scan_256(int v, int* total_out) {
  __shared__ int warp_sums[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int wprefix = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    int s = warp_sums[w];
    if (w < warp) wprefix += s;
    total += s;
  }
  __syncthreads();
  *total_out = total;
  return wprefix + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_chunk_sums_kernel(const int* __restrict__ in, int64_t n, int64_t chunk, int* __restrict__ block_sums,
                       const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const int64_t beg = static_cast<int64_t>(blockIdx.x) * chunk;
  const int64_t end = min(beg + chunk, n);
  int s = 0;
  for (int64_t i = beg + threadIdx.x; i < end; i += kScanThreads) s += in[i];
  int total;
  block_excl_scan_256(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanMaxBlocks)
scan_block_sums_kernel(int* __restrict__ block_sums, int nb, int* __restrict__ total_out,
                       const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  __shared__ int wsum[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int v = t < nb ? block_sums[t] : 0;
  int incl = warp_incl_scan(v);
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int s = wsum[lane];
    int si = warp_incl_scan(s);
    wsum[lane] = si - s;
  }
  __syncthreads();
  int excl = wsum[warp] + incl - v;
  if (t < nb) block_sums[t] = excl;
  if (t == nb - 1 && total_out) *total_out = excl + v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, int64_t n, int64_t chunk,
                  const int* __restrict__ block_offs, const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const int64_t beg = static_cast<int64_t>(blockIdx.x) * chunk;
  const int64_t end = min(beg + chunk, n);
  int carry = block_offs[blockIdx.x];
  for (int64_t base = beg; base < end; base += kScanTile) {
    int v[kScanItems];
    int tsum = 0;
    const int64_t i0 = base + static_cast<int64_t>(threadIdx.x) * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      v[k] = (i0 + k < end) ? in[i0 + k] : 0;
      tsum += v[k];
    }
    int total;
    int excl = block_excl_scan_256(tsum, &total) + carry;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (i0 + k < end) out[i0 + k] = excl;
      excl += v[k];
    }
    carry += total;
  }
}

// Single-CTA scan for small arrays (bins of a 2D plan, radix histograms of small point sets): one
// launch instead of three. Optional element transform: v -> ceil(v / div) when div > 0
// (subproblem counts, CalcSubproblemKernel nufft_plan.cu.cc:304-310, fused into the scan).
constexpr int kScanSmallMax = 32768;
__global__ void __launch_bounds__(1024)
scan_small_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int div, int* __restrict__ total_out,
                  const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    int v = i < n ? in[i] : 0;
    if (div > 0) v = (v + div - 1) / div;
    int incl = warp_incl_scan(v);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int s = wsum[lane];
      int si = warp_incl_scan(s);
      wsum[lane] = si - s;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + wsum[warp] + incl - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (t == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (t == 0 && total_out) *total_out = carry_s;
}

// tmp must hold kScanMaxBlocks ints. Returns the number of kernels launched.
// skip (optional, device): when *skip != 0 every kernel returns at once (opts.reuse_points).
inline int exclusive_scan_i32(const int* in, int* out, int64_t n, int* tmp, int* total_out,
                              cudaStream_t stream, const int* skip = nullptr) {
  if (n <= 0) {
    if (total_out) cudaMemsetAsync(total_out, 0, sizeof(int), stream);
    return 0;
  }
  if (n <= kScanSmallMax) {
    scan_small_kernel<<<1, 1024, 0, stream>>>(in, out, static_cast<int>(n), 0, total_out, skip);
    return 1;
  }
  int nb = static_cast<int>(std::min<int64_t>(kScanMaxBlocks, (n + 4 * kScanTile - 1) / (4 * kScanTile)));
  int64_t chunk = (n + nb - 1) / nb;
  chunk = ((chunk + kScanTile - 1) / kScanTile) * kScanTile;
  nb = static_cast<int>((n + chunk - 1) / chunk);
  scan_chunk_sums_kernel<<<nb, kScanThreads, 0, stream>>>(in, n, chunk, tmp, skip);
  scan_block_sums_kernel<<<1, kScanMaxBlocks, 0, stream>>>(tmp, nb, total_out, skip);
  scan_apply_kernel<<<nb, kScanThreads, 0, stream>>>(in, out, n, chunk, tmp, skip);
  return 3;
}

// ---------------------------------------------------------------------------------------------
// Stable LSD radix sort of the INDICES 0 .. n-1 by 32-bit keys, 8- or 10-bit digits. Tile = 8 warps
// x 256 consecutive elements per warp; a warp walks its 256 elements in 8 rounds of 32 lanes, so
// the (warp, round, lane) order IS the element order and ranks are stable by construction.
// Data movement per pass (round 2): the first pass reads only the keys (the value of element i is
// i), passes in between move (key, index) as ONE 8-byte pair -- one scattered store per element
// instead of two -- and the last pass writes only the index (the sorted keys are never read).
// ---------------------------------------------------------------------------------------------
constexpr int kSortWarps = 8;
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kSortRounds = 8;
constexpr int kSortPerWarp = 32 * kSortRounds;             // 256
constexpr int kSortTile = kSortWarps * kSortPerWarp;       // 2048
constexpr int kRadixBitsMax = 11;                          // 8-bit digits, or 10 / 11-bit when that saves a pass

// IN_PAIRS: the pass reads (key, index) pairs, else bare keys
template <int BITS, bool IN_PAIRS>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, const uint2* __restrict__ pairs, int64_t n, int shift, int nblk,
                  int* __restrict__ hist /*[1 << BITS][nblk]*/, const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  constexpr int RADIX = 1 << BITS;
  __shared__ int cnt[RADIX];
  for (int d = threadIdx.x; d < RADIX; d += kSortThreads) cnt[d] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortTile / kSortThreads; ++r) {
    int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&cnt[((IN_PAIRS ? pairs[i].x : keys[i]) >> shift) & (RADIX - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += kSortThreads) hist[static_cast<int64_t>(d) * nblk + blockIdx.x] = cnt[d];
}

// OUT_PAIRS: the pass writes (key, index) pairs, else (last pass) the indices only
template <int BITS, bool IN_PAIRS, bool OUT_PAIRS>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint2* __restrict__ pairs_in,
                     uint2* __restrict__ pairs_out, int* __restrict__ idx_out, int64_t n,
                     int shift, int nblk, const int* __restrict__ offs /*[1 << BITS][nblk] scanned*/,
                     const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  constexpr int RADIX = 1 << BITS;
  // running per-warp digit counts; a tile holds 2048 elements, so 16 bits are enough (and keep the
  // 11-bit variant inside the 48 KB of static shared memory)
  __shared__ unsigned short cnt[kSortWarps][RADIX];
  __shared__ int gbase[RADIX];             // global base of each digit for this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * RADIX; i += kSortThreads) (&cnt[0][0])[i] = 0;
  for (int d = threadIdx.x; d < RADIX; d += kSortThreads) gbase[d] = offs[static_cast<int64_t>(d) * nblk + blockIdx.x];
  __syncthreads();

  const int64_t wbase = static_cast<int64_t>(blockIdx.x) * kSortTile + warp * kSortPerWarp;
  uint32_t key[kSortRounds];
  int val[kSortRounds];
  int rank[kSortRounds];
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < n;
    if (IN_PAIRS) {
      const uint2 kv = ok ? pairs_in[i] : make_uint2(0xffffffffu, 0u);
      key[r] = kv.x;
      val[r] = static_cast<int>(kv.y);
    } else {
      key[r] = ok ? keys_in[i] : 0xffffffffu;
      val[r] = static_cast<int>(i);
    }
    // Invalid lanes use digit RADIX (out of range) so they never match a real digit.
    const unsigned dig = ok ? ((key[r] >> shift) & (RADIX - 1)) : RADIX;
    const unsigned peers = __match_any_sync(0xffffffffu, dig);
    const int before = __popc(peers & ((1u << lane) - 1u));
    int prev = 0;
    if (ok) prev = cnt[warp][dig];
    __syncwarp();
    if (ok && before == 0) cnt[warp][dig] = static_cast<unsigned short>(prev + __popc(peers));
    __syncwarp();
    rank[r] = prev + before;
  }
  __syncthreads();
  // Exclusive prefix over warps for each digit.
  for (int d = threadIdx.x; d < RADIX; d += kSortThreads) {
    int run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      int c = cnt[w][d];
      cnt[w][d] = static_cast<unsigned short>(run);
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const unsigned dig = (key[r] >> shift) & (RADIX - 1);
      const int64_t pos = static_cast<int64_t>(gbase[dig]) + cnt[warp][dig] + rank[r];
      if (OUT_PAIRS) pairs_out[pos] = make_uint2(key[r], static_cast<uint32_t>(val[r]));
      else idx_out[pos] = val[r];
    }
  }
}

__global__ void __launch_bounds__(256)
iota_kernel(int* __restrict__ out, int64_t n, const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = static_cast<int>(i);
}

template <int BITS>
inline void radix_pass(const uint32_t* keys, const uint2* pin, uint2* pout, int* idx_out, bool first, bool last,
                       int64_t n, int shift, int nblk, int* hist, int* scan_tmp, cudaStream_t stream, const int* skip,
                       int* launches) {
  if (first) radix_hist_kernel<BITS, false><<<nblk, kSortThreads, 0, stream>>>(keys, nullptr, n, shift, nblk, hist, skip);
  else radix_hist_kernel<BITS, true><<<nblk, kSortThreads, 0, stream>>>(nullptr, pin, n, shift, nblk, hist, skip);
  *launches += 1 + exclusive_scan_i32(hist, hist, (static_cast<int64_t>(1) << BITS) * nblk, scan_tmp, nullptr, stream, skip);
  if (first && last) radix_scatter_kernel<BITS, false, false><<<nblk, kSortThreads, 0, stream>>>(keys, nullptr, nullptr, idx_out, n, shift, nblk, hist, skip);
  else if (first) radix_scatter_kernel<BITS, false, true><<<nblk, kSortThreads, 0, stream>>>(keys, nullptr, pout, nullptr, n, shift, nblk, hist, skip);
  else if (last) radix_scatter_kernel<BITS, true, false><<<nblk, kSortThreads, 0, stream>>>(nullptr, pin, nullptr, idx_out, n, shift, nblk, hist, skip);
  else radix_scatter_kernel<BITS, true, true><<<nblk, kSortThreads, 0, stream>>>(nullptr, pin, pout, nullptr, n, shift, nblk, hist, skip);
  *launches += 1;
}

// idx_out[j] = index of the element with the j-th smallest key (low `key_bits` bits; ties keep the
// element order). pa / pb: scratch for n (key, index) pairs each (pb only when there are more than
// two passes); hist holds radix_hist_ints(n) ints; scan_tmp kScanMaxBlocks ints. Returns the number
// of kernels launched.
inline int radix_sort_index(const uint32_t* keys, uint2* pa, uint2* pb, int* idx_out, int64_t n, int key_bits,
                            int* hist, int* scan_tmp, cudaStream_t stream, const int* skip = nullptr) {
  int launches = 0;
  if (n <= 0) return 0;
  const int nblk = static_cast<int>((n + kSortTile - 1) / kSortTile);
  // 8-bit digits unless 10-bit digits save a whole pass (cfg2: 20 key bits -> 2 passes of 10).
  // 11-bit digits (instantiated for the histogram sizing only, off): cfg3's 22-bit window keys in 2
  // passes instead of 3 were measured SLOWER (set_points 1.26 vs 1.16 ms): the 2048-entry per-block
  // tables and the 8x larger histogram scan cost more than the pass they save.
  int bits = 8;
  int passes = key_bits <= 0 ? 0 : (key_bits + 7) / 8;
  {
    const int pb10 = key_bits <= 0 ? 0 : (key_bits + 9) / 10;
    if (pb10 < passes) { passes = pb10; bits = 10; }
  }
  if (passes == 0) {
    iota_kernel<<<static_cast<int>(std::min<int64_t>((n + 255) / 256, 4096)), 256, 0, stream>>>(idx_out, n, skip);
    return 1;
  }
  const uint2* pin = nullptr;
  uint2* pout = pa;
  for (int p = 0; p < passes; ++p) {
    const bool first = p == 0, last = p == passes - 1;
    if (bits == 10) radix_pass<10>(keys, pin, pout, idx_out, first, last, n, p * bits, nblk, hist, scan_tmp, stream, skip, &launches);
    else radix_pass<8>(keys, pin, pout, idx_out, first, last, n, p * bits, nblk, hist, scan_tmp, stream, skip, &launches);
    pin = pout;
    pout = (pout == pa) ? pb : pa;
  }
  return launches;
}

inline int64_t radix_hist_ints(int64_t n) {
  return (static_cast<int64_t>(1) << kRadixBitsMax) * ((n + kSortTile - 1) / kSortTile) + 1;
}

}  // namespace b200
