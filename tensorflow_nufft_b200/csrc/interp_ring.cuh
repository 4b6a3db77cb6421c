// interp_ring.cuh -- 3D type-2 interpolator with Z-SLAB STREAMING (complex64, ns <= 7).
//
// Same sums and the same quarter-warp gather as interp_qw.cuh (reference: InterpSubproblem3DKernel
// nufft_plan.cu.cc:1608-1706). interp_qw stages the whole (bin + halo) tile of a subproblem -- all
// bin_z + 8 planes it may need, 53 KB for 16 x 8 x 8 bins -- waits for it, then gathers: 4 CTAs per
// SM, every one of them idle while its tile is in flight (42 % of the cfg4 stall samples sit on the
// tile mbarrier). Here the points of a bin are sorted by the z start of their stencil, and a CTA
// keeps only a RING of 8 z-planes in shared memory (27 KB: 8 CTAs per SM): it gathers the points
// whose stencils start at plane rz while the TMA unit already loads the planes the next starts
// need into the slots the previous ones left behind. The bin depth no longer costs shared memory,
// so bins are 16 x 8 x 16: less z-halo per point.
#pragma once
#include <cuda.h>
#include <cuda_pipeline.h>

#include "dev_common.cuh"
#include "interp.cuh"
#include "interp_qw.cuh"
#include "spread.cuh"

namespace b200 {

constexpr int kInterpRing = 8;      // resident z-planes
constexpr int kRingMaxZ = 64;       // stencil z starts per tile the group table can hold (bin_z + 9 - ns <= 64)

inline size_t interp_ring_smem_bytes(const int* bin) {
  const size_t plane = static_cast<size_t>(bin[0] + kQwHaloX) * (bin[1] + 8);
  return kInterpRing * plane * sizeof(float2) + kInterpRing * sizeof(uint64_t) + (kRingMaxZ + 1) * sizeof(int);
}

// One CTA (WARPS warps) per (subproblem, transform). Points [p0, p0 + np) of the subproblem are
// sorted by rz = stencil z start relative to the tile, clamped to [0, WZ) (sort key (bin, rz) of
// fold_key_kernel, WZ = bin_z + 9 - ns); the group table below uses the same clamp, and a point
// whose true start differs from its group (coordinates outside the declared points_range) is
// written as zero, like in interp_qw.cuh.
template <int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
interp_ring3d_f32_kernel(int64_t M, GridGeom g, int ntr, const int* __restrict__ sub_total,
                         const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                         const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][6]*/,
                         const float2* __restrict__ fw, float2* __restrict__ c,
                         const __grid_constant__ CUtensorMap tmap, int use_tma, int zrange) {
  constexpr int RING = kInterpRing;
  constexpr int C4 = 6;
  extern __shared__ __align__(128) float4 smem4[];
  const int s = blockIdx.x / ntr;
  const int t = blockIdx.x - s * ntr;
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + kQwHaloX, TY = g.bin[1] + 8, TZ = g.bin[2] + 8;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = b / (g.nbins[0] * g.nbins[1]);
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = bz * g.bin[2] - 4;
  const int plane4 = TX * TY / 2;
  const int TXH = TX / 2;
  float4* ring = smem4;                                                       // [RING][plane4]; tile plane z in slot z & 7
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem4 + RING * plane4);        // one mbarrier per slot
  int* goff = reinterpret_cast<int*>(bars + RING);                           // goff[rz] = first point (relative) with start rz

  const float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;
  float2* ct = c + static_cast<int64_t>(t) * M;
  const bool interior = use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
                        oz >= 0 && oz + TZ <= g.nf[2];
  const int WZ = min(TZ - NS + 1, kRingMaxZ);   // admissible stencil starts 0 .. WZ - 1

  // ---- plane loads ----
  // mbarrier phases per slot (every thread keeps the same masks): next_par = parity the slot's next
  // load will complete, cur_par = parity of its latest load
  unsigned next_par = 0, cur_par = 0, used = 0;
  auto issue_plane = [&](int z) {   // tile plane z -> slot z & 7
    float4* dst = ring + (z & (RING - 1)) * plane4;
    const unsigned bit = 1u << (z & (RING - 1));
    // a look-ahead plane that no group needed may still be in flight: its phase must complete
    // before the barrier is armed again
    if (interior && tid == 0 && (used & bit)) mbar_wait(bars + (z & (RING - 1)), (cur_par >> (z & (RING - 1))) & 1u);
    used |= bit;
    cur_par = (cur_par & ~bit) | (next_par & bit);
    next_par ^= bit;
    if (interior) {
      if (tid == 0) {
        mbar_expect_tx(bars + (z & (RING - 1)), static_cast<uint32_t>(plane4 * sizeof(float4)));
        tma_load_4d(dst, &tmap, bars + (z & (RING - 1)), 2 * ox, oy, oz + z, t);
      }
    } else {
      const int gz = mod_idx(oz + z, g.nf[2]);
      for (int i = tid; i < plane4; i += WARPS * 32) {
        const int gx = mod_idx(ox + 2 * (i % TXH), g.nf[0]);
        const int gy = mod_idx(oy + i / TXH, g.nf[1]);
        __pipeline_memcpy_async(dst + i, fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx, 16);
      }
    }
  };

  for (int i = tid; i < kRingMaxZ + 1; i += WARPS * 32) goff[i] = 0;
  if (interior && tid < RING) mbar_init(bars + tid, 1);
  __syncthreads();
  // The first ring of planes leaves BEFORE the group table is built: the smallest stencil start of
  // the subproblem is already in its descriptor (subproblem_zrange_kernel), so the tile's first
  // 8 planes and the table's pass over start[] are one global round trip, not two.
  int issued_hi = 0;   // planes below this have been issued (monotone)
  if (zrange) {
    const int lo = (sd.w & 0xffff) - 32768 - oz;
    if (lo >= 0 && lo < WZ) {
      for (int z = lo; z < min(lo + RING, TZ); ++z) issue_plane(z);
      issued_hi = min(lo + RING, TZ);
    }
  }

  // ---- group table: points per stencil start (the points are sorted by it) ----
  for (int i = tid; i < np; i += WARPS * 32) {
    const int rz = start[p0 + i].z - oz;
    atomicAdd(&goff[min(max(rz, 0), WZ - 1) + 1], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 1; i <= WZ; ++i) { run += goff[i]; goff[i] = run; }   // goff[rz + 1] = end of group rz
  }
  __syncthreads();

  const int pt = lane >> 3;
  const int row = lane & 7;
  for (int rz = 0; rz < WZ; ++rz) {
    const int gbeg = goff[rz], gend = goff[rz + 1];
    if (gend == gbeg) continue;   // uniform over the CTA
    // every warp is done with the previous group: the planes below rz are dead, their slots free
    __syncthreads();
    const int want_hi = min(rz + RING, TZ);
    for (int z = max(issued_hi, rz); z < want_hi; ++z) issue_plane(z);
    issued_hi = max(issued_hi, want_hi);
    if (interior) {
      for (int z = rz; z < rz + NS; ++z) mbar_wait(bars + (z & (RING - 1)), (cur_par >> (z & (RING - 1))) & 1u);
    } else {
      __pipeline_commit();
      __pipeline_wait_prior(0);
      __syncthreads();
    }

    // ---- quarter-warp gather of the group's points (4 points per warp step) ----
    const int n = gend - gbeg;
    const int ngrp = (n + 3) >> 2;
    for (int grp = warp; grp < ngrp; grp += WARPS) {
      const int p = 4 * grp + pt;
      const bool valid = p < n;
      float re = 0.f, im = 0.f;
      int id = 0;
      if (valid) {
        const int64_t j = static_cast<int64_t>(p0) + gbeg + p;
        const float4 xa = wrec4[j * C4], xb = wrec4[j * C4 + 1];
        const float wy = reinterpret_cast<const float*>(wrec4 + j * C4 + 2)[row];
        const float4 za = wrec4[j * C4 + 4], zb4 = wrec4[j * C4 + 5];
        const int4 st = start[j];
        id = idx[j];
        const int rx = st.x - ox, ry = st.y - oy;
        const bool fits = rx >= 0 && rx + 8 <= TX && ry >= 0 && ry + NS <= TY && st.z - oz == rz;
        if (fits && row < NS) {
          const float wz[8] = {za.x, za.y, za.z, za.w, zb4.x, zb4.y, zb4.z, zb4.w};
          const int rowoff = ((ry + row) * TX + rx) >> 1;
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) {
            const float4* pz = ring + ((rz + dz) & (RING - 1)) * plane4 + rowoff;
            const float4 v0 = pz[0], v1 = pz[1], v2 = pz[2], v3 = pz[3];
            float pr, pi;
            qw_row_dot(v0, v1, v2, v3, xa, xb, pr, pi);
            qw_fma2(re, im, wz[dz], pr, pi);
          }
          re *= wy;
          im *= wy;
        }
      }
      float part[2] = {re, im};
      const float sum = qw_reduce<2>(part, lane);   // rows 0..3 of a quarter hold re, rows 4..7 im
      if (valid && (row & 3) == 0) reinterpret_cast<float*>(ct + id)[row >> 2] = sum;
    }
  }
}

}  // namespace b200
