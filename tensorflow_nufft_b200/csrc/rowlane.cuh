// rowlane.cuh -- shared-memory tile kernels in 2D and 3D for the cases the float tile kernels do
// not cover: complex128 (any kernel width <= 15) and complex64 with widths 8..15 (tol < 1e-6, or
// the low-upsampling mode sigma = 1.25). In 3D a lane walks its row through the ns z-planes of the
// stencil (the tile is bin + halo in all three dimensions, staged with wrapped asynchronous copies).
//
// Same sums as spread.cuh / interp.cuh (reference: SpreadSubproblem2DKernel / InterpSubproblem2DKernel
// nufft_plan.cu.cc:790-878, 1041-1110) (tol 1e-12 -> ns = 14: 196 cells of 16 bytes per point).
// Layout: one lane per ROW of the stencil window, each lane walks PXT cells of its row, one
// complex cell per access (128-bit for complex128, 64-bit for complex64); the tile pitch is ODD
// in cells, so the row-lanes of a quarter (half) warp hit distinct bank groups. The weight record is zero-padded to PXT x LP, so no lane needs the width.
//   * interp: LP lanes per point, 32 / LP points per warp instruction, shuffle reduction over the
//     rows, weights read straight from the sorted record array; tile staged by TMA (interior) or
//     wrapped cp.async.
//   * spread: one point per warp step, lanes = LP rows x 2 halves of the row; private tile per warp
//     (one-warp CTAs), plain load / DFMA / store, flushed with native f64 global reductions.
// Compared with the point-driven global kernels (REDG.F64 per cell / L2 gathers) this is the same
// move the float kernels make: the ns^2 cell updates stay on chip.
#pragma once
#include <cuda.h>
#include <cuda_pipeline.h>

#include "dev_common.cuh"
#include "interp.cuh"
#include "spread.cuh"

namespace b200 {

struct RowLaneGeom {
  int TX, TY, TZ; // tile extent in cells (TX odd; TZ = 1 in 2D)
  int hx, hy, hz; // tile origin = bin origin - (hx, hy, hz)
  int R, PX, PY;  // record stride, wy offset and wy / wz length, in reals (rows >= PY carry no weight)
  int ns;         // kernel width (z taps walked in 3D)
};

// align_x: the records' x start is moved down to an even cell (complex64): one more halo cell.
inline RowLaneGeom rowlane_geom(const int* bin, int rank, int ns, int pxt, int lp, int R, int PX, int PY, int align_x) {
  RowLaneGeom r;
  r.hx = (ns + 1) / 2 + (align_x ? 1 : 0);
  r.hy = (ns + 1) / 2;
  r.hz = rank > 2 ? (ns + 1) / 2 : 0;
  r.TX = (bin[0] + pxt + 2 + (align_x ? 2 : 0)) | 1;
  r.TY = bin[1] + lp + 2;
  r.TZ = rank > 2 ? bin[2] + ns + 1 : 1;
  r.R = R;
  r.PX = PX;
  r.PY = PY;
  r.ns = ns;
  return r;
}

inline size_t rowlane_smem_bytes(const RowLaneGeom& r, size_t cell_bytes) {
  return ((static_cast<size_t>(r.TX) * r.TY * r.TZ * cell_bytes + 127) & ~static_cast<size_t>(127)) + 16;
}

// ------------------------------------------------------------------------------------------------
// type 2
// ------------------------------------------------------------------------------------------------
template <typename F, int PXT, int LP, int WARPS, int RANK>
__global__ void __launch_bounds__(WARPS * 32)
interp_rowlane_kernel(int64_t M, GridGeom g, RowLaneGeom rl, const int* __restrict__ sub_total,
                      const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                      const int4* __restrict__ start, const F* __restrict__ wrec,
                      const Cplx<F>* __restrict__ fw, Cplx<F>* __restrict__ c,
                      const __grid_constant__ CUtensorMap tmap, int use_tma) {
  using C = Cplx<F>;
  extern __shared__ __align__(128) unsigned char tile_raw[];
  C* tile_rl = reinterpret_cast<C*>(tile_raw);
  const int s = blockIdx.x;
  // the subproblem count and this CTA's descriptor are independent loads (the descriptor
  // array has an entry for every launched CTA): one global round trip instead of two
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const int b = sd.x, p0 = sd.y, np = sd.z;
  const int TX = rl.TX, TY = rl.TY, TZ = RANK > 2 ? rl.TZ : 1;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - rl.hx, oy = by * g.bin[1] - rl.hy, oz = RANK > 2 ? bz * g.bin[2] - rl.hz : 0;
  const int plane = TX * TY;
  const int ncell = plane * TZ;
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(tile_rl) + ((static_cast<size_t>(ncell) * sizeof(C) + 127) & ~static_cast<size_t>(127)));
  const C* fwt = fw + static_cast<int64_t>(t) * g.nftot;
  C* ct = c + static_cast<int64_t>(t) * M;

  // 2D interior tiles: one TMA box. 3D (and tiles that straddle the periodic boundary): per-cell
  // asynchronous copies with index wrap (the odd tile pitch rules out 128-byte aligned plane boxes).
  const bool interior = RANK == 2 && use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1];
  if (interior) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, static_cast<uint32_t>(ncell * sizeof(C)));
      tma_load_3d(tile_rl, &tmap, bar, 2 * ox, oy, t);
    }
    mbar_wait(bar, 0);
  } else {
    for (int i = tid; i < ncell; i += WARPS * 32) {
      const int ix = i % TX;
      const int iy = (i / TX) % TY;
      const int iz = i / plane;
      const int gx = mod_idx(ox + ix, g.nf[0]);
      const int gy = mod_idx(oy + iy, g.nf[1]);
      const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
      __pipeline_memcpy_async(&tile_rl[i], fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx, sizeof(C));
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
  }

  constexpr int PW = 32 / LP;          // points per warp step
  const int pt = lane / LP;
  const int row = lane % LP;
  const int ngrp = (np + PW - 1) / PW;
  const int nz = RANK > 2 ? rl.ns : 1;
  for (int grp = warp; grp < ngrp; grp += WARPS) {
    const int p = grp * PW + pt;
    const bool valid = p < np;
    F re = F(0), im = F(0);
    int id = 0;
    if (valid) {
      const int64_t j = static_cast<int64_t>(p0) + p;
      const int4 st = start[j];
      id = idx[j];
      const int rx = st.x - ox, ry = st.y - oy, rz = RANK > 2 ? st.z - oz : 0;
      // Memory safety for coordinates outside the declared points_range (see interp.cuh).
      const bool fits = rx >= 0 && rx + PXT <= TX && ry >= 0 && ry + LP <= TY && (RANK < 3 || (rz >= 0 && rz + nz <= TZ));
      if (fits) {
        const F* wx = wrec + j * rl.R;
        const F wy = row < rl.PY ? wx[rl.PX + row] : F(0);
        const F* wz = wx + rl.PX + rl.PY;
        const C* ptr = tile_rl + (rz * TY + ry + row) * TX + rx;
        C w2[PXT / 2];
#pragma unroll
        for (int k = 0; k < PXT; k += 2) w2[k / 2] = *reinterpret_cast<const C*>(wx + k);   // two consecutive weights
        for (int dz = 0; dz < nz; ++dz) {
          F pr = F(0), pi = F(0);
#pragma unroll
          for (int k = 0; k < PXT; k += 2) {
            const C v0 = ptr[k], v1 = ptr[k + 1];
            pr += v0.x * w2[k / 2].x + v1.x * w2[k / 2].y;
            pi += v0.y * w2[k / 2].x + v1.y * w2[k / 2].y;
          }
          const F wzd = RANK > 2 ? wz[dz] : F(1);
          re += wzd * pr;
          im += wzd * pi;
          ptr += plane;
        }
        re *= wy;
        im *= wy;
      }
    }
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (valid && row == 0) ct[id] = make_cplx<F>(re, im);
  }
}

// ------------------------------------------------------------------------------------------------
// type 1: one-warp CTAs, private tile, one point per step. lane = half * 16 + row (LP <= 16):
// a quarter warp = 8 consecutive rows of the same half-row, conflict-free with the odd pitch.
// ------------------------------------------------------------------------------------------------
// 3D: WPT warps share the tile with exclusive ownership of its z-planes (plane z belongs to warp
// z % WPT): every warp walks all points of the subproblem and updates only the stencil planes it
// owns, so there are still no atomics and WPT warps hide each other's latency (the tile of a wide
// kernel fills most of the SM's shared memory: one CTA per SM).
template <typename F, int PXT, int LP, int RANK, int WPT>
__global__ void __launch_bounds__(32 * WPT)
spread_rowlane_kernel(int64_t M, GridGeom g, RowLaneGeom rl, const int* __restrict__ sub_total,
                      const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                      const int4* __restrict__ start, const F* __restrict__ wrec,
                      const Cplx<F>* __restrict__ c, Cplx<F>* __restrict__ fw) {
  using C = Cplx<F>;
  extern __shared__ __align__(128) unsigned char tile_raw[];
  C* tile_rl = reinterpret_cast<C*>(tile_raw);
  const int s = blockIdx.x;
  // the subproblem count and this CTA's descriptor are independent loads (the descriptor
  // array has an entry for every launched CTA): one global round trip instead of two
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  static_assert(WPT == 1 || RANK == 3, "plane ownership is 3D only");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const int b = sd.x, p0 = sd.y, np = sd.z;
  const int TX = rl.TX, TY = rl.TY, TZ = RANK > 2 ? rl.TZ : 1;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - rl.hx, oy = by * g.bin[1] - rl.hy, oz = RANK > 2 ? bz * g.bin[2] - rl.hz : 0;
  const int plane = TX * TY;
  const int ncell = plane * TZ;
  const C* ct = c + static_cast<int64_t>(t) * M;
  C* fwt = fw + static_cast<int64_t>(t) * g.nftot;

  for (int i = tid; i < ncell; i += 32 * WPT) tile_rl[i] = make_cplx<F>(F(0), F(0));
  if (WPT > 1) __syncthreads(); else __syncwarp();

  constexpr int HW = PXT / 2;          // cells per half row
  const int half = lane >> 4;
  const int row = lane & 15;
  const bool row_ok = row < LP;
  const int nz = RANK > 2 ? rl.ns : 1;

  // one-point software pipeline: the record of point p + 1 is fetched while point p is applied
  F wxh[HW];
  F wy = F(0);
  int4 st = make_int4(0, 0, 0, 0);
  C cj = make_cplx<F>(F(0), F(0));
  const F* wz_n = wrec;
  auto fetch = [&](int p) {
    if (p < np) {
      const int64_t j = static_cast<int64_t>(p0) + p;
      const F* wx = wrec + j * rl.R;
#pragma unroll
      for (int k = 0; k < HW; k += 2) {
        const C w2 = *reinterpret_cast<const C*>(wx + half * HW + k);   // two consecutive weights
        wxh[k] = w2.x;
        wxh[k + 1] = w2.y;
      }
      wy = (row_ok && row < rl.PY) ? wx[rl.PX + row] : F(0);
      wz_n = wx + rl.PX + rl.PY;
      st = start[j];
      cj = ct[idx[j]];
    }
  };
  fetch(0);
  for (int p = 0; p < np; ++p) {
    F w[HW];
#pragma unroll
    for (int k = 0; k < HW; ++k) w[k] = wxh[k];
    const F cr = cj.x * wy, ci = cj.y * wy;
    const int rx = st.x - ox, ry = st.y - oy, rz = RANK > 2 ? st.z - oz : 0;
    const F* wz = wz_n;
    fetch(p + 1);
    // Memory safety for coordinates outside the declared points_range: such a window does not lie
    // in this bin's tile and the point is dropped (the reference's behaviour is undefined there).
    const bool fits = rx >= 0 && rx + PXT <= TX && ry >= 0 && ry + LP <= TY && (RANK < 3 || (rz >= 0 && rz + nz <= TZ));
    if (fits && row_ok) {
      // first stencil plane owned by this warp: (rz + dz) % WPT == warp
      const int dz0 = WPT > 1 ? (((warp - rz) % WPT) + WPT) % WPT : 0;
      C* ptr = tile_rl + ((rz + dz0) * TY + ry + row) * TX + rx + half * HW;
      for (int dz = dz0; dz < nz; dz += WPT) {
        const F wzd = RANK > 2 ? wz[dz] : F(1);
        const F czr = cr * wzd, czi = ci * wzd;
        C v[HW];
#pragma unroll
        for (int k = 0; k < HW; ++k) v[k] = ptr[k];
#pragma unroll
        for (int k = 0; k < HW; ++k) {
          v[k].x += czr * w[k];
          v[k].y += czi * w[k];
          ptr[k] = v[k];
        }
        ptr += WPT * plane;
      }
    }
    __syncwarp();
  }
  if (WPT > 1) __syncthreads();

  // flush: native global reductions (REDG.F64 / REDG.F32x2), periodic wrap, untouched cells skipped
  for (int i = tid; i < ncell; i += 32 * WPT) {
    const C v = tile_rl[i];
    if (v.x == F(0) && v.y == F(0)) continue;
    const int ix = i % TX;
    const int iy = (i / TX) % TY;
    const int iz = i / plane;
    const int gx = mod_idx(ox + ix, g.nf[0]);
    const int gy = mod_idx(oy + iy, g.nf[1]);
    const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
    red_add(fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx, v);
  }
}

}  // namespace b200
