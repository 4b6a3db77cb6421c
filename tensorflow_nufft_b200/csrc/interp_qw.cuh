// interp_qw.cuh -- type-2 interpolator, quarter-warp-per-point gather from a shared-memory tile.
//
// Same sums as interp.cuh (reference: InterpSubproblem{2,3}DKernel nufft_plan.cu.cc:1041-1110,
// 1608-1706). interp_tile_f32_kernel lays the 32 lanes of a warp over ONE point's stencil and pays
// a 5-level butterfly per point; here a warp works on FOUR points at once: the 8 lanes of a
// quarter warp take the 8 rows of one point's window and each lane walks its row (8 cells = four
// 128-bit shared loads per z-plane). Per point this costs
//   * the same tile traffic (the stencil's bytes, inherent),
//   * a quarter of the instructions (every warp instruction serves 4 points),
//   * 1.5 shuffles instead of 4.5 (3-level reduction inside the quarter, re and im),
//   * no shared-memory staging of the records at all: a lane reads its point's weights straight
//     from the sorted record array (8 lanes share each 16-byte load), prefetched one group ahead.
// The tile pitch is bin_x + 10 cells = 2 * odd mod 16, which makes the 8 row-lanes of a quarter
// warp hit 8 distinct 16-byte bank groups: every 128-bit load is conflict-free.
#pragma once
#include <cuda.h>
#include <cuda_pipeline.h>

#include "dev_common.cuh"
#include "interp.cuh"
#include "spread.cuh"

namespace b200 {

constexpr int kQwHaloX = 10;   // tile x extent = bin_x + 10 (stencil reach 8 + 2 pad cells for the pitch)

template <int RANK>
inline size_t interp_qw_smem_bytes(const int* bin, int coils = 1) {
  const size_t ncell = static_cast<size_t>(bin[0] + kQwHaloX) * (bin[1] + 8) * (RANK > 2 ? bin[2] + 8 : 1);
  return ((coils * ncell * sizeof(float2) + 127) & ~static_cast<size_t>(127)) + 16;
}

// Transposing reduction over the 8 lanes of a quarter warp: N values per lane in, ONE value per
// lane out (N + ... shuffles instead of 3 N): at each xor step a lane keeps one half of its values
// and hands the other half to its partner. Afterwards lane row r holds the sum of value index
// r / (8 / N) (rows that share an index hold copies).
template <int N>
__device__ __forceinline__ float qw_reduce(float (&v)[N], int lane) {
  static_assert(N == 2 || N == 4 || N == 8, "2, 4 or 8 values");
  int n = N;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    if (n > 1) {
      const bool up = (lane & o) != 0;
      const int h = n / 2;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < h) {
          const float keep = up ? v[i + h] : v[i];
          const float send = up ? v[i] : v[i + h];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    }
  }
  return v[0];
}

// One stencil row of 8 complex cells (four float4 = cells 0..7 as (re, im) pairs) times the 8
// x-weights: (re, im) += w * (cell.re, cell.im) as ONE packed FFMA2 per cell (fma.rn.f32x2, the
// weight rides in the scalar-broadcast operand) instead of two FFMA. Same products, same order.
__device__ __forceinline__ void qw_fma2(float& re, float& im, float w, float cr, float ci) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %2};\n\t"
      "mov.b64 rb, {%3, %4};\n\t"
      "mov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(re), "+f"(im) : "f"(w), "f"(cr), "f"(ci));
}
__device__ __forceinline__ void qw_row_dot(const float4& v0, const float4& v1, const float4& v2, const float4& v3,
                                           const float4& xa, const float4& xb, float& pr, float& pi) {
  pr = v0.x * xa.x;
  pi = v0.y * xa.x;
  qw_fma2(pr, pi, xa.y, v0.z, v0.w);
  qw_fma2(pr, pi, xa.z, v1.x, v1.y);
  qw_fma2(pr, pi, xa.w, v1.z, v1.w);
  qw_fma2(pr, pi, xb.x, v2.x, v2.y);
  qw_fma2(pr, pi, xb.y, v2.z, v2.w);
  qw_fma2(pr, pi, xb.z, v3.x, v3.y);
  qw_fma2(pr, pi, xb.w, v3.z, v3.w);
}

// Gathers the points [0, np) of one subproblem from NC coil tiles (`tile4`, tile k at offset
// k * ncell / 2 float4; origin ox, oy, oz; pitch TX cells). Groups of 4 points are dealt
// round-robin to the `nwarps` warps. Records are read from global memory (sorted order, base
// index p0) ONCE for the NC coils. Output: ct[k * M + id].
template <int NS, int RANK, int NC, int PF, typename WaitTile>
__device__ __forceinline__ void qw_gather(WaitTile&& wait_tile, const float4* __restrict__ tile4, int TX, int TY, int TZ, int ox, int oy, int oz,
                                          int p0, int np, int warp, int nwarps, int lane,
                                          const int* __restrict__ idx, const int4* __restrict__ start,
                                          const float4* __restrict__ wrec4, float2* __restrict__ ct, int64_t M) {
  constexpr int C4 = 2 * RANK;
  constexpr int NV = 2 * NC;            // values per lane: (re, im) per coil
  constexpr int NR = NV > 8 ? 8 : NV;   // values per reduction (NC = 8: two reductions of 8)
  constexpr int DUP = 8 / NR;           // rows of a quarter that end up with the same value
  const int pt = lane >> 3;
  const int row = lane & 7;
  const int zstride4 = TY * TX / 2;
  const int tile_f4 = zstride4 * TZ;
  const int ngrp = (np + 3) >> 2;

  // Per-lane record of one point, prefetched PF groups ahead. PF = 2 on sparse 3D point sets (the
  // records stream from DRAM and one group of gathers is shorter than a DRAM round trip: cfg4
  // 1.97 -> 1.83 ms); PF = 1 elsewhere (the extra 24 registers cost a resident CTA: cfg3 as type 2
  // 1.20 -> 1.27 ms with PF = 2).
  struct Rec {
    float4 wxa, wxb, wza, wzb;
    float wy;
    int4 st;
    int id;
  };
  auto fetch = [&](int grp) {
    Rec r;
    r.wxa = r.wxb = r.wza = r.wzb = make_float4(0.f, 0.f, 0.f, 0.f);
    r.wy = 0.f;
    r.st = make_int4(0, 0, 0, 0);
    r.id = 0;
    const int p = 4 * grp + pt;
    if (grp < ngrp && p < np) {
      const int64_t j = static_cast<int64_t>(p0) + p;
      r.wxa = wrec4[j * C4];
      r.wxb = wrec4[j * C4 + 1];
      r.wy = reinterpret_cast<const float*>(wrec4 + j * C4 + 2)[row];
      if (RANK > 2) {
        r.wza = wrec4[j * C4 + 4];
        r.wzb = wrec4[j * C4 + 5];
      }
      r.st = start[j];
      r.id = idx[j];
    }
    return r;
  };
  Rec r1 = fetch(warp);
  Rec r2 = r1;
  if (PF == 2) r2 = fetch(warp + nwarps);
  wait_tile();   // the first records are in flight while the tile lands
  for (int grp = warp; grp < ngrp; grp += nwarps) {
    const float4 xa = r1.wxa, xb = r1.wxb, za = r1.wza, zb = r1.wzb;
    const float wy_c = r1.wy;
    const int4 st_c = r1.st;
    const int id_c = r1.id;
    if (PF == 2) {
      r1 = r2;
      r2 = fetch(grp + 2 * nwarps);
    } else {
      r1 = fetch(grp + nwarps);
    }

    const bool valid = 4 * grp + pt < np;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
    const int rx = st_c.x - ox, ry = st_c.y - oy, rz = RANK > 2 ? st_c.z - oz : 0;
    // Memory safety for coordinates outside the declared points_range (see interp.cuh).
    const bool fits = rx >= 0 && rx + 8 <= TX && ry >= 0 && ry + NS <= TY && (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
    if (valid && fits && row < NS) {
      const float4* ptr = tile4 + (((rz * TY + ry + row) * TX + rx) >> 1);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const float4* pk = ptr + k * tile_f4;
        float re = 0.f, im = 0.f;
        if (RANK == 2) {
          const float4 v0 = pk[0], v1 = pk[1], v2 = pk[2], v3 = pk[3];
          qw_row_dot(v0, v1, v2, v3, xa, xb, re, im);
        } else {
          const float wz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) {
            const float4* pz = pk + dz * zstride4;
            const float4 v0 = pz[0], v1 = pz[1], v2 = pz[2], v3 = pz[3];
            float pr, pi;
            qw_row_dot(v0, v1, v2, v3, xa, xb, pr, pi);
            qw_fma2(re, im, wz[dz], pr, pi);
          }
        }
        v[2 * k] = re * wy_c;
        v[2 * k + 1] = im * wy_c;
      }
    }
#pragma unroll
    for (int h = 0; h < NV / NR; ++h) {
      float part[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) part[i] = v[h * NR + i];
      const float sum = qw_reduce<NR>(part, lane);
      // lane row r holds value index h * NR + r / DUP = 2 * coil + (0: re, 1: im)
      const int vi = h * NR + row / DUP;
      if (valid && (row % DUP) == 0)
        reinterpret_cast<float*>(ct + static_cast<int64_t>(vi >> 1) * M + id_c)[vi & 1] = sum;
    }
  }
}

// One CTA (WARPS warps) per (subproblem, group of NC transforms); the NC coil tiles are staged by ONE TMA
// box copy (interior; the box spans NC transforms) or
// wrapped 16-byte cp.async copies (tiles that straddle the periodic boundary); the other resident
// CTAs of the SM hide the tile latency. (A persistent grid-stride variant and a two-stage tile ring
// were measured slower on every BASELINE config: static striding loses the hardware scheduler's
// load balancing between heavy and light subproblems.)
template <int NS, int RANK, int NC, int PF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
interp_qw_f32_kernel(int64_t M, GridGeom g, const int* __restrict__ sub_total,
                     const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                     const int4* __restrict__ start, const float4* __restrict__ wrec4,
                     const float2* __restrict__ fw, float2* __restrict__ c,
                     const __grid_constant__ CUtensorMap tmap, int use_tma, int zrange) {
  extern __shared__ __align__(128) float4 smem4[];
  const int s = blockIdx.x;
  // the subproblem count and this CTA's descriptor are independent loads (the descriptor
  // array has an entry for every launched CTA): one global round trip instead of two
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y * NC;   // first transform of this CTA's group
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + kQwHaloX, TY = g.bin[1] + 8;
  const int TZ = RANK > 2 ? g.bin[2] + 8 : 1;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = RANK > 2 ? bz * g.bin[2] - 4 : 0;
  const int ncell = TX * TY * TZ;
  const int TXH = TX / 2;
  float4* tile4 = smem4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem4) + ((static_cast<size_t>(NC) * ncell * sizeof(float2) + 127) & ~static_cast<size_t>(127)));

  const float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;
  const bool interior = use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
                        (RANK < 3 || (oz >= 0 && oz + TZ <= g.nf[2]));
  // 3D: only the z-planes the subproblem's stencils reach are loaded (sub_desc.w, see
  // subproblem_zrange_kernel); the tensor map's box is ONE plane and the planes complete on the
  // same mbarrier. 2D: one box for the NC coil tiles.
  int tz_lo = 0, tz_hi = TZ;
  if (RANK == 3 && zrange) {
    const int lo = (sd.w & 0xffff) - 32768 - oz, hi = ((sd.w >> 16) & 0xffff) - 32768 - oz + NS;
    if (lo >= 0 && hi <= TZ && lo < hi) { tz_lo = lo; tz_hi = hi; }
  }
  const int plane_cells = TX * TY;
  if (interior) {
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      if (RANK == 2) {
        mbar_expect_tx(bar, static_cast<uint32_t>(NC * ncell * sizeof(float2)));
        tma_load_3d(tile4, &tmap, bar, 2 * ox, oy, t);
      } else {
        mbar_expect_tx(bar, static_cast<uint32_t>((tz_hi - tz_lo) * plane_cells * sizeof(float2)));
        for (int z = tz_lo; z < tz_hi; ++z) tma_load_4d(tile4 + z * (plane_cells / 2), &tmap, bar, 2 * ox, oy, oz + z, t);
      }
    }
  } else {
    for (int i = tid + tz_lo * (plane_cells / 2); i < tz_hi * (plane_cells / 2); i += WARPS * 32) {
      const int ix = i % TXH;
      const int iy = (i / TXH) % TY;
      const int iz = i / (TXH * TY);
      const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
      const int gy = mod_idx(oy + iy, g.nf[1]);
      const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
      const float2* src = fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx;
#pragma unroll
      for (int k = 0; k < NC; ++k) __pipeline_memcpy_async(&tile4[k * (ncell / 2) + i], src + k * g.nftot, 16);
    }
    __pipeline_commit();
  }
  auto wait_tile = [&]() {
    if (interior) {
      mbar_wait(bar, 0);
    } else {
      __pipeline_wait_prior(0);
      __syncthreads();
    }
  };
  qw_gather<NS, RANK, NC, PF>(wait_tile, tile4, TX, TY, TZ, ox, oy, oz, p0, np, warp, WARPS, lane, idx, start, wrec4,
                          c + static_cast<int64_t>(t) * M, M);
}

}  // namespace b200
