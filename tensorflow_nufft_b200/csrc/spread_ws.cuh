// spread_ws.cuh -- window-sorted type-1 spreader (complex64, ns <= 7, rank 2 or 3).
//
// Same sums as spread.cuh (reference: SpreadSubproblem{2,3}DKernel nufft_plan.cu.cc:790-878,
// 1404-1510), different schedule. The bin-sort key is refined to (bin, stencil window) so that the
// points of a subproblem arrive grouped by the position of their stencil inside the bin's tile.
// One warp owns one tile. Lanes are laid over the stencil (row r = lane / QX, cell pair
// q = lane % QX); for a RUN of points with the same window every lane's target cells are the same,
// so the run is accumulated in REGISTERS (4 FFMA per point per lane in 2D, 28 in 3D) and the
// shared-memory tile is touched once per run (LDS.128 / FADD / STS.128) instead of once per point.
// In 3D a lane keeps the whole z-column of the tile under its (x pair, y row) in registers, the z
// start of each point selects one of TZ-NS+1 unrolled update variants (warp-uniform switch).
// Shared-memory traffic per point drops from 8 (2D) / 56 (3D) read-modify-write wavefronts to
// 8 / run-length (2D) and 8*TZ / run-length (3D); no atomics in shared memory; tile flushed to the
// fine grid with REDG.E.ADD.F32x4.
#pragma once
#include "dev_common.cuh"
#include "spread.cuh"

namespace b200 {

template <int RANK, int NC> struct WsRec {
  // words: [0..7]   wx[8]
  //        [8]      2 * off2d + flag; off2d = window offset in cells (-1: dropped point),
  //                 flag = 1 if this point opens a new run (window differs from the previous point)
  //        [9]      tz (3D: tile z of the stencil start)        [10..11] pad
  //        [12..19] wy[8]
  //        [20 + 2 k]  c_k = strength of this point in coil k (re, im), k < NC
  //        [kWz ...]   wz[8]                                       (3D)
  // (c_k * wy[r] is formed on the fly: 2 FMUL per coil per point, in exchange for a stage record of
  //  ~30 words instead of 12 + 16 NC, i.e. more resident CTAs per SM.)
  // 128-bit shared loads always cost 4 wavefronts (one per quarter warp) even when lanes share
  // addresses, so the per-point broadcast data is read with 64/32-bit loads.
  static constexpr int kWy = 12;
  static constexpr int kC = 20;
  static constexpr int kWz = 20 + ((2 * NC + 3) / 4) * 4;
  static constexpr int kRaw = kWz + (RANK == 3 ? 8 : 0);
  static constexpr int kStride = ((kRaw / 4) % 2 == 1) ? kRaw : kRaw + 4;   // 4 * odd words: conflict-free staging
  static_assert((kStride / 4) % 2 == 1, "stage stride must be an odd multiple of 4 words");
};

template <int NS, int TZ, int TZ0>
struct ColumnUpdate {
  // v[TZ0 + dz] += wz[dz] * cx  for dz < NS, with compile-time register indices
  __device__ static __forceinline__ void run(float4 (&v)[TZ], const float (&wz)[8], const float4& cx) {
#pragma unroll
    for (int dz = 0; dz < NS; ++dz) {
      if (TZ0 + dz < TZ) {
        v[TZ0 + dz].x += wz[dz] * cx.x;
        v[TZ0 + dz].y += wz[dz] * cx.y;
        v[TZ0 + dz].z += wz[dz] * cx.z;
        v[TZ0 + dz].w += wz[dz] * cx.w;
      }
    }
  }
};

template <int NS, int TZ, int K>
struct ColumnDispatch {
  __device__ static __forceinline__ void run(int tz, float4 (&v)[TZ], const float (&wz)[8], const float4& cx) {
    if (tz == K) ColumnUpdate<NS, TZ, K>::run(v, wz, cx);
    else ColumnDispatch<NS, TZ, K - 1>::run(tz, v, wz, cx);
  }
};
template <int NS, int TZ>
struct ColumnDispatch<NS, TZ, -1> {
  __device__ static __forceinline__ void run(int, float4 (&)[TZ], const float (&)[8], const float4&) {}
};

// TZ = bin_z + 8 for RANK 3 (compile time: the z-column lives in registers), 1 for RANK 2.
// NC = coils (transforms) handled by one CTA (2D only): the per-point work that does not depend on
// the coil -- weight loads, run detection, loop control -- is shared by NC accumulator sets and NC
// tiles, which cuts instructions and shared-memory traffic per point-transform by ~NC/2.
template <int NS, int RANK, int TZ, int NC>
__global__ void __launch_bounds__(32)
spread_ws_f32_kernel(int64_t M, GridGeom g, const int* __restrict__ sub_total,
                     const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                     const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][2*RANK]*/,
                     const float2* __restrict__ c, float2* __restrict__ fw,
                     const __grid_constant__ CUtensorMap tmap_out, int use_tma) {
  constexpr int QX = (NS + 2) / 2;
  static_assert(QX * NS <= 32, "stencil slab must fit one warp");
  constexpr int C4 = 2 * RANK;
  static_assert(RANK == 2 || NC == 1, "multi-coil CTAs are 2D only");
  using Rec = WsRec<RANK, NC>;
  constexpr int SW = Rec::kStride;
  constexpr int BS = 32;   // points per staged batch (16 was measured: no gain)
  extern __shared__ __align__(128) float4 smem4[];

  const int s = blockIdx.x;
  // the subproblem count and this CTA's descriptor are independent loads (the descriptor
  // array has an entry for every launched CTA): one global round trip instead of two
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int lane = threadIdx.x;
  const int t = blockIdx.y;
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + 8, TY = g.bin[1] + 8;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = RANK > 2 ? bz * g.bin[2] - 4 : 0;
  const int plane = TX * TY;
  const int ncell = plane * TZ;
  float4* tile4 = smem4;                                        // [NC][ncell / 2]
  float2* tile = reinterpret_cast<float2*>(tile4);
  float* stage = reinterpret_cast<float*>(smem4 + NC * (ncell / 2));   // [BS + 1][SW]

  for (int i = lane; i < NC * (ncell / 2); i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < NS;
  const int rr = row_ok ? r : 0;
  const int lane_off = rr * TX + 2 * q;
  const int zstride4 = plane / 2;

  const float2* ct = c + static_cast<int64_t>(t) * NC * M;      // coil k: ct + k * M
  float2* fwt = fw + static_cast<int64_t>(t) * NC * g.nftot;

  // ---- register prefetch of this lane's point of the next batch ----
  float4 w4[C4];
  int4 st_n = make_int4(0, 0, 0, 0);
  float2 c_n[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) c_n[k] = make_float2(0.f, 0.f);
  int id_n2 = 0;
  auto fetch = [&](int bb) {
    const int pl = bb * BS + lane;
    if (lane < BS && pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < C4; ++k) w4[k] = wrec4[j * C4 + k];
      st_n = start[j];
#pragma unroll
      for (int k = 0; k < NC; ++k) c_n[k] = ct[static_cast<int64_t>(k) * M + id_n2];
    }
    const int pl2 = (bb + 1) * BS + lane;
    if (lane < BS && pl2 < np) id_n2 = idx[p0 + pl2];
  };
  int last_off = -2;   // window of the last point of the previous batch (forces a flag at the start)
  auto stage_write = [&](int bb) {
    const int pl = bb * BS + lane;
    float4* rec4 = reinterpret_cast<float4*>(stage + lane * SW);
    int off = -1, tz = 0;
    if (lane < BS && pl < np) {
      const int rx = st_n.x - ox, ry = st_n.y - oy, rz = RANK > 2 ? st_n.z - oz : 0;
      // Memory safety for coordinates outside the declared points_range: the stencil does not lie
      // in this bin's tile and the point is dropped (the reference's behaviour is undefined there).
      const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                        (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
      if (fits) { off = ry * TX + rx; tz = rz; }
    }
    // run flag: compare with the previous point's window (previous lane, or the last point of the
    // previous batch for lane 0)
    const int prev = __shfl_up_sync(0xffffffffu, off, 1);
    const int flag = (lane == 0 ? (off != last_off) : (off != prev)) ? 1 : 0;
    last_off = __shfl_sync(0xffffffffu, off, BS - 1);
    if (lane >= BS) return;
    rec4[0] = w4[0];
    rec4[1] = w4[1];
    rec4[2] = make_float4(__int_as_float(off * 2 + flag), __int_as_float(tz), 0.f, 0.f);
    rec4[3] = w4[2];
    rec4[4] = w4[3];
    float* recf = reinterpret_cast<float*>(rec4);
#pragma unroll
    for (int k = 0; k < NC; ++k) *reinterpret_cast<float2*>(recf + Rec::kC + 2 * k) = c_n[k];
    if (RANK > 2) {
      *reinterpret_cast<float4*>(recf + Rec::kWz) = w4[C4 - 2];
      *reinterpret_cast<float4*>(recf + Rec::kWz + 4) = w4[C4 - 1];
    }
  };
  if (lane < BS && lane < np) id_n2 = idx[p0 + lane];
  fetch(0);

  // ---- run accumulators: v[coil][plane] ----
  constexpr int NV = NC * TZ;
  float4 v[NV];
#pragma unroll
  for (int z = 0; z < NV; ++z) v[z] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1;   // 2D window offset (cells) of the open run, -1 = none
  int zlo = TZ, zhi = 0;   // planes touched by the open run (warp-uniform)
  auto flush_run = [&]() {
    if (cur >= 0 && row_ok) {
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        float4* ptr = reinterpret_cast<float4*>(tile + static_cast<size_t>(k) * ncell + cur + lane_off);
#pragma unroll
        for (int z = 0; z < TZ; ++z) {
          if (RANK == 2 || (z >= zlo && z < zhi)) {
            float4 tv = ptr[z * zstride4];
            tv.x += v[k * TZ + z].x; tv.y += v[k * TZ + z].y; tv.z += v[k * TZ + z].z; tv.w += v[k * TZ + z].w;
            ptr[z * zstride4] = tv;
          }
        }
      }
    }
#pragma unroll
    for (int z = 0; z < NV; ++z) v[z] = make_float4(0.f, 0.f, 0.f, 0.f);
    zlo = TZ;
    zhi = 0;
    __syncwarp();
  };

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    stage_write(bb);
    __syncwarp();
    if (bb + 1 < nbatch) fetch(bb + 1);

    const int cnt = min(BS, np - bb * BS);
    // The stage holds BS + 1 records so that the prefetch of point p + 1 never needs a guard.
    float2 wx = *reinterpret_cast<const float2*>(stage + 2 * q);
    int of = __float_as_int(stage[8]);
    float wy = stage[Rec::kWy + rr];
    float2 cc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) cc[k] = *reinterpret_cast<const float2*>(stage + Rec::kC + 2 * k);
#pragma unroll 2
    for (int p = 0; p < cnt; ++p) {
      const float* rec = stage + p * SW;
      const float2 wx_c = wx;
      const int of_c = of;
      float2 cw_c[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) cw_c[k] = make_float2(cc[k].x * wy, cc[k].y * wy);
      wx = *reinterpret_cast<const float2*>(rec + SW + 2 * q);
      of = __float_as_int(rec[SW + 8]);
      wy = rec[SW + Rec::kWy + rr];
#pragma unroll
      for (int k = 0; k < NC; ++k) cc[k] = *reinterpret_cast<const float2*>(rec + SW + Rec::kC + 2 * k);
      if (of_c & 1) {          // warp-uniform: this point opens a new run
        flush_run();
        cur = of_c >> 1;
      }
      // dropped points carry cur = -1 and are accumulated into registers that are never stored
      if constexpr (RANK == 2) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          v[k].x += cw_c[k].x * wx_c.x; v[k].y += cw_c[k].y * wx_c.x;
          v[k].z += cw_c[k].x * wx_c.y; v[k].w += cw_c[k].y * wx_c.y;
        }
      } else {
        const float4 cx = make_float4(cw_c[0].x * wx_c.x, cw_c[0].y * wx_c.x, cw_c[0].x * wx_c.y, cw_c[0].y * wx_c.y);
        float wz[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t2 = *reinterpret_cast<const float2*>(rec + Rec::kWz + 2 * k);
          wz[2 * k] = t2.x;
          wz[2 * k + 1] = t2.y;
        }
        const int tz = __float_as_int(rec[9]);
        zlo = min(zlo, tz);
        zhi = max(zhi, tz + NS);
        ColumnDispatch<NS, TZ, TZ - NS>::run(tz, v, wz, cx);
      }
    }
    __syncwarp();
  }
  flush_run();

  // Flush the tiles: interior tiles by TMA reduce-add (one instruction per coil), tiles that
  // straddle the periodic boundary with two complex cells per REDG.ADD.F32x4 (index wrap, zero
  // pairs skipped).
  if (use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
      (RANK < 3 || (oz >= 0 && oz + TZ <= g.nf[2]))) {
    fence_proxy_async_smem();   // every lane: its generic-proxy tile writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (RANK == 2) tma_reduce_add_3d(&tmap_out, tile4 + k * (ncell / 2), 2 * ox, oy, t * NC + k);
        else tma_reduce_add_4d(&tmap_out, tile4 + k * (ncell / 2), 2 * ox, oy, oz, t * NC + k);
      }
      tma_store_commit_and_wait_read();
    }
    return;
  }
  const int TXH = TX / 2;
  for (int i = lane; i < ncell / 2; i += 32) {
    const int ix = i % TXH;
    const int iy = (i / TXH) % TY;
    const int iz = i / (TXH * TY);
    const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
    const int gy = mod_idx(oy + iy, g.nf[1]);
    const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
    const int64_t cell = (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const float4 tv = tile4[k * (ncell / 2) + i];
      if (tv.x == 0.f && tv.y == 0.f && tv.z == 0.f && tv.w == 0.f) continue;
      red_add(reinterpret_cast<float4*>(fwt + static_cast<int64_t>(k) * g.nftot + cell), tv);
    }
  }
}

template <int RANK, int NC>
inline size_t spread_ws_smem_bytes(const int* bin) {
  const size_t ncell = static_cast<size_t>(bin[0] + 8) * (bin[1] + 8) * (RANK > 2 ? bin[2] + 8 : 1);
  return NC * ncell * sizeof(float2) + 33 * WsRec<RANK, NC>::kStride * sizeof(float);
}

}  // namespace b200
