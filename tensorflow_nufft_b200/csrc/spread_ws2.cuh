// spread_ws2.cuh -- window-sorted 2D type-1 spreader, second generation (complex64, ns <= 7).
//
// Same sums as spread.cuh / spread_ws.cuh (reference: SpreadSubproblem2DKernel
// nufft_plan.cu.cc:790-878), same schedule as spread_ws.cuh (one warp per tile, lanes over the
// stencil window, a RUN of points with the same window accumulated in registers), with three
// changes that cut instructions and shared-memory wavefronts per point-transform:
//   * the window is (NS + 1) rows tall and starts on an EVEN row (the record kernel shifts wy by
//     the parity of the stencil's first row and zero-pads, exactly as it does for x): windows are
//     2 x 2 cells apart instead of 2 x 1, so runs are twice as long and the tile is touched half
//     as often; for NS = 7 all 32 lanes are busy (8 rows x 4 cell pairs);
//   * per point the lane forms w = wx * wy once (2 FMUL) and every coil costs 4 FFMA with its
//     broadcast strength, instead of 2 FMUL + 4 FFMA per coil;
//   * the first point of a run writes the accumulators (FMUL) instead of zeroing + FFMA.
#pragma once
#include "dev_common.cuh"
#include "interp.cuh"
#include "spread.cuh"
#include "spread_ws.cuh"

namespace b200 {

template <int NS, int NC>
__global__ void __launch_bounds__(32)
spread_ws2_f32_kernel(int64_t M, GridGeom g, const int* __restrict__ sub_total,
                      const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                      const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][4]*/,
                      const float2* __restrict__ c, float2* __restrict__ fw,
                      const __grid_constant__ CUtensorMap tmap_out, int use_tma) {
  constexpr int QX = (NS + 2) / 2;
  constexpr int ROWS = NS + 1;
  static_assert(QX * ROWS <= 32, "stencil window must fit one warp");
  using Rec = WsRec<2, NC>;
  constexpr int SW = Rec::kStride;
  constexpr int BS = 32;
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
  const int by = b / g.nbins[0];
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4;
  const int ncell = TX * TY;
  float4* tile4 = smem4;                                               // [NC][ncell / 2]
  float2* tile = reinterpret_cast<float2*>(tile4);
  float* stage = reinterpret_cast<float*>(smem4 + NC * (ncell / 2));   // [BS + 1][SW]

  for (int i = lane; i < NC * (ncell / 2); i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < ROWS;
  const int rr = row_ok ? r : 0;
  const int lane_off = rr * TX + 2 * q;

  const float2* ct = c + static_cast<int64_t>(t) * NC * M;
  float2* fwt = fw + static_cast<int64_t>(t) * NC * g.nftot;

  // ---- register prefetch of this lane's point of the next batch ----
  float4 w4[4];
  int4 st_n = make_int4(0, 0, 0, 0);
  float2 c_n[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) c_n[k] = make_float2(0.f, 0.f);
  int id_n2 = 0;
  auto fetch = [&](int bb) {
    const int pl = bb * BS + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < 4; ++k) w4[k] = wrec4[j * 4 + k];
      st_n = start[j];
#pragma unroll
      for (int k = 0; k < NC; ++k) c_n[k] = ct[static_cast<int64_t>(k) * M + id_n2];
    }
    const int pl2 = (bb + 1) * BS + lane;
    if (pl2 < np) id_n2 = idx[p0 + pl2];
  };
  int last_off = -2;
  auto stage_write = [&](int bb) {
    const int pl = bb * BS + lane;
    float4* rec4 = reinterpret_cast<float4*>(stage + lane * SW);
    int off = -1;
    if (pl < np) {
      const int rx = st_n.x - ox, ry = st_n.y - oy;
      // Memory safety for coordinates outside the declared points_range: the window does not lie
      // in this bin's tile and the point is dropped (the reference's behaviour is undefined there).
      const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + ROWS <= TY;
      if (fits) off = ry * TX + rx;
    }
    const int prev = __shfl_up_sync(0xffffffffu, off, 1);
    const int flag = (lane == 0 ? (off != last_off) : (off != prev)) ? 1 : 0;
    last_off = __shfl_sync(0xffffffffu, off, BS - 1);
    rec4[0] = w4[0];
    rec4[1] = w4[1];
    rec4[2] = make_float4(__int_as_float(off * 2 + flag), 0.f, 0.f, 0.f);
    rec4[3] = w4[2];
    rec4[4] = w4[3];
    float* recf = reinterpret_cast<float*>(rec4);
#pragma unroll
    for (int k = 0; k < NC; ++k) *reinterpret_cast<float2*>(recf + Rec::kC + 2 * k) = c_n[k];
  };
  if (lane < np) id_n2 = idx[p0 + lane];
  fetch(0);

  float4 v[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1;   // window offset (cells) of the open run, -1 = none / dropped points
  auto store_run = [&]() {
    if (cur >= 0 && row_ok) {
      float4 tv[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k)
        tv[k] = *reinterpret_cast<const float4*>(tile + static_cast<size_t>(k) * ncell + cur + lane_off);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        tv[k].x += v[k].x; tv[k].y += v[k].y; tv[k].z += v[k].z; tv[k].w += v[k].w;
        *reinterpret_cast<float4*>(tile + static_cast<size_t>(k) * ncell + cur + lane_off) = tv[k];
      }
    }
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
      const float w0 = wx.x * wy, w1 = wx.y * wy;
      const int of_c = of;
      float2 cc_c[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) cc_c[k] = cc[k];
      wx = *reinterpret_cast<const float2*>(rec + SW + 2 * q);
      of = __float_as_int(rec[SW + 8]);
      wy = rec[SW + Rec::kWy + rr];
#pragma unroll
      for (int k = 0; k < NC; ++k) cc[k] = *reinterpret_cast<const float2*>(rec + SW + Rec::kC + 2 * k);
      if (of_c & 1) {          // warp-uniform: this point opens a new run
        store_run();
        cur = of_c >> 1;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          v[k].x = cc_c[k].x * w0; v[k].y = cc_c[k].y * w0;
          v[k].z = cc_c[k].x * w1; v[k].w = cc_c[k].y * w1;
        }
      } else {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          v[k].x += cc_c[k].x * w0; v[k].y += cc_c[k].y * w0;
          v[k].z += cc_c[k].x * w1; v[k].w += cc_c[k].y * w1;
        }
      }
    }
    __syncwarp();
  }
  store_run();

  // Flush the tiles. Interior tiles: ONE TMA reduce-add per coil (the TMA unit reads the tile and
  // adds it to the fine grid in L2; no LSU work at all). Tiles that straddle the periodic boundary:
  // two complex cells per REDG.ADD.F32x4 with index wrap, zero pairs skipped.
  if (use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1]) {
    fence_proxy_async_smem();   // every lane: its generic-proxy tile writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NC; ++k) tma_reduce_add_3d(&tmap_out, tile4 + k * (ncell / 2), 2 * ox, oy, t * NC + k);
      tma_store_commit_and_wait_read();   // the tile must stay allocated until it has been read
    }
    return;
  }
  const int TXH = TX / 2;
  for (int i = lane; i < ncell / 2; i += 32) {
    const int ix = i % TXH;
    const int iy = i / TXH;
    const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
    const int gy = mod_idx(oy + iy, g.nf[1]);
    const int64_t cell = static_cast<int64_t>(gy) * g.nf[0] + gx;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const float4 tv = tile4[k * (ncell / 2) + i];
      if (tv.x == 0.f && tv.y == 0.f && tv.z == 0.f && tv.w == 0.f) continue;
      red_add(reinterpret_cast<float4*>(fwt + static_cast<int64_t>(k) * g.nftot + cell), tv);
    }
  }
}

}  // namespace b200
