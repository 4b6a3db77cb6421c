// spread_sweep.cuh -- type-1 spreaders that SWEEP a register-resident window along x
// (complex64, ns <= 7). Third generation of the 2D spreader and second of the 3D one.
//
// Same sums as spread.cuh / spread_ws2.cuh (reference: SpreadSubproblem{2,3}DKernel
// nufft_plan.cu.cc:790-878,1404-1510). What changed against spread_ws2.cuh, and why
// (profiles/r02_pipe_probe.jsonl, r01_ncu_ws2_cfg2_summary.txt): that kernel spends its time in the
// load/store pipe -- 8 broadcast strength loads per point for 8 coils and a 128-bit read-modify-write
// of the whole 8 x 8 window, for 8 coils, at the end of every run of ~7 points -- with 2 accumulator
// cells per lane.
//   * A lane now owns a whole ROW of the window: 8 x-cells x Y coils (2D) or x 2 z-planes (3D) = 32
//     accumulator floats. Per point it loads the 8 x-weights (warp-uniform, 2 x LDS.128), one
//     y-weight, Y strengths (one LDS.128 for Y = 2) and issues 32 FFMA (16 packed FFMA2): 5 shared
//     loads per 32 FFMA instead of 11.
//   * Points are sorted by (bin, window row, window column), so consecutive runs of a sweep are
//     windows that move 2 cells to the right. The accumulators ROTATE instead of being flushed: the
//     stage record holds the x-weights already rotated to the register slots (cell pair c lives in
//     slot c & 3), so the inner loop has no phase, and a window step flushes only the ONE outgoing
//     cell pair (a 128-bit read-modify-write per coil) instead of all four.
//   * Interior tiles leave through the TMA unit (cp.reduce.async.bulk.tensor add), as before.
// The tile pitch is bin_x + 10 cells (= 2 mod 16), which makes the 8 row-lanes of a quarter warp
// hit 8 distinct 16-byte bank groups.
#pragma once
#include <type_traits>

#include "dev_common.cuh"
#include "interp.cuh"
#include "spread.cuh"

namespace b200 {

constexpr int kSweepHaloX = 10;   // tile x extent = bin_x + 10

template <int Y> struct SweepRec2 {
  // words: [0..7]   x-weights, rotated: float2 slot s = weights of the tile cell pair with (pair & 3) == s
  //        [8..15]  wy[8] (rows of the even-aligned window)
  //        [16]     header = (window << 1) | new-run flag, window = wy_index * 64 + wx_index, -1 = dropped point
  //        [20 + 2 k] strength of this point in coil k (re, im), k < 4 Y
  static constexpr int kWy = 8, kHdr = 16, kC = 20;
  static constexpr int kRaw = 20 + 8 * Y;
  static constexpr int kStride = ((kRaw / 4) % 2 == 1) ? kRaw : kRaw + 4;   // 4 * odd words: conflict-free staging
  static_assert((kStride / 4) % 2 == 1, "stage stride must be an odd multiple of 4 words");
};

// acc (cells a, b) += s * (wa, wb): one packed FFMA2 (fma.rn.f32x2), or two FFMA.
template <int PACK>
__device__ __forceinline__ void fma_pair(float2& acc, float s, float2 w) {
  if (PACK) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(acc.x), "+f"(acc.y) : "f"(s), "f"(w.x), "f"(w.y));
  } else {
    acc.x = fmaf(s, w.x, acc.x);
    acc.y = fmaf(s, w.y, acc.y);
  }
}

template <int Y>
inline size_t spread_sweep2d_smem_bytes(const int* bin) {
  const size_t ncell = static_cast<size_t>(bin[0] + kSweepHaloX) * (bin[1] + 8);
  return 4 * Y * ncell * sizeof(float2) + 33 * SweepRec2<Y>::kStride * sizeof(float);
}

// One warp per (subproblem, group of NC = 4 Y coils). lane = g * 8 + r: r = row of the window,
// g = coil sub-group (coils g * Y .. g * Y + Y - 1 of the CTA's group).
template <int NS, int Y, int PACK>
__global__ void __launch_bounds__(32)
spread_sweep2d_f32_kernel(int64_t M, GridGeom g, int ngroups, const int* __restrict__ sub_total,
                          const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                          const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][4]*/,
                          const float2* __restrict__ c, float2* __restrict__ fw,
                          const __grid_constant__ CUtensorMap tmap_out, int use_tma) {
  static_assert(NS <= 7, "8-cell windows");
  constexpr int NC = 4 * Y;
  using Rec = SweepRec2<Y>;
  constexpr int SW = Rec::kStride;
  constexpr int BS = 32;
  constexpr int kWinStride = 64;   // window id = wy_index * 64 + wx_index (bin_x <= 96)
  extern __shared__ __align__(128) float4 smem4[];

  // coil groups of one subproblem are adjacent CTAs: they read the same records while those are in L2
  const int s = blockIdx.x / ngroups;
  const int t = blockIdx.x - s * ngroups;
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int lane = threadIdx.x;
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + kSweepHaloX, TY = g.bin[1] + 8;
  const int bx = b % g.nbins[0];
  const int by = b / g.nbins[0];
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4;
  const int ncell = TX * TY;
  float4* tile4 = smem4;                                               // [NC][ncell / 2]
  float* stage = reinterpret_cast<float*>(smem4 + NC * (ncell / 2));   // [BS + 1][SW]

  for (int i = lane; i < NC * (ncell / 2); i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int r = lane & 7;
  const int cg = lane >> 3;
  const bool row_ok = r < NS + 1;   // rows of the even-aligned window that can carry weight
  // this lane's row r of coil (cg * Y + k): float4 index of cell pair 0 of tile row 0
  const int lane_tile4 = (cg * Y) * (ncell / 2) + r * (TX / 2);

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
  int last_win = -2;
  auto stage_write = [&](int bb) {
    const int pl = bb * BS + lane;
    float* rec = stage + lane * SW;
    int win = -1;
    if (pl < np) {
      const int rx = st_n.x - ox, ry = st_n.y - oy;
      // Memory safety for coordinates outside the declared points_range: the window does not lie
      // in this bin's tile and the point is dropped (the reference's behaviour is undefined there).
      // (the zero-padded window may use the two pad columns; rows >= NS + 1 carry no weight and are masked)
      const bool fits = rx >= 0 && rx + 8 <= TX && ry >= 0 && ry + NS + 1 <= TY && ((rx | ry) & 1) == 0;
      if (fits) win = (ry >> 1) * kWinStride + (rx >> 1);
    }
    const int prev = __shfl_up_sync(0xffffffffu, win, 1);
    const int flag = (lane == 0 ? (win != last_win) : (win != prev)) ? 1 : 0;
    last_win = __shfl_sync(0xffffffffu, win, BS - 1);
    // x-weights rotated to their register slots: window cell pair i -> slot (wx_index + i) & 3
    const int rot = win & 3;
    if (win < 0) {   // dropped point: zero weights AND zero strengths (nothing, not even a NaN, reaches the tile)
      w4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      w4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float2* wxs = reinterpret_cast<float2*>(rec);
    wxs[(rot + 0) & 3] = make_float2(w4[0].x, w4[0].y);
    wxs[(rot + 1) & 3] = make_float2(w4[0].z, w4[0].w);
    wxs[(rot + 2) & 3] = make_float2(w4[1].x, w4[1].y);
    wxs[(rot + 3) & 3] = make_float2(w4[1].z, w4[1].w);
    float4* rec4 = reinterpret_cast<float4*>(rec);
    rec4[2] = w4[2];
    rec4[3] = w4[3];
    rec4[4] = make_float4(__int_as_float((win << 1) | flag), 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NC; k += 2)
      rec4[5 + k / 2] = win >= 0 ? make_float4(c_n[k].x, c_n[k].y, c_n[k + 1].x, c_n[k + 1].y)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (lane < np) id_n2 = idx[p0 + lane];
  fetch(0);

  // ---- accumulators: slot s = tile cell pair with (pair & 3) == s; per coil (re_a, re_b), (im_a, im_b) ----
  float2 are[4][Y], aim[4][Y];
#pragma unroll
  for (int sl = 0; sl < 4; ++sl)
#pragma unroll
    for (int k = 0; k < Y; ++k) { are[sl][k] = make_float2(0.f, 0.f); aim[sl][k] = make_float2(0.f, 0.f); }
  int cur_wx = 0, cur_wy = 0;
  bool have = false;

  // adds the accumulators of slot SL to tile cell pair `pair` of this lane's row, and clears them
  auto flush_slot = [&](auto sl_tag, int pair) {
    constexpr int SL = decltype(sl_tag)::value;
    float4* ptr = tile4 + lane_tile4 + cur_wy * TX + pair;   // row 2 * cur_wy + r: (2 cur_wy) * (TX / 2) = cur_wy * TX
#pragma unroll
    for (int k = 0; k < Y; ++k) {
      if (row_ok) {
        float4 tv = ptr[k * (ncell / 2)];
        tv.x += are[SL][k].x; tv.y += aim[SL][k].x; tv.z += are[SL][k].y; tv.w += aim[SL][k].y;
        ptr[k * (ncell / 2)] = tv;
      }
      are[SL][k] = make_float2(0.f, 0.f);
      aim[SL][k] = make_float2(0.f, 0.f);
    }
  };
  auto flush_pair = [&](int pair) {
    switch (pair & 3) {
      case 0: flush_slot(std::integral_constant<int, 0>{}, pair); break;
      case 1: flush_slot(std::integral_constant<int, 1>{}, pair); break;
      case 2: flush_slot(std::integral_constant<int, 2>{}, pair); break;
      default: flush_slot(std::integral_constant<int, 3>{}, pair); break;
    }
  };
  // a run with window `win` starts: retire the cell pairs the open window leaves behind
  auto open_window = [&](int win) {
    if (have) {
      int n = 4;
      if (win >= 0 && (win / kWinStride) == cur_wy) n = min(4, (win % kWinStride) - cur_wx);
      for (int i = 0; i < n; ++i) flush_pair(cur_wx + i);
      __syncwarp();
    }
    have = win >= 0;
    if (have) { cur_wx = win % kWinStride; cur_wy = win / kWinStride; }
  };

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    stage_write(bb);
    __syncwarp();
    if (bb + 1 < nbatch) fetch(bb + 1);

    const int cnt = min(BS, np - bb * BS);
    // The stage holds BS + 1 records so that the prefetch of point p + 1 never needs a guard.
    float4 wxa = *reinterpret_cast<const float4*>(stage);
    float4 wxb = *reinterpret_cast<const float4*>(stage + 4);
    float wy = stage[Rec::kWy + r];
    int hdr = __float_as_int(stage[Rec::kHdr]);
    float2 cc[Y];
#pragma unroll
    for (int k = 0; k < Y; ++k) cc[k] = *reinterpret_cast<const float2*>(stage + Rec::kC + 2 * (cg * Y + k));
#pragma unroll 2
    for (int p = 0; p < cnt; ++p) {
      const float* nxt = stage + (p + 1) * SW;
      const float4 xa = wxa, xb = wxb;
      const float wy_c = wy;
      const int hdr_c = hdr;
      float2 cc_c[Y];
#pragma unroll
      for (int k = 0; k < Y; ++k) cc_c[k] = cc[k];
      wxa = *reinterpret_cast<const float4*>(nxt);
      wxb = *reinterpret_cast<const float4*>(nxt + 4);
      wy = nxt[Rec::kWy + r];
      hdr = __float_as_int(nxt[Rec::kHdr]);
#pragma unroll
      for (int k = 0; k < Y; ++k) cc[k] = *reinterpret_cast<const float2*>(nxt + Rec::kC + 2 * (cg * Y + k));
      if (hdr_c & 1) open_window(hdr_c >> 1);   // warp-uniform
#pragma unroll
      for (int k = 0; k < Y; ++k) {
        const float cre = cc_c[k].x * wy_c, cim = cc_c[k].y * wy_c;
        fma_pair<PACK>(are[0][k], cre, make_float2(xa.x, xa.y));
        fma_pair<PACK>(aim[0][k], cim, make_float2(xa.x, xa.y));
        fma_pair<PACK>(are[1][k], cre, make_float2(xa.z, xa.w));
        fma_pair<PACK>(aim[1][k], cim, make_float2(xa.z, xa.w));
        fma_pair<PACK>(are[2][k], cre, make_float2(xb.x, xb.y));
        fma_pair<PACK>(aim[2][k], cim, make_float2(xb.x, xb.y));
        fma_pair<PACK>(are[3][k], cre, make_float2(xb.z, xb.w));
        fma_pair<PACK>(aim[3][k], cim, make_float2(xb.z, xb.w));
      }
    }
    __syncwarp();
  }
  open_window(-1);   // retire the last window

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
  __syncwarp();
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
