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
//     slot c & 3), so the inner loop has no phase, and a window step moves only ONE cell pair: the
//     outgoing pair is stored to the tile and the incoming pair's tile value is loaded as the new
//     accumulator start (the accumulators have the tile's (re, im) layout: one 128-bit store and one
//     128-bit load per coil, no adds, no dependent load-add-store chain).
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

// (re, im) += w * (cr, ci): one packed FFMA2 (fma.rn.f32x2; ptxas folds the {w, w} pack into the
// instruction's scalar-broadcast operand form), or two FFMA.
template <int PACK>
__device__ __forceinline__ void fma_cell(float& re, float& im, float w, float2 c) {
  if (PACK) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(re), "+f"(im) : "f"(w), "f"(c.x), "f"(c.y));
  } else {
    re = fmaf(w, c.x, re);
    im = fmaf(w, c.y, im);
  }
}

// c [T][M] -> c_pm [M][T] (complex64), 32 x 32 tiles through shared memory: both sides coalesced.
__global__ void __launch_bounds__(256)
transpose_strengths_kernel(const float2* __restrict__ c, float2* __restrict__ c_pm, int64_t M, int T) {
  __shared__ float2 tile[32][33];
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int t0 = blockIdx.y * 32;
  for (int ty = threadIdx.y; ty < 32; ty += 8) {
    const int t = t0 + ty;
    const int64_t i = i0 + threadIdx.x;
    if (t < T && i < M) tile[ty][threadIdx.x] = c[static_cast<int64_t>(t) * M + i];
  }
  __syncthreads();
  for (int iy = threadIdx.y; iy < 32; iy += 8) {
    const int64_t i = i0 + iy;
    const int t = t0 + threadIdx.x;
    if (t < T && i < M) c_pm[i * T + t] = tile[threadIdx.x][iy];
  }
}

template <int Y>
inline size_t spread_sweep2d_smem_bytes(const int* bin) {
  const size_t ncell = static_cast<size_t>(bin[0] + kSweepHaloX) * (bin[1] + 8);
  return 4 * Y * ncell * sizeof(float2) + 35 * SweepRec2<Y>::kStride * sizeof(float);
}

// One warp per (subproblem, group of NC = 4 Y coils). lane = g * 8 + r: r = row of the window,
// g = coil sub-group (coils g * Y .. g * Y + Y - 1 of the CTA's group).
template <int NS, int Y, int PACK, int PM>
__global__ void __launch_bounds__(32)
spread_sweep2d_f32_kernel(int64_t M, GridGeom g, int ngroups, const int* __restrict__ sub_total,
                          const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                          const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][4]*/,
                          const float2* __restrict__ c_in /* PM: point-major [M][ngroups * NC], else [T][M] */,
                          float2* __restrict__ fw, const __grid_constant__ CUtensorMap tmap_out, int use_tma) {
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
  float* stage = reinterpret_cast<float*>(smem4 + NC * (ncell / 2));   // [BS + 3][SW]

  for (int i = lane; i < NC * (ncell / 2); i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int r = lane & 7;
  const int cg = lane >> 3;
  const bool row_ok = r < NS + 1;   // rows of the even-aligned window that can carry weight
  // this lane's row r of coil (cg * Y + k): float4 index of cell pair 0 of tile row 0
  const int lane_tile4 = (cg * Y) * (ncell / 2) + r * (TX / 2);

  float2* fwt = fw + static_cast<int64_t>(t) * NC * g.nftot;
  // Strength rows: a point's NC strengths are NC * 8 contiguous bytes of c_pm; LPP lanes fetch one
  // point (16 bytes each), so one load instruction covers PPI points and touches PPI cache lines.
  constexpr int LPP = NC / 2;          // lanes per point
  constexpr int PPI = 32 / LPP;        // points per load instruction
  const int Ttot = ngroups * NC;
  const float4* crow = reinterpret_cast<const float4*>(c_in + static_cast<int64_t>(t) * NC) + (lane % LPP);
  const float2* ct = c_in + static_cast<int64_t>(t) * NC * M;   // !PM: coil-major strengths of this group

  // ---- register prefetch of this lane's point of the next batch ----
  float4 w4[4];
  int4 st_n = make_int4(0, 0, 0, 0);
  float4 c_n[LPP];   // c_n[q]: 16 bytes (2 coils) of point q * PPI + lane / LPP of the batch
#pragma unroll
  for (int q = 0; q < LPP; ++q) c_n[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  int id_n2 = 0;
  auto fetch = [&](int bb) {
    const int pl = bb * BS + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < 4; ++k) w4[k] = wrec4[j * 4 + k];
      st_n = start[j];
    }
    if (PM) {
#pragma unroll
      for (int q = 0; q < LPP; ++q) {
        const int pq = q * PPI + lane / LPP;                     // point of the batch this lane fetches a piece of
        const int idq = __shfl_sync(0xffffffffu, id_n2, pq);
        if (bb * BS + pq < np) c_n[q] = crow[static_cast<int64_t>(idq) * (Ttot / 2)];
      }
    } else if (pl < np) {   // coil-major: NC scattered 8-byte gathers for this lane's own point
#pragma unroll
      for (int q = 0; q < LPP; ++q) {
        const float2 a = ct[static_cast<int64_t>(2 * q) * M + id_n2], b2 = ct[static_cast<int64_t>(2 * q + 1) * M + id_n2];
        c_n[q] = make_float4(a.x, a.y, b2.x, b2.y);
      }
    }
    const int pl2 = (bb + 1) * BS + lane;
    if (pl2 < np) id_n2 = idx[p0 + pl2];
  };
  int last_win = -2;
  unsigned run_mask = 0;   // bit p: point p of the staged batch opens a new run (warp-uniform)
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
    run_mask = __ballot_sync(0xffffffffu, flag);
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
    for (int q = 0; q < LPP; ++q) {
      if (PM) {
        const int pq = q * PPI + lane / LPP;
        const int wq = __shfl_sync(0xffffffffu, win, pq);          // dropped point: zero strengths
        reinterpret_cast<float4*>(stage + pq * SW + Rec::kC)[lane % LPP] =
            wq >= 0 ? c_n[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        rec4[5 + q] = win >= 0 ? c_n[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  if (lane < np) id_n2 = idx[p0 + lane];
  fetch(0);

  // ---- accumulators: slot s = the tile cell pair with (pair & 3) == s, held in the TILE's own layout
  // (re_a, im_a, re_b, im_b) per coil, so that a slot moves to / from the tile with one 128-bit
  // access and no arithmetic: a window step STORES the outgoing pair and LOADS the incoming one
  // (whose tile value becomes the accumulator's start value) instead of read-add-write. ----
  float4 acc[4][Y];
#pragma unroll
  for (int sl = 0; sl < 4; ++sl)
#pragma unroll
    for (int k = 0; k < Y; ++k) acc[sl][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur_wx = 0, cur_wy = 0;
  bool have = false;

  auto store_slot = [&](auto sl_tag, int pair) {
    constexpr int SL = decltype(sl_tag)::value;
    float4* ptr = tile4 + lane_tile4 + cur_wy * TX + pair;   // row 2 * cur_wy + r: (2 cur_wy) * (TX / 2) = cur_wy * TX
    if (row_ok) {
#pragma unroll
      for (int k = 0; k < Y; ++k) ptr[k * (ncell / 2)] = acc[SL][k];
    }
  };
  auto load_slot = [&](auto sl_tag, int pair) {
    constexpr int SL = decltype(sl_tag)::value;
    const float4* ptr = tile4 + lane_tile4 + cur_wy * TX + pair;
    if (row_ok) {   // rows without weight keep their zero accumulators and are never stored
#pragma unroll
      for (int k = 0; k < Y; ++k) acc[SL][k] = ptr[k * (ncell / 2)];
    }
  };
  auto store_pair = [&](int pair) {
    switch (pair & 3) {
      case 0: store_slot(std::integral_constant<int, 0>{}, pair); break;
      case 1: store_slot(std::integral_constant<int, 1>{}, pair); break;
      case 2: store_slot(std::integral_constant<int, 2>{}, pair); break;
      default: store_slot(std::integral_constant<int, 3>{}, pair); break;
    }
  };
  auto load_pair = [&](int pair) {
    switch (pair & 3) {
      case 0: load_slot(std::integral_constant<int, 0>{}, pair); break;
      case 1: load_slot(std::integral_constant<int, 1>{}, pair); break;
      case 2: load_slot(std::integral_constant<int, 2>{}, pair); break;
      default: load_slot(std::integral_constant<int, 3>{}, pair); break;
    }
  };
  // A run with window `win` starts (-1: none). Same sweep (same window row): the window moved d
  // pairs to the right; the d outgoing pairs go back to the tile, the d incoming ones are loaded.
  // New sweep: all four pairs are stored, and the new window's four pairs are loaded after a warp
  // barrier (another lane owned those tile rows in the previous sweep).
  auto open_window = [&](int win) {
    const int nwx = win % kWinStride, nwy = win / kWinStride;
    const bool same_sweep = have && win >= 0 && nwy == cur_wy;
    if (same_sweep) {
      const int d = min(4, nwx - cur_wx);
      for (int i = 0; i < d; ++i) store_pair(cur_wx + i);
      for (int i = 0; i < d; ++i) load_pair(nwx + 4 - d + i);
      cur_wx = nwx;
      return;
    }
    if (have) {
      for (int i = 0; i < 4; ++i) store_pair(cur_wx + i);
    }
    __syncwarp();
    have = win >= 0;
    if (have) {
      cur_wx = nwx;
      cur_wy = nwy;
      for (int i = 0; i < 4; ++i) load_pair(cur_wx + i);
    }
  };

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    stage_write(bb);
    __syncwarp();
    if (bb + 1 < nbatch) fetch(bb + 1);

    const int cnt = min(BS, np - bb * BS);
    // Three points in flight: the shared-memory loads of point p + 3 are issued right after point p
    // has been consumed, two points (~45 instructions) ahead of their first use; with 7 one-warp
    // CTAs per SM nothing else hides the load latency. The stage holds BS + 3 records so that the
    // look-ahead never needs a guard. The run flags travel as a ballot mask, the header word is
    // read only when a run starts.
    struct PRec { float4 xa, xb; float wy; float2 cc[Y]; };
    auto ld = [&](PRec& R, int p) {
      const float* rec = stage + p * SW;
      R.xa = *reinterpret_cast<const float4*>(rec);
      R.xb = *reinterpret_cast<const float4*>(rec + 4);
      R.wy = rec[Rec::kWy + r];
      if (Y == 2) {
        const float4 c4 = *reinterpret_cast<const float4*>(rec + Rec::kC + 4 * cg);
        R.cc[0] = make_float2(c4.x, c4.y);
        R.cc[Y - 1] = make_float2(c4.z, c4.w);
      } else {
#pragma unroll
        for (int k = 0; k < Y; ++k) R.cc[k] = *reinterpret_cast<const float2*>(rec + Rec::kC + 2 * (cg * Y + k));
      }
    };
    auto comp = [&](const PRec& R, int p) {
      if ((run_mask >> p) & 1u) open_window(__float_as_int(stage[p * SW + Rec::kHdr]) >> 1);   // warp-uniform
#pragma unroll
      for (int k = 0; k < Y; ++k) {
        const float2 cw = make_float2(R.cc[k].x * R.wy, R.cc[k].y * R.wy);
        // cell (re, im) += w_cell * (Re, Im)(c wy): packed FFMA2 with a scalar-broadcast weight
        fma_cell<PACK>(acc[0][k].x, acc[0][k].y, R.xa.x, cw);
        fma_cell<PACK>(acc[0][k].z, acc[0][k].w, R.xa.y, cw);
        fma_cell<PACK>(acc[1][k].x, acc[1][k].y, R.xa.z, cw);
        fma_cell<PACK>(acc[1][k].z, acc[1][k].w, R.xa.w, cw);
        fma_cell<PACK>(acc[2][k].x, acc[2][k].y, R.xb.x, cw);
        fma_cell<PACK>(acc[2][k].z, acc[2][k].w, R.xb.y, cw);
        fma_cell<PACK>(acc[3][k].x, acc[3][k].y, R.xb.z, cw);
        fma_cell<PACK>(acc[3][k].z, acc[3][k].w, R.xb.w, cw);
      }
    };
    PRec A, B, C;
    ld(A, 0);
    ld(B, 1);
    ld(C, 2);
    int p = 0;
    for (; p + 3 <= cnt; p += 3) {
      comp(A, p);
      ld(A, p + 3);
      comp(B, p + 1);
      ld(B, p + 4);
      comp(C, p + 2);
      ld(C, p + 5);
    }
    if (p < cnt) {
      comp(A, p);
      if (p + 1 < cnt) comp(B, p + 1);
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
