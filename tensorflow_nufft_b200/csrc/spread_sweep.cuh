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
// The tile pitch is bin_x + 10 cells = an odd number of 16-byte cell pairs, which makes the 8
// row-lanes of a quarter warp hit 8 distinct 16-byte bank groups.
#pragma once
#include <type_traits>

#include "dev_common.cuh"
#include "interp.cuh"
#include "points.cuh"
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

// The same with 128-bit accesses on both sides: 64 points x 32 coils per CTA, a lane reads two
// adjacent points of one coil and writes two adjacent coils of one point (needs M even, T even and a
// 16-byte aligned source). Shared tile [point][coil pair ^ (point / 2 & 7)]: the float4 reads of
// the write-out are conflict free, the transposing 8-byte stores two-way.
__global__ void __launch_bounds__(256)
transpose_strengths_wide_kernel(const float2* __restrict__ c, float2* __restrict__ c_pm, int64_t M, int T) {
  __shared__ __align__(16) float2 tile[64 * 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int t0 = blockIdx.y * 32;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = t0 + (tid >> 5) + 8 * k;
    const int64_t i = i0 + 2 * lane;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T && i < M) v[k] = *reinterpret_cast<const float4*>(c + static_cast<int64_t>(t) * M + i);   // M even: i + 1 < M
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ty = (tid >> 5) + 8 * k;
    const int r = 2 * lane;   // rows r, r + 1 share the swizzle (r / 2 & 7)
    const int col = 2 * ((ty >> 1) ^ (lane & 7)) + (ty & 1);
    tile[r * 32 + col] = make_float2(v[k].x, v[k].y);
    tile[(r + 1) * 32 + col] = make_float2(v[k].z, v[k].w);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int iy = (tid >> 4) + 16 * k, q = tid & 15;
    const int64_t i = i0 + iy;
    const int t = t0 + 2 * q;
    if (i < M && t < T) {   // T even: t + 1 < T
      const float4 w = *reinterpret_cast<const float4*>(&tile[iy * 32 + 2 * (q ^ ((iy >> 1) & 7))]);
      *reinterpret_cast<float4*>(c_pm + i * T + t) = w;
    }
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

  // ---- register prefetch of this lane's point of the next batch (a second set, two batches deep as
  // in the 3D kernel, was measured: 0.99 vs 0.96 ms on cfg2 -- 196 registers) ----
  struct Pre {
    float4 w4[4];
    int4 st;
    float4 c[LPP];   // c[q]: 16 bytes (2 coils) of point q * PPI + lane / LPP of the batch
    int id;          // point id of the batch this set fetches next
  };
  Pre P0;
  auto fetch = [&](Pre& P, int bb) {
    const int pl = bb * BS + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < 4; ++k) P.w4[k] = wrec4[j * 4 + k];
      P.st = start[j];
    }
    if (PM) {
#pragma unroll
      for (int q = 0; q < LPP; ++q) {
        const int pq = q * PPI + lane / LPP;                     // point of the batch this lane fetches a piece of
        const int idq = __shfl_sync(0xffffffffu, P.id, pq);
        if (bb * BS + pq < np) P.c[q] = crow[static_cast<int64_t>(idq) * (Ttot / 2)];
      }
    } else if (pl < np) {   // coil-major: NC scattered 8-byte gathers for this lane's own point
#pragma unroll
      for (int q = 0; q < LPP; ++q) {
        const float2 a = ct[static_cast<int64_t>(2 * q) * M + P.id], b2 = ct[static_cast<int64_t>(2 * q + 1) * M + P.id];
        P.c[q] = make_float4(a.x, a.y, b2.x, b2.y);
      }
    }
    const int pl2 = (bb + 1) * BS + lane;
    if (pl2 < np) P.id = idx[p0 + pl2];
  };
  int last_win = -2;
  unsigned run_mask = 0;   // bit p: point p of the staged batch opens a new run (warp-uniform)
  auto stage_write = [&](Pre& P, int bb) {
    float4 (&w4)[4] = P.w4;
    float4 (&c_n)[LPP] = P.c;
    const int pl = bb * BS + lane;
    float* rec = stage + lane * SW;
    int win = -1;
    if (pl < np) {
      const int rx = P.st.x - ox, ry = P.st.y - oy;
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
#pragma unroll
  for (int q = 0; q < LPP; ++q) P0.c[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  P0.st = make_int4(0, 0, 0, 0);
  P0.id = lane < np ? idx[p0 + lane] : 0;
  fetch(P0, 0);

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

  // Slot sl holds the tile cell pair P in [wx, wx + 4) with (P & 3) == sl.
  auto pair_of = [](int wx, int sl) { return wx + ((sl - wx) & 3); };
  float4* rowp = tile4 + lane_tile4;   // this lane's row of the open sweep, coil k at rowp[k * (ncell / 2) + pair]
  // A run with window `win` starts (-1: none). Same sweep (same window row): the window moved d
  // pairs to the right; every slot computes which pair it holds and whether that pair leaves the
  // window -- then it is stored and the pair of the new window that takes the slot is loaded
  // (predicated 128-bit accesses, no branch on the slot index). New sweep: all four pairs are
  // stored, and the new window's four pairs are loaded after a warp barrier (another lane owned
  // those tile rows in the previous sweep).
  auto open_window = [&](int win) {
    const int nwx = win % kWinStride, nwy = win / kWinStride;
    if (have && win >= 0 && nwy == cur_wy) {
      const int d = nwx - cur_wx;
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int P = pair_of(cur_wx, sl);
        if (row_ok && P - cur_wx < d) {
          const int Q = pair_of(nwx, sl);
#pragma unroll
          for (int k = 0; k < Y; ++k) {
            rowp[k * (ncell / 2) + P] = acc[sl][k];
            acc[sl][k] = rowp[k * (ncell / 2) + Q];
          }
        }
      }
      cur_wx = nwx;
      return;
    }
    if (have && row_ok) {
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int P = pair_of(cur_wx, sl);
#pragma unroll
        for (int k = 0; k < Y; ++k) rowp[k * (ncell / 2) + P] = acc[sl][k];
      }
    }
    __syncwarp();
    have = win >= 0;
    if (!have) return;
    cur_wx = nwx;
    cur_wy = nwy;
    rowp = tile4 + lane_tile4 + cur_wy * TX;   // row 2 * cur_wy + r: (2 cur_wy) * (TX / 2) = cur_wy * TX
    if (row_ok) {   // rows without weight keep their zero accumulators and are never stored
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int P = pair_of(cur_wx, sl);
#pragma unroll
        for (int k = 0; k < Y; ++k) acc[sl][k] = rowp[k * (ncell / 2) + P];
      }
    }
  };

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    stage_write(P0, bb);
    __syncwarp();
    if (bb + 1 < nbatch) fetch(P0, bb + 1);
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

// ------------------------------------------------------------------------------------------------
// 3D: the same sweep with one coil per CTA. lane = zg * 8 + r: r = row (y) of the 8 x 8 x 8 window
// (even-aligned in all three dimensions: the record kernel shifts and zero-pads wx, wy AND wz),
// zg = pair of z-planes (2 zg, 2 zg + 1). A lane holds 8 x-cells x 2 planes = 32 accumulator floats,
// rotating along x. Points are sorted by (bin, window z, window y, window x); a sweep is a run of
// windows with the same (z, y). Per point: 2 x LDS.128 (x-weights, warp-uniform), y-weight, the
// lane's two z-weights (LDS.64), the strength (LDS.64), 6 FMUL, 16 FFMA2 -- against a 128-bit
// read-modify-write of the shared-memory tile per lane and z-plane in spread.cuh (7 per point).
// ------------------------------------------------------------------------------------------------
struct SweepRec3 {
  // words: [0..7] x-weights rotated to slots, [8..15] wy, [16..23] wz, [24] header, [26..27] strength
  static constexpr int kWy = 8, kWz = 16, kHdr = 24, kC = 26;
  static constexpr int kStride = 28;   // 4 * 7 words: conflict-free staging
};

constexpr int kSweepRing = 8;   // z-planes of the tile resident in shared memory (one window depth)

inline size_t spread_sweep3d_smem_bytes(const int* bin) {
  const size_t plane = static_cast<size_t>(bin[0] + kSweepHaloX) * (bin[1] + 8);
  return kSweepRing * plane * sizeof(float2) + 35 * SweepRec3::kStride * sizeof(float);
}

// Z-SLAB STREAMING: the sort order is (bin, window z, window y, window x), so the window's z start
// never decreases inside a subproblem. Shared memory therefore holds only a RING of 8 z-planes of
// the (bin + halo) tile -- the planes the current window depth [2 wz, 2 wz + 8) covers -- instead
// of all bin_z + 8: when wz advances, the planes left behind are complete, leave through the TMA
// unit (one reduce-add per plane) and are cleared for reuse. 27 KB per one-warp CTA instead of 53,
// i.e. 7 resident warps per SM instead of 4 (the kernel is latency-bound), whatever the bin depth.
// OTF = 1 (opt-in, single-transform plans): the stencil weights are evaluated HERE, in the staging
// step, from the folded coordinates (same functions as stencil_record8_kernel, so the same bits)
// instead of being written by set_points (96 + 16 bytes per point) and read back once. Measured on
// cfg3: set_points 1.18 -> 0.71 ms, but the spreader 1.29 -> 4.85 ms (a lane evaluates its point's
// 24 kernel values alone: ~1400 serial instructions per batch in a one-warp CTA), so it is OFF.
template <int NS, int PACK, int OTF>
__global__ void __launch_bounds__(32)
spread_sweep3d_f32_kernel(int64_t M, GridGeom g, int ntr, const int* __restrict__ sub_total,
                          const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                          const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][6]*/,
                          const float4* __restrict__ folded4 /*[M]: x, y, z, 0 (OTF)*/, float es_beta, float es_c, float es_hw,
                          const float2* __restrict__ c, float2* __restrict__ fw,
                          const __grid_constant__ CUtensorMap tmap_out, int use_tma) {
  static_assert(NS <= 7, "8-cell windows");
  using Rec = SweepRec3;
  constexpr int SW = Rec::kStride;
  constexpr int BS = 32;
  constexpr int kWinStride = 64;   // window id = (wz * 64 + wy) * 64 + wx
  constexpr int RING = kSweepRing;
  extern __shared__ __align__(128) float4 smem4[];

  const int s = blockIdx.x / ntr;
  const int t = blockIdx.x - s * ntr;
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int lane = threadIdx.x;
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + kSweepHaloX, TY = g.bin[1] + 8, TZ = g.bin[2] + 8;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = b / (g.nbins[0] * g.nbins[1]);
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = bz * g.bin[2] - 4;
  const int plane4 = TX * TY / 2;   // float4 per z-plane
  float4* tile4 = smem4;            // [RING][plane4]: tile plane z lives in ring slot z & 7
  float* stage = reinterpret_cast<float*>(smem4 + RING * plane4);   // [BS + 3][SW]
  const bool interior = use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
                        oz >= 0 && oz + TZ <= g.nf[2];

  for (int i = lane; i < RING * plane4; i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int r = lane & 7;
  const int zg = lane >> 3;
  const bool ok0 = r < NS + 1 && 2 * zg < NS + 1, ok1 = r < NS + 1 && 2 * zg + 1 < NS + 1;
  const int lane_row4 = r * (TX / 2);

  const float2* ct = c + static_cast<int64_t>(t) * M;
  float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;

  // ---- register prefetch, two batches deep (see the 2D kernel) ----
  struct Pre {
    float4 w4[OTF ? 1 : 6];   // OTF: w4[0] = folded coordinates (x, y, z, 0)
    int4 st;
    float2 c;
    int id;   // point id of the batch this set fetches NEXT
  };
  Pre P0, P1;
  auto fetch = [&](Pre& P, int bb) {
    const int pl = bb * BS + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
      if (OTF) {
        P.w4[0] = folded4[P.id];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) P.w4[k] = wrec4[j * 6 + k];
        P.st = start[j];
      }
      P.c = ct[P.id];
    }
    const int pl2 = (bb + 2) * BS + lane;
    if (pl2 < np) P.id = idx[p0 + pl2];
  };
  int last_win = -2;
  unsigned run_mask = 0;
  auto stage_write = [&](Pre& P, int bb) {
    float4 w4[6];
    int4 st_c = P.st;
    if (OTF) {
      // stencil start (moved down to an even cell) and the 8 zero-padded, shifted weights per dimension
      const float xs[3] = {P.w4[0].x, P.w4[0].y, P.w4[0].z};
      int sts[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int i1 = static_cast<int>(ceilf(sub_rn(xs[d], es_hw)));
        const float x1 = sub_rn(static_cast<float>(i1), xs[d]);
        const int shift = i1 & 1;
        sts[d] = i1 - shift;
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = es_eval_fast<float>(add_rn(x1, static_cast<float>(k - shift)), es_beta, es_c, es_hw);
        w4[2 * d] = make_float4(w[0], w[1], w[2], w[3]);
        w4[2 * d + 1] = make_float4(w[4], w[5], w[6], w[7]);
      }
      st_c = make_int4(sts[0], sts[1], sts[2], 0);
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) w4[k] = P.w4[OTF ? 0 : k];
    }
    float2 c_n = P.c;
    const int pl = bb * BS + lane;
    float* rec = stage + lane * SW;
    int win = -1;
    if (pl < np) {
      const int rx = st_c.x - ox, ry = st_c.y - oy, rz = st_c.z - oz;
      // Memory safety for coordinates outside the declared points_range (see the 2D kernel).
      const bool fits = rx >= 0 && rx + 8 <= TX && ry >= 0 && ry + NS + 1 <= TY && rz >= 0 && rz + NS + 1 <= TZ &&
                        ((rx | ry | rz) & 1) == 0;
      if (fits) win = ((rz >> 1) * kWinStride + (ry >> 1)) * kWinStride + (rx >> 1);
    }
    const int prev = __shfl_up_sync(0xffffffffu, win, 1);
    const int flag = (lane == 0 ? (win != last_win) : (win != prev)) ? 1 : 0;
    last_win = __shfl_sync(0xffffffffu, win, BS - 1);
    run_mask = __ballot_sync(0xffffffffu, flag);
    const int rot = win & 3;
    if (win < 0) {
      w4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      w4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      c_n = make_float2(0.f, 0.f);
    }
    float2* wxs = reinterpret_cast<float2*>(rec);
    wxs[(rot + 0) & 3] = make_float2(w4[0].x, w4[0].y);
    wxs[(rot + 1) & 3] = make_float2(w4[0].z, w4[0].w);
    wxs[(rot + 2) & 3] = make_float2(w4[1].x, w4[1].y);
    wxs[(rot + 3) & 3] = make_float2(w4[1].z, w4[1].w);
    float4* rec4 = reinterpret_cast<float4*>(rec);
    rec4[2] = w4[2];
    rec4[3] = w4[3];
    rec4[4] = w4[4];
    rec4[5] = w4[5];
    rec4[6] = make_float4(__int_as_float((win << 1) | flag), 0.f, c_n.x, c_n.y);
  };
  P0.st = P1.st = make_int4(0, 0, 0, 0);
  P0.c = P1.c = make_float2(0.f, 0.f);
  P0.id = lane < np ? idx[p0 + lane] : 0;
  P1.id = BS + lane < np ? idx[p0 + BS + lane] : 0;
  fetch(P0, 0);
  if (np > BS) fetch(P1, 1);

  // ---- accumulators: [slot][plane], tile layout (re_a, im_a, re_b, im_b) ----
  float4 acc[4][2];
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) { acc[sl][0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[sl][1] = make_float4(0.f, 0.f, 0.f, 0.f); }
  int cur_wx = 0, cur_wy = 0, cur_wz = -1;
  int zbase = -1;           // tile planes [zbase, zbase + 8) are resident (zbase = 2 * window z of the open depth)
  float4* row0 = tile4;     // this lane's row in its two planes of the open window: row0[pair], row1[pair]
  float4* row1 = tile4;
  bool have = false;

  // Retires the tile planes [z0, z1): adds them to the fine grid and clears their ring slots.
  auto retire_planes = [&](int z0, int z1) {
    if (z1 <= z0) return;
    __syncwarp();
    if (interior) {
      fence_proxy_async_smem();   // every lane: its generic-proxy tile writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        for (int z = z0; z < z1; ++z) tma_reduce_add_4d(&tmap_out, tile4 + (z & (RING - 1)) * plane4, 2 * ox, oy, oz + z, t);
        tma_store_commit_and_wait_read();   // the slots are cleared next: the TMA unit must have read them
      }
      __syncwarp();
    } else {
      const int TXH = TX / 2;
      for (int z = z0; z < z1; ++z) {
        const float4* pl = tile4 + (z & (RING - 1)) * plane4;
        const int gz = mod_idx(oz + z, g.nf[2]);
        for (int i = lane; i < plane4; i += 32) {
          const float4 v = pl[i];
          if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
          const int gx = mod_idx(ox + 2 * (i % TXH), g.nf[0]);
          const int gy = mod_idx(oy + i / TXH, g.nf[1]);
          red_add(reinterpret_cast<float4*>(fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx), v);
        }
      }
      __syncwarp();
    }
    for (int z = z0; z < z1; ++z) {
      float4* pl = tile4 + (z & (RING - 1)) * plane4;
      for (int i = lane; i < plane4; i += 32) pl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
  };

  // Slot sl holds the tile cell pair P in [wx, wx + 4) with (P & 3) == sl.
  auto pair_of = [](int wx, int sl) { return wx + ((sl - wx) & 3); };
  auto store_all = [&]() {
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int P = pair_of(cur_wx, sl);
      if (ok0) row0[P] = acc[sl][0];
      if (ok1) row1[P] = acc[sl][1];
    }
  };
  auto load_all = [&]() {
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int P = pair_of(cur_wx, sl);
      if (ok0) acc[sl][0] = row0[P];
      if (ok1) acc[sl][1] = row1[P];
    }
  };
  // A run with window `win` starts (-1: none). Branch-free on the slot index: every slot computes
  // which pair it holds and whether that pair leaves the window (predicated 128-bit store + load).
  auto open_window = [&](int win) {
    const int nwx = win % kWinStride, nwyz = win / kWinStride;
    const int nwy = nwyz % kWinStride, nwz = nwyz / kWinStride;
    if (have && win >= 0 && nwy == cur_wy && nwz == cur_wz) {   // same sweep: the window moved right
      const int d = nwx - cur_wx;
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int P = pair_of(cur_wx, sl);
        if (P - cur_wx < d) {             // leaves: store it, load the pair of the new window that takes the slot
          const int Q = pair_of(nwx, sl);
          if (ok0) { row0[P] = acc[sl][0]; acc[sl][0] = row0[Q]; }
          if (ok1) { row1[P] = acc[sl][1]; acc[sl][1] = row1[Q]; }
        }
      }
      cur_wx = nwx;
      return;
    }
    if (have) store_all();
    __syncwarp();   // other lanes owned the new sweep's rows before
    have = win >= 0;
    if (!have) return;
    if (zbase < 0) zbase = 2 * nwz;
    if (2 * nwz > zbase) {   // deeper window: the planes left behind are complete
      retire_planes(zbase, min(2 * nwz, zbase + RING));
      zbase = 2 * nwz;
    }
    cur_wx = nwx;
    cur_wy = nwy;
    cur_wz = nwz;
    const int zp = (2 * nwz + 2 * zg) & (RING - 1);   // even, so the lane's two planes are ring slots zp, zp + 1
    row0 = tile4 + zp * plane4 + 2 * nwy * (TX / 2) + lane_row4;
    row1 = row0 + plane4;
    load_all();
  };

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    // the two prefetch sets alternate; the per-point code below exists once
    if (bb & 1) {
      stage_write(P1, bb);
      __syncwarp();
      if (bb + 2 < nbatch) fetch(P1, bb + 2);
    } else {
      stage_write(P0, bb);
      __syncwarp();
      if (bb + 2 < nbatch) fetch(P0, bb + 2);
    }
    const int cnt = min(BS, np - bb * BS);
    struct PRec { float4 xa, xb; float wy; float2 wz; float2 cc; };
    auto ld = [&](PRec& R, int p) {
      const float* rec = stage + p * SW;
      R.xa = *reinterpret_cast<const float4*>(rec);
      R.xb = *reinterpret_cast<const float4*>(rec + 4);
      R.wy = rec[Rec::kWy + r];
      R.wz = *reinterpret_cast<const float2*>(rec + Rec::kWz + 2 * zg);
      R.cc = *reinterpret_cast<const float2*>(rec + Rec::kC);
    };
    auto comp = [&](const PRec& R, int p) {
      if ((run_mask >> p) & 1u) open_window(__float_as_int(stage[p * SW + Rec::kHdr]) >> 1);   // warp-uniform
      const float cre = R.cc.x * R.wy, cim = R.cc.y * R.wy;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float wz = k == 0 ? R.wz.x : R.wz.y;
        const float2 cw = make_float2(cre * wz, cim * wz);
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
  open_window(-1);   // stores the last window
  if (zbase >= 0) retire_planes(zbase, min(zbase + RING, TZ));
}

}  // namespace b200
