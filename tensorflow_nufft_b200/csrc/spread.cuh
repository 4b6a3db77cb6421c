// spread.cuh -- type-1 spreader: fw[g] += sum_j c_j prod_d phi(g_d - x_{j,d}) with periodic wrap.
// Behavioural reference: SpreadSubproblem{2,3}DKernel nufft_plan.cu.cc:790-878,1404-1510 and the
// CPU spread_subproblem_{1,2,3}d nufft_plan.cc:1463-1636 (exactly ns taps per dimension).
//
// Two engines:
//  * spread_global_kernel   point-driven, native vector reductions (REDG.ADD.F32x2 / F64) straight
//                           into the fine grid; any rank / width / precision. Fallback + cross-check.
//  * spread_tile_f32_kernel one warp = one subproblem (<= msub points of one bin) with a PRIVATE
//                           shared-memory tile (bin + halo). Lanes are laid over the stencil
//                           (row = lane / QX, cell pair = lane % QX), so every shared-memory update
//                           is a plain 128-bit load / FFMA / store with no atomics (shared-memory float
//                           atomicAdd is a CAS loop on sm_100: ATOMS.CAST.SPIN) and no bank
//                           conflicts (tile pitch = 8 mod 16 cells). The tile is flushed with
//                           REDG.E.ADD.F32x4 (two complex cells per reduction), zero cells skipped.
#pragma once
#include <cuda.h>

#include "dev_common.cuh"

namespace b200 {

struct GridGeom {
  int rank;
  int nf[3];
  int bin[3];
  int nbins[3];
  int64_t nftot;
};

__device__ __forceinline__ int wrap_idx(int g, int n) {
  g = g < 0 ? g + n : g;
  g = g >= n ? g - n : g;
  return g;
}
__device__ __forceinline__ int mod_idx(int g, int n) {
  int r = g % n;
  return r < 0 ? r + n : r;
}

// ---------------------------------------------------------------------------------------------
// Point-driven global-reduction spreader. One thread per (point, stencil row); loops over the
// transforms of the batch so the weights are read once.
// ---------------------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256)
spread_global_kernel(int64_t M, int ntr, GridGeom g, int ns, int R, int PX, int PY, const int* __restrict__ idx,
                     const int4* __restrict__ start, const F* __restrict__ wrec,
                     const Cplx<F>* __restrict__ c, Cplx<F>* __restrict__ fw) {
  const int rows = g.rank == 1 ? 1 : (g.rank == 2 ? ns : ns * ns);
  const int64_t total = M * rows;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int64_t j = w / rows;
    const int row = static_cast<int>(w - j * rows);
    const int dy = g.rank > 1 ? row % ns : 0;
    const int dz = g.rank > 2 ? row / ns : 0;
    const int4 st = start[j];
    const F* wx = wrec + j * R;
    const F* wy = wx + PX;
    const F* wz = wy + PY;
    F kyz = F(1);
    int64_t rowoff = 0;
    if (g.rank > 1) {
      kyz = wy[dy];
      rowoff = static_cast<int64_t>(mod_idx(st.y + dy, g.nf[1])) * g.nf[0];
    }
    if (g.rank > 2) {
      kyz = kyz * wz[dz];
      rowoff += static_cast<int64_t>(mod_idx(st.z + dz, g.nf[2])) * g.nf[0] * g.nf[1];
    }
    const int pid = idx[j];
    for (int t = 0; t < ntr; ++t) {
      const Cplx<F> cj = c[static_cast<int64_t>(t) * M + pid];
      Cplx<F>* out = fw + static_cast<int64_t>(t) * g.nftot + rowoff;
      for (int k = 0; k < ns; ++k) {
        const F kx = wx[st.w + k];
        const int gx = mod_idx(st.x + st.w + k, g.nf[0]);
        red_add(out + gx, make_cplx<F>(kyz * (cj.x * kx), kyz * (cj.y * kx)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tile spreader (complex64, ns <= 7, rank 2 or 3).
// Work item = (subproblem, transform); CTA = WPT warps sharing ONE tile of (bin + 8)^rank float2
// cells. WPT = 1 in 2D. In 3D the tile's z-planes are dealt round-robin to the WPT warps
// (plane p belongs to warp p % WPT), every warp walks all points of the subproblem and updates
// only the planes it owns: exclusive ownership, so still no atomics, with WPT times the warps
// per tile for latency hiding.
// Points are consumed in batches of 32*WPT through a double-buffered shared-memory stage: lane l
// of warp w fetches the record of point (32 w + l) of the NEXT batch from HBM (weights, start
// index, strength via idx) while the CTA processes the current one; at the batch boundary each
// lane writes {tile offset, c*wx[8], wy[8], wz[8]} for its point, one barrier, and the loop
// goes on. All shared accesses in the inner loop are conflict-free 128-bit (tile pitch = 8 mod
// 16 cells, stage record stride = 4 mod 32 words).
// ---------------------------------------------------------------------------------------------
// TMA reduce-add helpers (defined in interp.cuh, which includes this header)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2);
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3);
__device__ __forceinline__ void tma_store_commit_and_wait_read();
__device__ __forceinline__ void fence_proxy_async_smem();

template <int RANK> struct StageRec {
  // spread: words [0..7] wx, [8..23] cw[r] = {Re c * wy[r], Im c * wy[r]}, [24] tile offset (cells)
  //         or -1, [25] tile z of the stencil start, [26..27] pad, [28..35] wz (3D)
  // interp: words [0..7] wx, [8..15] wy, [16..23] wz, [24] tile offset or -1
  // Per-point broadcast data is read with 64/32-bit loads: a 128-bit shared load costs 4
  // wavefronts (one per quarter warp) even when all lanes share addresses.
  static constexpr int kWords = RANK == 3 ? 44 : 36;   // stride = 12 / 4 mod 32: conflict-free staging
};

template <int NS, int RANK, int WPT>
__global__ void __launch_bounds__(WPT * 32)
spread_tile_f32_kernel(int64_t M, GridGeom g, const int* __restrict__ sub_total,
                       const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                       const int4* __restrict__ start, const float4* __restrict__ wrec4 /*[M][2*RANK]*/,
                       const float2* __restrict__ c, float2* __restrict__ fw,
                       const __grid_constant__ CUtensorMap tmap_out, int use_tma, int zrange) {
  constexpr int QX = (NS + 2) / 2;      // float4 lanes per stencil row: covers NS+1 cells
  static_assert(QX * NS <= 32, "stencil slab must fit one warp");
  constexpr int C4 = 2 * RANK;          // float4 chunks per weight record
  constexpr int SW = StageRec<RANK>::kWords;
  constexpr int BS = 32;                // points per batch (staged by warp 0, one point per lane)
  constexpr int NBUF = WPT > 1 ? 2 : 1;
  extern __shared__ __align__(128) float4 smem4[];

  const int s = blockIdx.x;
  // the subproblem count and this CTA's descriptor are independent loads (the descriptor
  // array has an entry for every launched CTA): one global round trip instead of two
  const int nsub_live = *sub_total;
  const int4 sd = sub_desc[s];
  if (s >= nsub_live) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.y;
  const int b = sd.x, p0 = sd.y, np = sd.z;

  const int TX = g.bin[0] + 8, TY = g.bin[1] + 8;
  const int TZ = RANK > 2 ? g.bin[2] + 8 : 1;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = RANK > 2 ? bz * g.bin[2] - 4 : 0;
  const int ncell = TX * TY * TZ;
  float4* tile4 = smem4;
  float2* tile = reinterpret_cast<float2*>(tile4);
  float* stage = reinterpret_cast<float*>(smem4 + ncell / 2);   // [NBUF][BS][SW]

  for (int i = tid; i < ncell / 2; i += WPT * 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < NS;
  const int lane_off = r * TX + 2 * q;   // cells, within a z-plane, relative to the stencil start
  const int zstride4 = TY * TX / 2;

  const float2* ct = c + static_cast<int64_t>(t) * M;
  float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;

  // ---- software pipeline registers (warp 0): this lane's point of the next batch ----
  float4 w4[C4];
  int4 st_n = make_int4(0, 0, 0, 0);
  float2 c_n = make_float2(0.f, 0.f);
  int id_n2 = 0;
  auto fetch = [&](int bb) {
    // weights + start + strength (through the id fetched one call earlier) for batch bb, and the
    // point id for batch bb + 1
    const int pl = bb * BS + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < C4; ++k) w4[k] = wrec4[j * C4 + k];
      st_n = start[j];
      c_n = ct[id_n2];
    }
    const int pl2 = (bb + 1) * BS + lane;
    if (pl2 < np) id_n2 = idx[p0 + pl2];
  };
  auto stage_write = [&](int bb) {
    const int pl = bb * BS + lane;
    float4* rec4 = reinterpret_cast<float4*>(stage + ((bb % NBUF) * BS + lane) * SW);
    int off = -1, tz = 0;
    if (pl < np) {
      const int rx = st_n.x - ox, ry = st_n.y - oy, rz = RANK > 2 ? st_n.z - oz : 0;
      // Memory safety for coordinates outside the declared points_range: such a stencil does not
      // lie in this bin's tile and the point is dropped (the reference's behaviour is undefined).
      const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                        (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
      if (fits) { off = (rz * TY + ry) * TX + rx; tz = rz; }
    }
    rec4[0] = w4[0];
    rec4[1] = w4[1];
    const float cre = c_n.x, cim = c_n.y;
    rec4[2] = make_float4(cre * w4[2].x, cim * w4[2].x, cre * w4[2].y, cim * w4[2].y);
    rec4[3] = make_float4(cre * w4[2].z, cim * w4[2].z, cre * w4[2].w, cim * w4[2].w);
    rec4[4] = make_float4(cre * w4[3].x, cim * w4[3].x, cre * w4[3].y, cim * w4[3].y);
    rec4[5] = make_float4(cre * w4[3].z, cim * w4[3].z, cre * w4[3].w, cim * w4[3].w);
    rec4[6] = make_float4(__int_as_float(off), __int_as_float(tz), 0.f, 0.f);
    if (RANK > 2) { rec4[7] = w4[C4 - 2]; rec4[8] = w4[C4 - 1]; }
  };
  if (warp == 0) {
    if (lane < np) id_n2 = idx[p0 + lane];
    fetch(0);
  }

  const int nbatch = (np + BS - 1) / BS;
  for (int bb = 0; bb < nbatch; ++bb) {
    if (warp == 0) stage_write(bb);
    __syncthreads();   // stage visible (and, first time, the zeroed tile)
    if (warp == 0 && bb + 1 < nbatch) fetch(bb + 1);

    const float* sbuf = stage + (bb % NBUF) * BS * SW;
    const int cnt = min(BS, np - bb * BS);
    // Inner loop, software-pipelined by hand: the small per-point loads (header, wx pair, wy) of
    // point p+1 are issued before the tile read-modify-write of point p, so that the dependent
    // chain per point is only LDS.128 -> FFMA -> STS.128.
    const int rr = row_ok ? r : 0;
    float2 of = *reinterpret_cast<const float2*>(sbuf + 24);
    float2 wx = *reinterpret_cast<const float2*>(sbuf + 2 * q);
    float2 cw = *reinterpret_cast<const float2*>(sbuf + 8 + 2 * rr);
    for (int p = 0; p < cnt; ++p) {
      const float* rec = sbuf + p * SW;
      const float2 of_c = of, wx_c = wx, cw_c = cw;
      if (p + 1 < cnt) {
        const float* nxt = rec + SW;
        of = *reinterpret_cast<const float2*>(nxt + 24);
        wx = *reinterpret_cast<const float2*>(nxt + 2 * q);
        cw = *reinterpret_cast<const float2*>(nxt + 8 + 2 * rr);
      }
      const int off = __float_as_int(of_c.x);
      if (off >= 0) {
        const float4 cx = make_float4(cw_c.x * wx_c.x, cw_c.y * wx_c.x, cw_c.x * wx_c.y, cw_c.y * wx_c.y);
        if (RANK == 2) {
          if (row_ok) {
            float4* ptr = reinterpret_cast<float4*>(tile + off + lane_off);
            float4 v = *ptr;
            v.x += cx.x; v.y += cx.y; v.z += cx.z; v.w += cx.w;
            *ptr = v;
          }
        } else {
          const int tz = __float_as_int(of_c.y);
          // first stencil plane owned by this warp: (tz + dz) % WPT == warp   (WPT is a power of two)
          static_assert((WPT & (WPT - 1)) == 0, "WPT must be a power of two");
          const int dz0 = (warp - tz) & (WPT - 1);
          if (row_ok && dz0 < NS) {
            float4* ptr = reinterpret_cast<float4*>(tile + off + lane_off) + dz0 * zstride4;
            const float* wzp = rec + 28 + dz0;
            constexpr int KMAX = (NS + WPT - 1) / WPT;
            float4 v[KMAX];
            float w[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k == 0 || dz0 + k * WPT < NS) {
                v[k] = ptr[k * WPT * zstride4];
                w[k] = wzp[k * WPT];
              }
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k == 0 || dz0 + k * WPT < NS) {
                v[k].x += w[k] * cx.x; v[k].y += w[k] * cx.y; v[k].z += w[k] * cx.z; v[k].w += w[k] * cx.w;
                ptr[k * WPT * zstride4] = v[k];
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (NBUF == 1) __syncwarp();
  }
  __syncthreads();

  // Flush. Interior tiles: ONE TMA reduce-add (the TMA unit reads the tile and adds it to the fine
  // grid in L2). Tiles that straddle the periodic boundary: two complex cells per
  // REDG.ADD.F32x4 with index wrap, untouched (zero) pairs skipped.
  if (use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
      (RANK < 3 || (oz >= 0 && oz + TZ <= g.nf[2]))) {
    fence_proxy_async_smem();   // every thread: its generic-proxy tile writes -> visible to the TMA unit
    __syncthreads();
    if (tid == 0) {
      if (RANK == 2) {
        tma_reduce_add_3d(&tmap_out, tile4, 2 * ox, oy, t);
      } else {
        // 3D: the tensor-map box is ONE z-plane; only the planes the subproblem's stencils reach
        // (sub_desc.w, subproblem_zrange_kernel) are sent: the reduce-add traffic in L2 is what
        // bounds sparse point sets (a stack-of-stars bin holds one k_z: 7 of 10 planes).
        int tz_lo = 0, tz_hi = TZ;
        if (zrange) {
          const int lo = (sd.w & 0xffff) - 32768 - oz, hi = ((sd.w >> 16) & 0xffff) - 32768 - oz + NS;
          if (lo >= 0 && hi <= TZ && lo < hi) { tz_lo = lo; tz_hi = hi; }
        }
        for (int z = tz_lo; z < tz_hi; ++z)
          tma_reduce_add_4d(&tmap_out, tile4 + z * (TX * TY / 2), 2 * ox, oy, oz + z, t);
      }
      tma_store_commit_and_wait_read();
    }
    return;
  }
  const int TXH = TX / 2;
  for (int i = tid; i < ncell / 2; i += WPT * 32) {
    const float4 v = tile4[i];
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    const int ix = i % TXH;
    const int iy = (i / TXH) % TY;
    const int iz = i / (TXH * TY);
    const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
    const int gy = mod_idx(oy + iy, g.nf[1]);
    const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
    float2* dst = fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx;
    red_add(reinterpret_cast<float4*>(dst), v);
  }
}

template <int RANK, int WPT>
inline size_t spread_tile_smem_bytes(const int* bin) {
  const size_t ncell = static_cast<size_t>(bin[0] + 8) * (bin[1] + 8) * (RANK > 2 ? bin[2] + 8 : 1);
  return ncell * sizeof(float2) + static_cast<size_t>(WPT > 1 ? 2 : 1) * 32 * StageRec<RANK>::kWords * sizeof(float);
}

}  // namespace b200
