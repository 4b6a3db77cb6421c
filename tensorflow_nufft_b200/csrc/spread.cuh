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
// Shared-memory tile spreader (complex64, ns <= 7, rank 2 or 3). One warp per CTA.
// Work item = (subproblem, transform). Tile = (bin + 8)^rank cells of float2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_bin_of_subproblem(const int* __restrict__ sub_start, int nb, int s) {
  // largest b with sub_start[b] <= s  (sub_start is an exclusive scan, non-decreasing)
  int lo = 0, hi = nb - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (sub_start[mid] <= s) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int NS, int RANK>
__global__ void __launch_bounds__(32)
spread_tile_f32_kernel(int64_t M, GridGeom g, int msub, const int* __restrict__ sub_total,
                       const int* __restrict__ sub_start, const int* __restrict__ bin_start,
                       const int* __restrict__ bin_sizes, const int* __restrict__ idx,
                       const int4* __restrict__ start, const float* __restrict__ wrec /*[M][8*RANK]*/,
                       const float2* __restrict__ c, float2* __restrict__ fw) {
  constexpr int QX = (NS + 2) / 2;      // float4 lanes per stencil row: covers NS+1 cells
  constexpr int RPI = NS;               // rows per instruction (lanes r = lane / QX < NS active)
  static_assert(QX * RPI <= 32, "stencil slab must fit one warp");
  constexpr int R = 8 * RANK;           // floats per weight record
  extern __shared__ float4 tile4[];     // [TZ][TY][TX/2] pairs of cells
  float2* tile = reinterpret_cast<float2*>(tile4);

  const int s = blockIdx.x;
  if (s >= *sub_total) return;
  const int lane = threadIdx.x;
  const int t = blockIdx.y;
  const int nbtot = g.nbins[0] * g.nbins[1] * g.nbins[2];
  const int b = find_bin_of_subproblem(sub_start, nbtot, s);
  const int within = s - sub_start[b];
  const int p0 = bin_start[b] + within * msub;
  const int np = min(msub, bin_sizes[b] - within * msub);

  const int TX = g.bin[0] + 8, TY = g.bin[1] + 8;
  const int TZ = RANK > 2 ? g.bin[2] + 8 : 1;
  const int bx = b % g.nbins[0];
  const int by = (b / g.nbins[0]) % g.nbins[1];
  const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
  const int ox = bx * g.bin[0] - 4, oy = by * g.bin[1] - 4, oz = RANK > 2 ? bz * g.bin[2] - 4 : 0;
  const int ncell = TX * TY * TZ;

  for (int i = lane; i < ncell / 2; i += 32) tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < RPI;
  const int lane_off = r * TX + 2 * q;   // cells, within a z-plane, relative to the point's base

  const float2* ct = c + static_cast<int64_t>(t) * M;
  float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;

  for (int base = 0; base < np; base += 32) {
    // Each lane prefetches one point's header: strength (gather through idx) and tile offset.
    const int jl = p0 + base + lane;
    float2 c_l = make_float2(0.f, 0.f);
    int off_l = 0;
    if (base + lane < np) {
      const int4 st = start[jl];
      const int rx = st.x - ox, ry = st.y - oy, rz = RANK > 2 ? st.z - oz : 0;
      // Memory safety for coordinates outside the declared points_range: such a stencil does not
      // lie in this bin's tile and the point is dropped (the reference's behaviour is undefined).
      const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                        (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
      if (fits) {
        c_l = ct[idx[jl]];
        off_l = (rz * TY + ry) * TX + rx;
      }
    }
    const int cnt = min(32, np - base);
    for (int p = 0; p < cnt; ++p) {
      const int64_t j = p0 + base + p;
      const float wv = lane < R ? wrec[j * R + lane] : 0.f;
      const float cre = __shfl_sync(0xffffffffu, c_l.x, p);
      const float cim = __shfl_sync(0xffffffffu, c_l.y, p);
      const int boff = __shfl_sync(0xffffffffu, off_l, p);
      const float wxa = __shfl_sync(0xffffffffu, wv, 2 * q);
      const float wxb = __shfl_sync(0xffffffffu, wv, 2 * q + 1);
      const float wyr = __shfl_sync(0xffffffffu, wv, 8 + (row_ok ? r : 0));
      const float4 cx = make_float4(cre * wxa, cim * wxa, cre * wxb, cim * wxb);
      if (RANK == 2) {
        if (row_ok) {
          float4* ptr = reinterpret_cast<float4*>(tile + boff + lane_off);
          float4 v = *ptr;
          v.x += wyr * cx.x; v.y += wyr * cx.y; v.z += wyr * cx.z; v.w += wyr * cx.w;
          *ptr = v;
        }
      } else {
        float4 v[NS];
        float wz[NS];
#pragma unroll
        for (int dz = 0; dz < NS; ++dz) wz[dz] = __shfl_sync(0xffffffffu, wv, 16 + dz);
        if (row_ok) {
          float4* ptr = reinterpret_cast<float4*>(tile + boff + lane_off);
          const int zstride = TY * TX / 2;
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) v[dz] = ptr[dz * zstride];
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) {
            const float w = wyr * wz[dz];
            v[dz].x += w * cx.x; v[dz].y += w * cx.y; v[dz].z += w * cx.z; v[dz].w += w * cx.w;
          }
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) ptr[dz * zstride] = v[dz];
        }
      }
      __syncwarp();
    }
  }

  // Flush: two complex cells per REDG.ADD.F32x4; periodic wrap; untouched (zero) pairs skipped.
  const int TXH = TX / 2;
  for (int i = lane; i < ncell / 2; i += 32) {
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

}  // namespace b200
