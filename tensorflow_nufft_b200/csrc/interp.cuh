// interp.cuh -- type-2 interpolator: c_j = sum_g fw[g] prod_d phi(g_d - x_{j,d}), periodic wrap.
// Behavioural reference: InterpNuptsDriven{2,3}DKernel nufft_plan.cu.cc:963-1039,1513-1606,
// InterpSubproblem{2,3}DKernel :1041-1110,1608-1706, CPU interp_line/square/cube
// nufft_plan.cc:1309-1461.
//
//  * interp_global_kernel    one thread per point, gathers ns^d cells through L1/L2. Any rank /
//                            width / precision. Fallback + cross-check.
//  * interp_tile_f32_kernel  one CTA per subproblem: the (bin + halo) tile of the fine grid is
//                            staged in shared memory, then each warp takes points of the
//                            subproblem, lanes laid over the stencil (row = lane / QX, cell pair =
//                            lane % QX): 128-bit conflict-free shared loads, butterfly reduction.
#pragma once
#include "dev_common.cuh"
#include "spread.cuh"

namespace b200 {

template <typename F>
__global__ void __launch_bounds__(128)
interp_global_kernel(int64_t M, int ntr, GridGeom g, int ns, int R, int PX, int PY, const int* __restrict__ idx,
                     const int4* __restrict__ start, const F* __restrict__ wrec,
                     const Cplx<F>* __restrict__ fw, Cplx<F>* __restrict__ c) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < M; j += stride) {
    const int4 st = start[j];
    const int pid = idx[j];
    const F* kx = wrec + j * R + st.w;
    const F* wy = wrec + j * R + PX;
    const F* wz = wy + PY;
    const int nz = g.rank > 2 ? ns : 1, ny = g.rank > 1 ? ns : 1;
    for (int t = 0; t < ntr; ++t) {
      const Cplx<F>* in = fw + static_cast<int64_t>(t) * g.nftot;
      F re = F(0), im = F(0);
      for (int dz = 0; dz < nz; ++dz) {
        const F kz = g.rank > 2 ? wz[dz] : F(1);
        const int64_t oz = g.rank > 2 ? static_cast<int64_t>(mod_idx(st.z + dz, g.nf[2])) * g.nf[0] * g.nf[1] : 0;
        for (int dy = 0; dy < ny; ++dy) {
          const F kyz = g.rank > 1 ? wy[dy] * kz : kz;
          const int64_t oy = oz + (g.rank > 1 ? static_cast<int64_t>(mod_idx(st.y + dy, g.nf[1])) * g.nf[0] : 0);
          for (int k = 0; k < ns; ++k) {
            const int gx = mod_idx(st.x + st.w + k, g.nf[0]);
            const Cplx<F> v = in[oy + gx];
            const F w = kx[k] * kyz;
            re += v.x * w;
            im += v.y * w;
          }
        }
      }
      c[static_cast<int64_t>(t) * M + pid] = make_cplx<F>(re, im);
    }
  }
}

template <int NS, int RANK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
interp_tile_f32_kernel(int64_t M, GridGeom g, int msub, const int* __restrict__ sub_total,
                       const int* __restrict__ sub_start, const int* __restrict__ bin_start,
                       const int* __restrict__ bin_sizes, const int* __restrict__ idx,
                       const int4* __restrict__ start, const float* __restrict__ wrec,
                       const float2* __restrict__ fw, float2* __restrict__ c) {
  constexpr int QX = (NS + 2) / 2;
  constexpr int RPI = NS;
  constexpr int R = 8 * RANK;
  extern __shared__ float4 tile4[];
  float2* tile = reinterpret_cast<float2*>(tile4);

  const int s = blockIdx.x;
  if (s >= *sub_total) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
  const int TXH = TX / 2;

  const float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;
  float2* ct = c + static_cast<int64_t>(t) * M;

  // Stage the tile (two cells per 128-bit load; nf and the tile origin are even so a pair never
  // straddles the periodic boundary).
  for (int i = threadIdx.x; i < ncell / 2; i += WARPS * 32) {
    const int ix = i % TXH;
    const int iy = (i / TXH) % TY;
    const int iz = i / (TXH * TY);
    const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
    const int gy = mod_idx(oy + iy, g.nf[1]);
    const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
    tile4[i] = *reinterpret_cast<const float4*>(fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx);
  }
  __syncthreads();

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < RPI;
  const int lane_off = r * TX + 2 * q;

  for (int base = warp * 32; base < np; base += WARPS * 32) {
    const int jl = p0 + base + lane;
    int off_l = -1;
    int pid_l = 0;
    if (base + lane < np) {
      pid_l = idx[jl];
      const int4 st = start[jl];
      const int rx = st.x - ox, ry = st.y - oy, rz = RANK > 2 ? st.z - oz : 0;
      const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                        (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
      off_l = fits ? (rz * TY + ry) * TX + rx : -1;
    }
    float2 res_l = make_float2(0.f, 0.f);
    const int cnt = min(32, np - base);
    for (int p = 0; p < cnt; ++p) {
      const int64_t j = p0 + base + p;
      const float wv = lane < R ? wrec[j * R + lane] : 0.f;
      const int boff = __shfl_sync(0xffffffffu, off_l, p);
      const float wxa = __shfl_sync(0xffffffffu, wv, 2 * q);
      const float wxb = __shfl_sync(0xffffffffu, wv, 2 * q + 1);
      const float wyr = __shfl_sync(0xffffffffu, wv, 8 + (row_ok ? r : 0));
      float re = 0.f, im = 0.f;
      if (RANK == 2) {
        if (row_ok && boff >= 0) {
          const float4 v = *reinterpret_cast<const float4*>(tile + boff + lane_off);
          re = wyr * (v.x * wxa + v.z * wxb);
          im = wyr * (v.y * wxa + v.w * wxb);
        }
      } else {
        float wz[NS];
#pragma unroll
        for (int dz = 0; dz < NS; ++dz) wz[dz] = __shfl_sync(0xffffffffu, wv, 16 + dz);
        if (row_ok && boff >= 0) {
          const float4* ptr = reinterpret_cast<const float4*>(tile + boff + lane_off);
          const int zstride = TY * TX / 2;
          float4 v[NS];
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) v[dz] = ptr[dz * zstride];
#pragma unroll
          for (int dz = 0; dz < NS; ++dz) {
            re += wz[dz] * (v[dz].x * wxa + v[dz].z * wxb);
            im += wz[dz] * (v[dz].y * wxa + v[dz].w * wxb);
          }
          re *= wyr;
          im *= wyr;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
      }
      if (lane == p) res_l = make_float2(re, im);
    }
    if (base + lane < np) ct[pid_l] = res_l;
  }
}

}  // namespace b200
