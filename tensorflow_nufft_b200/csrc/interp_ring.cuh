// interp_ring.cuh -- type-2 interpolator, persistent CTAs with a two-stage tile ring.
//
// Same sums as interp.cuh (reference: InterpSubproblem{2,3}DKernel nufft_plan.cu.cc:1041-1110,
// 1608-1706; c_j = sum_g fw[g] prod_d phi(g_d - x_{j,d}), periodic wrap), different schedule:
//   * grid = resident CTAs only (148 x occupancy); CTA b walks subproblems b, b + G, b + 2G, ...
//     and, inside a subproblem, the transforms of the batch;
//   * the (bin + halo) tile of item i + 1 is in flight (ONE TMA box copy, or 16-byte cp.async with
//     index wrap for tiles that straddle the periodic boundary) while the warps gather from the
//     tile of item i: the tile-load latency that dominated the one-tile-per-CTA kernel on sparse
//     point sets (stack-of-stars: tens of points per tile) is hidden;
//   * the stencil records of a subproblem (weights + stencil start + point id) are copied ONCE by
//     cp.async into shared memory and reused by every transform of the batch, prefetched one
//     subproblem ahead; nothing on the critical path is a dependent global load;
//   * points are dealt to the warps four at a time, so a tile with 40 points keeps all warps busy.
#pragma once
#include <cuda.h>
#include <cuda_pipeline.h>

#include "dev_common.cuh"
#include "interp.cuh"
#include "spread.cuh"

namespace b200 {

template <int RANK, int WARPS>
inline size_t interp_ring_smem_bytes(const int* bin, int msub) {
  const size_t ncell = static_cast<size_t>(bin[0] + 8) * (bin[1] + 8) * (RANK > 2 ? bin[2] + 8 : 1);
  const size_t tile_f4 = (ncell / 2 + 7) & ~static_cast<size_t>(7);
  const size_t rec_f4 = static_cast<size_t>(msub) * (2 * RANK + 1);
  return 2 * tile_f4 * sizeof(float4) + 2 * rec_f4 * sizeof(float4) + 2 * static_cast<size_t>(msub) * sizeof(int) + 32;
}

template <int NS, int RANK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
interp_ring_f32_kernel(int64_t M, GridGeom g, int ntr, int msub, const int* __restrict__ sub_total,
                       const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                       const int4* __restrict__ start, const float4* __restrict__ wrec4,
                       const float2* __restrict__ fw, float2* __restrict__ c,
                       const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int QX = (NS + 2) / 2;
  constexpr int C4 = 2 * RANK;       // float4 chunks of weights per point
  constexpr int F4 = C4 + 1;         // + the stencil start (int4)
  constexpr int NT = WARPS * 32;
  extern __shared__ __align__(128) float4 smem4[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int TX = g.bin[0] + 8, TY = g.bin[1] + 8;
  const int TZ = RANK > 2 ? g.bin[2] + 8 : 1;
  const int ncell = TX * TY * TZ;
  const int TXH = TX / 2;
  const int tile_f4 = (ncell / 2 + 7) & ~7;                          // 128-byte aligned tile pitch
  float4* tiles = smem4;                                             // [2][tile_f4]
  float4* recs = smem4 + 2 * tile_f4;                                // [2][msub][F4]
  int* idbuf = reinterpret_cast<int*>(recs + 2 * static_cast<size_t>(msub) * F4);   // [2][msub]
  uint64_t* bars = reinterpret_cast<uint64_t*>(idbuf + 2 * msub);    // [2]  (msub is a multiple of 4)

  const int nsub = *sub_total;
  const int G = gridDim.x;
  const int first = blockIdx.x;
  if (first >= nsub) return;
  const int nloc = (nsub - first + G - 1) / G;

  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  __syncthreads();

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < NS;
  const int lane_off = r * TX + 2 * q;
  const int zstride4 = TY * TX / 2;

  struct Sub { int p0, np, ox, oy, oz; bool interior; };
  auto decode = [&](const int4 sd) {
    Sub s;
    const int b = sd.x;
    s.p0 = sd.y;
    s.np = sd.z;
    const int bx = b % g.nbins[0];
    const int by = (b / g.nbins[0]) % g.nbins[1];
    const int bz = RANK > 2 ? b / (g.nbins[0] * g.nbins[1]) : 0;
    s.ox = bx * g.bin[0] - 4;
    s.oy = by * g.bin[1] - 4;
    s.oz = RANK > 2 ? bz * g.bin[2] - 4 : 0;
    s.interior = use_tma && s.ox >= 0 && s.ox + TX <= g.nf[0] && s.oy >= 0 && s.oy + TY <= g.nf[1] &&
                 (RANK < 3 || (s.oz >= 0 && s.oz + TZ <= g.nf[2]));
    return s;
  };
  auto load_desc = [&](int k) {
    return k < nloc ? sub_desc[first + k * G] : make_int4(0, 0, 0, 0);
  };
  // Starts the copy of one tile into ring slot `buf` (does not wait).
  auto issue_tile = [&](const Sub& s, int t, int buf) {
    float4* dst = tiles + buf * tile_f4;
    if (s.interior) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this slot are done
        mbar_expect_tx(&bars[buf], static_cast<uint32_t>(ncell * sizeof(float2)));
        if (RANK == 2) tma_load_3d(dst, &tmap, &bars[buf], 2 * s.ox, s.oy, t);
        else tma_load_4d(dst, &tmap, &bars[buf], 2 * s.ox, s.oy, s.oz, t);
      }
    } else {
      const float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;
      for (int i = tid; i < ncell / 2; i += NT) {
        const int ix = i % TXH;
        const int iy = (i / TXH) % TY;
        const int iz = i / (TXH * TY);
        const int gx = mod_idx(s.ox + 2 * ix, g.nf[0]);
        const int gy = mod_idx(s.oy + iy, g.nf[1]);
        const int gz = RANK > 2 ? mod_idx(s.oz + iz, g.nf[2]) : 0;
        __pipeline_memcpy_async(&dst[i], fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx, 16);
      }
    }
  };
  // Starts the copy of a subproblem's records into record slot `rbuf`.
  auto issue_records = [&](const Sub& s, int rbuf) {
    float4* dst = recs + static_cast<size_t>(rbuf) * msub * F4;
    const int total = s.np * F4;
    for (int e = tid; e < total; e += NT) {
      const int j = e / F4;
      const int k = e - j * F4;
      const float4* src = k < C4 ? wrec4 + (static_cast<int64_t>(s.p0) + j) * C4 + k
                                 : reinterpret_cast<const float4*>(start + s.p0 + j);
      __pipeline_memcpy_async(&dst[e], src, 16);
    }
    int* idst = idbuf + rbuf * msub;
    for (int j = tid; j < s.np; j += NT) __pipeline_memcpy_async(&idst[j], idx + s.p0 + j, 4);
  };

  // Descriptor queue: the descriptor of subproblem k + 1 is needed when the last transform of
  // subproblem k starts; it is fetched two subproblems ahead so that it never stalls.
  Sub cur = decode(load_desc(0));
  int4 d1 = load_desc(1);
  int4 d2 = load_desc(2);
  issue_records(cur, 0);
  issue_tile(cur, 0, 0);
  __pipeline_commit();

  uint32_t phases = 0u;   // bit b = parity of ring slot b's mbarrier
  int buf = 0, rbuf = 0;
  for (int k = 0; k < nloc; ++k) {
    const Sub nxt = decode(d1);
    const bool has_next_sub = k + 1 < nloc;
    const float4* recs4 = recs + static_cast<size_t>(rbuf) * msub * F4;
    const int* ids = idbuf + rbuf * msub;
    const int np = cur.np;
    const int ngrp = (np + 3) >> 2;
    for (int t = 0; t < ntr; ++t) {
      // ---- start the loads of the next item ----
      if (t + 1 < ntr) {
        issue_tile(cur, t + 1, buf ^ 1);
      } else if (has_next_sub) {
        issue_records(nxt, rbuf ^ 1);
        issue_tile(nxt, 0, buf ^ 1);
      }
      __pipeline_commit();   // one group per item, possibly empty
      // ---- wait for the current item ----
      __pipeline_wait_prior(1);
      if (cur.interior) {
        mbar_wait(&bars[buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
      }
      __syncthreads();

      const float2* tile = reinterpret_cast<const float2*>(tiles + buf * tile_f4);
      float2* ct = c + static_cast<int64_t>(t) * M;
      for (int g0 = warp; g0 < ngrp; g0 += 8 * WARPS) {
        float2 res_l = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int k8 = 0; k8 < 8; ++k8) {
          const int gi = g0 + k8 * WARPS;
          if (gi >= ngrp) break;
          const int pb = gi * 4;
          float re[4], im[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            re[u] = 0.f;
            im[u] = 0.f;
            const int p = pb + u;
            if (p < np) {
              const float* rec = reinterpret_cast<const float*>(recs4 + p * F4);
              const int rx = __float_as_int(rec[4 * C4]) - cur.ox;
              const int ry = __float_as_int(rec[4 * C4 + 1]) - cur.oy;
              const int rz = RANK > 2 ? __float_as_int(rec[4 * C4 + 2]) - cur.oz : 0;
              // Memory safety for coordinates outside the declared points_range (see interp.cuh).
              const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                                (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
              if (row_ok && fits) {
                const int off = (rz * TY + ry) * TX + rx;
                const float2 wx = *reinterpret_cast<const float2*>(rec + 2 * q);
                const float wy = rec[8 + r];
                const float4* ptr = reinterpret_cast<const float4*>(tile + off + lane_off);
                if (RANK == 2) {
                  const float4 v = *ptr;
                  re[u] = wy * (v.x * wx.x + v.z * wx.y);
                  im[u] = wy * (v.y * wx.x + v.w * wx.y);
                } else {
                  float4 v[NS];
#pragma unroll
                  for (int dz = 0; dz < NS; ++dz) v[dz] = ptr[dz * zstride4];
                  float wz[8];
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    const float2 t2 = *reinterpret_cast<const float2*>(rec + 16 + 2 * kk);
                    wz[2 * kk] = t2.x;
                    wz[2 * kk + 1] = t2.y;
                  }
                  float ar = 0.f, ai = 0.f;
#pragma unroll
                  for (int dz = 0; dz < NS; ++dz) {
                    ar += wz[dz] * (v[dz].x * wx.x + v[dz].z * wx.y);
                    ai += wz[dz] * (v[dz].y * wx.x + v[dz].w * wx.y);
                  }
                  re[u] = ar * wy;
                  im[u] = ai * wy;
                }
              }
            }
          }
          // Transposing butterfly (see interp_tile_f32_kernel): 18 shuffles per 4 points.
          {
            const int p4 = 4 * k8;
            const bool hi16 = lane & 16, hi8 = lane & 8;
            float a0 = hi16 ? re[0] : re[1], k0 = hi16 ? re[1] : re[0];
            float a1 = hi16 ? re[2] : re[3], k1 = hi16 ? re[3] : re[2];
            k0 += __shfl_xor_sync(0xffffffffu, a0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, a1, 16);
            float a2 = hi8 ? k0 : k1, kr = hi8 ? k1 : k0;
            kr += __shfl_xor_sync(0xffffffffu, a2, 8);
            float b0 = hi16 ? im[0] : im[1], m0 = hi16 ? im[1] : im[0];
            float b1 = hi16 ? im[2] : im[3], m1 = hi16 ? im[3] : im[2];
            m0 += __shfl_xor_sync(0xffffffffu, b0, 16);
            m1 += __shfl_xor_sync(0xffffffffu, b1, 16);
            float b2 = hi8 ? m0 : m1, ki = hi8 ? m1 : m0;
            ki += __shfl_xor_sync(0xffffffffu, b2, 8);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
              kr += __shfl_xor_sync(0xffffffffu, kr, o);
              ki += __shfl_xor_sync(0xffffffffu, ki, o);
            }
            const int src_lane = ((lane - p4) & 1 ? 16 : 0) | ((lane - p4) & 2 ? 8 : 0);
            const float rr = __shfl_sync(0xffffffffu, kr, src_lane);
            const float ri = __shfl_sync(0xffffffffu, ki, src_lane);
            if (lane >= p4 && lane < p4 + 4) res_l = make_float2(rr, ri);
          }
        }
        // lane l holds point 4 * (g0 + (l / 4) * WARPS) + l % 4
        const int p = 4 * (g0 + (lane >> 2) * WARPS) + (lane & 3);
        if (p < np) ct[ids[p]] = res_l;
      }
      __syncthreads();   // everyone is done with ring slot `buf` (and, after the last transform, `rbuf`)
      buf ^= 1;
    }
    cur = nxt;
    d1 = d2;
    d2 = load_desc(k + 3);
    rbuf ^= 1;
  }
}

}  // namespace b200
