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
#include <cuda.h>
#include <cuda_pipeline.h>

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

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers, raw PTX for sm_100a ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// One box of the fine grid -> shared memory. Coordinates are in elements of the tensor map
// (float32; innermost dim = 2 * nf0 interleaved re/im), fastest dimension first.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Tile -> fine grid with an element-wise ADD done by the TMA unit (cp.reduce.async.bulk.tensor):
// one instruction flushes a whole spreader tile; elements outside the tensor are skipped.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One CTA (WARPS warps) per (subproblem, transform). The (bin + halo) tile is copied into shared
// memory by ONE TMA box copy (cp.async.bulk.tensor, completion on an mbarrier) when the tile lies
// inside the grid, or with 16-byte cp.async (LDGSTS) copies with index wrapping when it straddles
// the periodic boundary (TMA zero-fills out-of-bounds, it does not wrap). Then each warp takes
// batches of 32 points: lane l
// prefetches the record of its point of the NEXT batch (weights, start, point id) while the warp
// gathers the current batch from the tile; records are staged per warp in shared memory
// ({tile offset, wx[8], wy[8], wz[8]}), results are reduced by butterfly shuffles and scattered to
// c[idx] once per batch.
template <int NS, int RANK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
interp_tile_f32_kernel(int64_t M, GridGeom g, const int* __restrict__ sub_total,
                       const int4* __restrict__ sub_desc, const int* __restrict__ idx,
                       const int4* __restrict__ start, const float4* __restrict__ wrec4,
                       const float2* __restrict__ fw, float2* __restrict__ c,
                       const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int QX = (NS + 2) / 2;
  constexpr int C4 = 2 * RANK;
  constexpr int SW = StageRec<RANK>::kWords;
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
  const int TXH = TX / 2;
  float4* tile4 = smem4;
  const float2* tile = reinterpret_cast<const float2*>(tile4);
  float* stage = reinterpret_cast<float*>(smem4 + ncell / 2) + warp * 32 * SW;   // per warp [32][SW]
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem4 + ncell / 2) + WARPS * 32 * SW);

  const float2* fwt = fw + static_cast<int64_t>(t) * g.nftot;
  float2* ct = c + static_cast<int64_t>(t) * M;

  // ---- prefetch the first batch's records while the tile streams in ----
  float4 w4[C4];
  int4 st_n = make_int4(0, 0, 0, 0);
  int id_n = 0;
  auto fetch = [&](int first) {
    const int pl = first + lane;
    if (pl < np) {
      const int64_t j = p0 + pl;
#pragma unroll
      for (int k = 0; k < C4; ++k) w4[k] = wrec4[j * C4 + k];
      st_n = start[j];
      id_n = idx[j];
    }
  };
  fetch(warp * 32);

  const bool interior = use_tma && ox >= 0 && ox + TX <= g.nf[0] && oy >= 0 && oy + TY <= g.nf[1] &&
                        (RANK < 3 || (oz >= 0 && oz + TZ <= g.nf[2]));
  if (interior) {
    // TMA: one elected thread arms the mbarrier with the tile's byte count and issues the box copy.
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, static_cast<uint32_t>(ncell * sizeof(float2)));
      if (RANK == 2) tma_load_3d(tile4, &tmap, bar, 2 * ox, oy, t);
      else tma_load_4d(tile4, &tmap, bar, 2 * ox, oy, oz, t);
    }
    mbar_wait(bar, 0);
  } else {
    // Stage the tile (two cells per 16-byte async copy; nf and the tile origin are even so a pair
    // never straddles the periodic boundary).
    for (int i = tid; i < ncell / 2; i += WARPS * 32) {
      const int ix = i % TXH;
      const int iy = (i / TXH) % TY;
      const int iz = i / (TXH * TY);
      const int gx = mod_idx(ox + 2 * ix, g.nf[0]);
      const int gy = mod_idx(oy + iy, g.nf[1]);
      const int gz = RANK > 2 ? mod_idx(oz + iz, g.nf[2]) : 0;
      const float2* src = fwt + (static_cast<int64_t>(gz) * g.nf[1] + gy) * g.nf[0] + gx;
      __pipeline_memcpy_async(&tile4[i], src, 16);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
  }

  const int q = lane % QX;
  const int r = lane / QX;
  const bool row_ok = r < NS;
  const int lane_off = r * TX + 2 * q;
  const int zstride4 = TY * TX / 2;

  for (int first = warp * 32; first < np; first += WARPS * 32) {
    // ---- stage this lane's point, then prefetch the next batch ----
    const int id_cur = id_n;
    {
      const int pl = first + lane;
      float4* rec4 = reinterpret_cast<float4*>(stage + lane * SW);
      int off = -1;
      if (pl < np) {
        const int rx = st_n.x - ox, ry = st_n.y - oy, rz = RANK > 2 ? st_n.z - oz : 0;
        const bool fits = rx >= 0 && rx + 2 * QX <= TX && ry >= 0 && ry + NS <= TY &&
                          (RANK < 3 || (rz >= 0 && rz + NS <= TZ));
        if (fits) off = (rz * TY + ry) * TX + rx;
      }
#pragma unroll
      for (int k = 0; k < C4; ++k) rec4[k] = w4[k];
      rec4[6] = make_float4(__int_as_float(off), 0.f, 0.f, 0.f);
    }
    __syncwarp();
    fetch(first + WARPS * 32);

    float2 res_l = make_float2(0.f, 0.f);
    const int cnt = min(32, np - first);
    // Points are processed four at a time so that the butterfly reduction can be shared: the four
    // (re, im) partial sums are folded with a transposing reduction (9 + 9 shuffles per 4 points
    // instead of 40), and the four gathers are independent (ILP across points).
    for (int p4 = 0; p4 < cnt; p4 += 4) {
      float re[4], im[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        re[u] = 0.f;
        im[u] = 0.f;
        const int p = p4 + u;
        if (p < cnt) {
          const float* rec = stage + p * SW;
          const int off = __float_as_int(rec[24]);
          if (row_ok && off >= 0) {
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
              for (int k = 0; k < 4; ++k) {
                const float2 t2 = *reinterpret_cast<const float2*>(rec + 16 + 2 * k);
                wz[2 * k] = t2.x;
                wz[2 * k + 1] = t2.y;
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
      // Transposing butterfly: after the xor-16 and xor-8 steps each lane holds ONE of the four
      // sums (selected by lane bits 4 and 3); three more steps finish it.
      {
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
        // lane bits (4,3) = (hi16, hi8) select the point: kept value index = 2*hi8 + hi16
        const int which = (hi8 ? 2 : 0) + (hi16 ? 1 : 0);
        // lanes 0, 16, 8, 24 hold points p4+0, p4+1, p4+2, p4+3; hand each to lane (p4 + which)
        const int src_lane = ((lane - p4) & 1 ? 16 : 0) | ((lane - p4) & 2 ? 8 : 0);
        const float rr = __shfl_sync(0xffffffffu, kr, src_lane);
        const float ri = __shfl_sync(0xffffffffu, ki, src_lane);
        (void)which;
        if (lane >= p4 && lane < p4 + 4) res_l = make_float2(rr, ri);
      }
    }
    if (first + lane < np) ct[id_cur] = res_l;
    __syncwarp();
  }
}

template <int RANK, int WARPS>
inline size_t interp_tile_smem_bytes(const int* bin) {
  const size_t ncell = static_cast<size_t>(bin[0] + 8) * (bin[1] + 8) * (RANK > 2 ? bin[2] + 8 : 1);
  return ncell * sizeof(float2) + static_cast<size_t>(WARPS) * 32 * StageRec<RANK>::kWords * sizeof(float) + 16;
}

}  // namespace b200
