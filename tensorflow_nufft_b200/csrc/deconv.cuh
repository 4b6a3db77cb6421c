// deconv.cuh -- fused deconvolve+crop (type-1 step 3) and amplify+zero-pad (type-2 step 1).
// Behavioural reference: Deconvolve{1,2,3}DKernel / Amplify{1,2,3}DKernel nufft_plan.cu.cc:326-435,
// CPU deconvolve_{1,2,3}d nufft_plan.cc:729-881 (CMCL mode order: index i <-> k = i - N/2; fine
// index w = k >= 0 ? k : nf + k; factor prod_d phihat_d[|k_d|]; no 1/N anywhere).
// The arithmetic follows the CPU plan's association ((1/p3)/p2 * v) / p1 so that the parity
// oracle's roundings are reproduced.
#pragma once
#include "dev_common.cuh"

namespace b200 {

struct ModeGeom {
  int rank;
  int n[3];    // modes per dim
  int nf[3];   // fine size per dim
  int64_t ntot, nftot;
};

// Row-based mapping: one CTA per (row of the fastest axis, transform). The slow-axis indices and the
// slow-axis prefactor (1/p3)/p2 are computed once per row, no per-element integer division; x is
// walked with coalesced accesses.
// grid: (n2*n3, ntr). fk[t][i3][i2][i1] = fw[t][wrap(k)] / factor
template <typename F>
__global__ void __launch_bounds__(256)
deconvolve_kernel(ModeGeom m, const F* __restrict__ p1, const F* __restrict__ p2, const F* __restrict__ p3,
                  const Cplx<F>* __restrict__ fw, Cplx<F>* __restrict__ fk) {
  const int row = blockIdx.x;
  const int t = blockIdx.y;
  const int i2 = m.rank > 1 ? row % m.n[1] : 0;
  const int i3 = m.rank > 2 ? row / m.n[1] : 0;
  F pre = F(1);
  int64_t in_row = 0;
  if (m.rank > 2) {
    const int k3 = i3 - m.n[2] / 2;
    const int w3 = k3 >= 0 ? k3 : m.nf[2] + k3;
    pre = pre / p3[abs(k3)];
    in_row += static_cast<int64_t>(w3) * m.nf[0] * m.nf[1];
  }
  if (m.rank > 1) {
    const int k2 = i2 - m.n[1] / 2;
    const int w2 = k2 >= 0 ? k2 : m.nf[1] + k2;
    pre = pre / p2[abs(k2)];
    in_row += static_cast<int64_t>(w2) * m.nf[0];
  }
  const Cplx<F>* src = fw + static_cast<int64_t>(t) * m.nftot + in_row;
  Cplx<F>* dst = fk + static_cast<int64_t>(t) * m.ntot + static_cast<int64_t>(row) * m.n[0];
  const int half = m.n[0] / 2;
  for (int i1 = threadIdx.x; i1 < m.n[0]; i1 += blockDim.x) {
    const int k1 = i1 - half;
    const int w1 = k1 >= 0 ? k1 : m.nf[0] + k1;
    const F f1 = p1[abs(k1)];
    const Cplx<F> v = src[w1];
    dst[i1] = make_cplx<F>((pre * v.x) / f1, (pre * v.y) / f1);
  }
}

// Two adjacent complex cells as one store: 16 bytes for complex64 (one STG.128), 2 x 16 for complex128.
__device__ __forceinline__ void store_pair(float2* dst, float2 a, float2 b) {
  *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_pair(double2* dst, double2 a, double2 b) {
  dst[0] = a;
  dst[1] = b;
}

constexpr int kAmplifyRowsPerCta = 8;   // one warp per fine-grid row

// grid: (ceil(nf2*nf3 / 8), ntr), 256 threads: ONE WARP per row of the fastest axis. Writes EVERY
// fine cell of the row: amplified mode or zero (no memset pass). 7/8 of a 3D fine grid is zero
// padding, so this is mostly a fill: the row bookkeeping (which slow-axis modes exist, the
// slow-axis prefactor) is done once per warp and every lane then issues 16-byte stores of two
// cells (a CTA-per-row version with 8-byte stores spent 77 instructions per thread on 2 cells and
// was issue-bound at 3.4 TB/s). nf[0] is even by construction (fine sizes are even).
template <typename F>
__global__ void __launch_bounds__(256)
amplify_kernel(ModeGeom m, const F* __restrict__ p1, const F* __restrict__ p2, const F* __restrict__ p3,
               const Cplx<F>* __restrict__ fk, Cplx<F>* __restrict__ fw) {
  // rank 1 has a single row: the whole CTA walks it
  const int lane = m.rank == 1 ? threadIdx.x : (threadIdx.x & 31);
  const int step = m.rank == 1 ? 2 * blockDim.x : 64;
  const int64_t nrows = m.nftot / m.nf[0];
  const int64_t row = m.rank == 1 ? 0 : static_cast<int64_t>(blockIdx.x) * kAmplifyRowsPerCta + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const int t = blockIdx.y;
  const int w2 = m.rank > 1 ? static_cast<int>(row % m.nf[1]) : 0;
  const int w3 = m.rank > 2 ? static_cast<int>(row / m.nf[1]) : 0;
  // mode k_d present iff w_d <= kmax_d (k = w) or w_d >= nf_d + kmin_d (k = w - nf)
  bool ok = true;
  F pre = F(1);
  int64_t src_row = 0;
  if (m.rank > 2) {
    const int kmax = (m.n[2] - 1) / 2, kmin = -(m.n[2] / 2);
    const int k3 = w3 <= kmax ? w3 : w3 - m.nf[2];
    ok = ok && (w3 <= kmax || w3 >= m.nf[2] + kmin);
    if (ok) {
      pre = pre / p3[abs(k3)];
      src_row += static_cast<int64_t>(k3 + m.n[2] / 2) * m.n[0] * m.n[1];
    }
  }
  if (m.rank > 1) {
    const int kmax = (m.n[1] - 1) / 2, kmin = -(m.n[1] / 2);
    const int k2 = w2 <= kmax ? w2 : w2 - m.nf[1];
    const bool ok2 = (w2 <= kmax || w2 >= m.nf[1] + kmin);
    if (ok && ok2) {
      pre = pre / p2[abs(k2)];
      src_row += static_cast<int64_t>(k2 + m.n[1] / 2) * m.n[0];
    }
    ok = ok && ok2;
  }
  Cplx<F>* dst = fw + static_cast<int64_t>(t) * m.nftot + row * m.nf[0];
  const Cplx<F> zero = make_cplx<F>(F(0), F(0));
  if (!ok) {
    for (int w1 = 2 * lane; w1 < m.nf[0]; w1 += step) store_pair(dst + w1, zero, zero);
    return;
  }
  const Cplx<F>* src = fk + static_cast<int64_t>(t) * m.ntot + src_row;
  const int kmax = (m.n[0] - 1) / 2, kmin = -(m.n[0] / 2);
  const int half = m.n[0] / 2;
  auto cell = [&](int w1) {
    Cplx<F> out = zero;
    if (w1 <= kmax || w1 >= m.nf[0] + kmin) {
      const int k1 = w1 <= kmax ? w1 : w1 - m.nf[0];
      const F f1 = p1[abs(k1)];
      const Cplx<F> v = src[k1 + half];
      out = make_cplx<F>((pre * v.x) / f1, (pre * v.y) / f1);
    }
    return out;
  };
  for (int w1 = 2 * lane; w1 < m.nf[0]; w1 += step) store_pair(dst + w1, cell(w1), cell(w1 + 1));
}

template <typename F>
__global__ void __launch_bounds__(256)
scale_kernel(int64_t n, F s, Cplx<F>* __restrict__ a) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    a[i].x *= s;
    a[i].y *= s;
  }
}

}  // namespace b200
