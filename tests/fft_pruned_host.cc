// CPU harness for tensorflow_nufft_b200/csrc/fft_pruned.cuh (TEST INFRASTRUCTURE): runs the kernels'
// own load / butterfly / store functions thread by thread, block by block, so that the index
// arithmetic of the pruned FFT passes is checked against numpy.fft without a GPU.
#include <cmath>
#include <cstring>
#include <vector>

#include "../tensorflow_nufft_b200/csrc/fft_pruned.cuh"

using namespace b200;

namespace {

struct HostExec {
  int type, rank, sign, ntr;
  const int* n;
  float2* fw;
  float2* f;
  const float* fac[3];
  int bad = 0;

  std::vector<float2> twiddles(int logn) const {
    std::vector<float2> tw(fft_tw_count(logn) + 1);
    fft_fill_twiddles(logn, sign, tw.data());
    return tw;
  }

  template <int LOGN, int KIND>
  void col_t(const FftColGeom& g, const float* pa, const float* po) {
    using A = FftAlg<LOGN>;
    constexpr int LOGW = fft_logw(LOGN);
    std::vector<float2> s(fft_col_smem_bytes(LOGN) / sizeof(float2));
    const std::vector<float2> tw = twiddles(LOGN);
    const float sg = static_cast<float>(sign);
    for (int t = 0; t < ntr; ++t)
      for (int o = 0; o < g.outer_count; ++o)
        for (int gx = 0; gx < (g.N0 >> LOGW); ++gx) {
          const FftColCtx cx = fft_col_ctx(g, LOGW, gx, o, t, fw, f);
          for (int tid = 0; tid < kFftColThreads; ++tid)
            fft_col_first<LOGN, KIND>(g, cx, o, tid, kFftColThreads, pa, po, fac[0], s.data(), sg);
          if (A::NM >= 1)
            for (int tid = 0; tid < kFftColThreads; ++tid) fft_col_middle<LOGN, 1>(tid, kFftColThreads, s.data(), tw.data(), sg);
          if (A::NM >= 2)
            for (int tid = 0; tid < kFftColThreads; ++tid) fft_col_middle<LOGN, (A::NM >= 2 ? 2 : 1)>(tid, kFftColThreads, s.data(), tw.data(), sg);
          for (int tid = 0; tid < kFftColThreads; ++tid)
            fft_col_last<LOGN, KIND>(g, cx, o, tid, kFftColThreads, pa, po, fac[0], s.data(), tw.data(), sg);
        }
  }
  template <int LOGN>
  void col_k(int kind, const FftColGeom& g, const float* pa, const float* po) {
    if (kind == kFftPlain) col_t<LOGN, kFftPlain>(g, pa, po);
    else if (kind == kFftFromModes) col_t<LOGN, kFftFromModes>(g, pa, po);
    else col_t<LOGN, kFftToModes>(g, pa, po);
  }
  void col(int axis, int kind, const FftColGeom& g, int axis_a, int axis_o) {
    const float* pa = axis_a >= 0 ? fac[axis_a] : nullptr;
    const float* po = axis_o >= 0 ? fac[axis_o] : nullptr;
    switch (fft_log2(n[axis])) {
      case 6: col_k<6>(kind, g, pa, po); break;
      case 7: col_k<7>(kind, g, pa, po); break;
      case 8: col_k<8>(kind, g, pa, po); break;
      case 9: col_k<9>(kind, g, pa, po); break;
      case 10: col_k<10>(kind, g, pa, po); break;
      default: bad = 1;
    }
  }

  template <int LOGN>
  void row_t(const FftRowGeom& g, long long rows) {
    using A = FftAlg<LOGN>;
    constexpr int RW = fft_rows_per_cta(LOGN);
    std::vector<float2> s(fft_row_smem_bytes(LOGN) / sizeof(float2));
    const std::vector<float2> tw = twiddles(LOGN);
    const float sg = static_cast<float>(sign);
    for (int t = 0; t < ntr; ++t)
      for (long long b = 0; b < rows / RW; ++b) {
        for (int tid = 0; tid < kFftThreads; ++tid) fft_row_first<LOGN>(g, b * RW, t, tid, fw, s.data(), sg);
        if (A::NM >= 1)
          for (int tid = 0; tid < kFftThreads; ++tid) fft_row_middle<LOGN, 1>(tid, s.data(), tw.data(), sg);
        if (A::NM >= 2)
          for (int tid = 0; tid < kFftThreads; ++tid) fft_row_middle<LOGN, (A::NM >= 2 ? 2 : 1)>(tid, s.data(), tw.data(), sg);
        for (int tid = 0; tid < kFftThreads; ++tid) fft_row_last<LOGN>(g, b * RW, t, tid, fw, s.data(), tw.data(), sg);
      }
  }
  void row(const FftRowGeom& g, long long rows) {
    switch (fft_log2(g.n0)) {
      case 6: row_t<6>(g, rows); break;
      case 7: row_t<7>(g, rows); break;
      case 8: row_t<8>(g, rows); break;
      case 9: row_t<9>(g, rows); break;
      case 10: row_t<10>(g, rows); break;
      default: bad = 1;
    }
  }
};

}  // namespace

// f: [ntr][N2][N1][N0] complex64, fw: [ntr][n2][n1][n0] complex64 (x fastest), fac_d: n_d / 2 + 1
// deconvolution factors. type 2: f -> fw (fw need not be initialised); type 1: fw -> f (fw is
// overwritten with intermediate values). Returns 0, or 1 when the sizes are not eligible.
extern "C" int fft_pruned_host(int type, int rank, const int* n, const int* N, int sign, int ntr, float* fw, float* f,
                               const float* fac0, const float* fac1, const float* fac2) {
  long long NN[3] = {N[0], N[1], rank > 2 ? N[2] : 1};
  if (!fft_pruned_ok(rank, n, NN)) return 1;
  const float* fac[3] = {fac0, fac1, fac2};
  std::vector<float> rec[3];   // reciprocal factor tables, rounded once from double (as plan.cu does)
  for (int d = 0; d < rank; ++d) {
    rec[d].resize(n[d] / 2 + 1);
    for (size_t k = 0; k < rec[d].size(); ++k) rec[d][k] = static_cast<float>(1.0 / static_cast<double>(fac[d][k]));
  }
  HostExec ex{type, rank, sign, ntr, n, reinterpret_cast<float2*>(fw), reinterpret_cast<float2*>(f),
              {rec[0].data(), rec[1].data(), rec[2].data()}};
  fft_pruned_sequence(type, rank, n, N, ex);
  return ex.bad;
}
