// fft_pruned.cuh -- the oversampled FFT of a complex64 plan as PRUNED, FUSED one-axis passes.
//
// Stage replaced: cuFFT behind initialize_fft / ThenFft (nufft_plan.cu.cc:2227-2285, :2148-2152)
// together with Amplify{2,3}DKernel / Deconvolve{2,3}DKernel (:326-435). Of the sigma^d fine cells
// only the N^d central modes carry data on the uniform side of the transform, so a full in-place
// FFT moves mostly zeros (type 2) or computes outputs nobody reads (type 1). Here each axis is one
// kernel, and each kernel touches only what the next one needs:
//
//   type 2, 3D (cfg4, nf = 512^3, N = 256^3; G = one fine grid = 1.07 GB)
//     z pass   reads the MODES f (amplified on the fly: no amplify kernel, no zero fill), only the
//              N0 x N1 populated (x, y) columns, writes all z               0.125 G in, 0.25 G out
//     y pass   populated x only, reads the populated y rows, writes all y   0.25 G in,  0.5 G out
//     x pass   every row, reads the populated half, writes all of it        0.5 G in,   1 G out
//   = 2.6 G per transform instead of 8 G (amplify fill + three full in-place passes).
//   type 1 runs the mirror image (x pass keeps only populated x, ..., the last pass divides by the
//   deconvolution factors and writes the modes: no deconvolve kernel). 2D: y and x passes.
//
// The transform of one line is an in-place decimation-in-time FFT, radix 8 (last pass radix 2 or 4
// when log2 n is not a multiple of 3), organised so that an ITEM -- 8 samples t + q n/8 of a line,
// held in registers -- is the unit of every pass: the first pass takes its samples straight from
// global memory (digit reversal = where it puts them in shared memory), the last pass sends its
// results straight to global memory; shared memory only carries the exchanges in between (two for
// n = 512). Strided (y / z) passes: a CTA holds a bundle of W adjacent x columns as s[p][W], every
// global access is a W * 8-byte segment. x pass: 2048 / n rows per CTA, consecutive threads take
// consecutive x (coalesced), rows padded by one cell per 64 so that the exchanges are conflict
// free. Powers of two 64 .. 1024 only; other sizes, complex128 and 1D stay on cuFFT (plan.cu).
//
// The index arithmetic is plain C++ (FFT_HD functions taking block / thread ids) so that
// tests/fft_pruned_host.cc can run the very same code on the CPU against numpy.fft.
#pragma once
#include <cmath>
#include <cstdint>
#if defined(__CUDACC__)
#define FFT_HD __host__ __device__ __forceinline__
#else
#include <vector_functions.h>
#include <vector_types.h>
#define FFT_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define FFT_LDG(p) __ldg(p)
#else
#define FFT_LDG(p) (*(p))
#endif

namespace b200 {

constexpr int kFftThreads = 256;      // x pass: one item per thread
constexpr int kFftColThreads = 512;   // strided passes: 3 CTAs of 64 KB per SM = 48 warps (40 registers)
constexpr int kFftMinLog = 6, kFftMaxLog = 10;
// strided passes: columns per bundle: 16 (128-byte segments), 8 for n = 1024 (64 KB of shared memory)
constexpr int fft_logw(int logn) { return logn >= 10 ? 3 : 4; }
// x pass: rows per CTA (one item per thread) and the padded row pitch (= 8 mod 16 cells)
constexpr int fft_rows_per_cta(int logn) { return kFftThreads / ((1 << logn) / 8); }
constexpr int fft_row_pitch(int logn) { return (((1 << logn) + ((1 << logn) >> 6) + 7) / 16) * 16 + 8; }
inline size_t fft_col_smem_bytes(int logn) { return (static_cast<size_t>(1) << (logn + fft_logw(logn))) * sizeof(float2); }
inline size_t fft_row_smem_bytes(int logn) { return static_cast<size_t>(fft_rows_per_cta(logn)) * fft_row_pitch(logn) * sizeof(float2); }

// mode index i in [0, N) <-> fine index w in [0, n): k = i - N/2, w = k >= 0 ? k : n + k
// (CMCL order, nufft_plan.cu.cc:326-379)
FFT_HD int fft_w_of_mode(int i, int N, int n) {
  const int k = i - N / 2;
  return k >= 0 ? k : n + k;
}
FFT_HD int fft_mode_of_w(int w, int N, int n) {   // -1: zero padding
  if (w <= (N - 1) / 2) return w + N / 2;
  if (w >= n - N / 2) return w - n + N / 2;
  return -1;
}

// branch-free "w is a populated fine index" for loads / stores that must stay predicated
struct FftPop {
  int hi, lo;   // populated: w <= hi or w >= lo
};
FFT_HD FftPop fft_pop(int N, int n, int active) {
  FftPop p;
  p.hi = active ? (N - 1) / 2 : n;
  p.lo = active ? n - N / 2 : 0;
  return p;
}
FFT_HD bool fft_is_pop(const FftPop& p, int w) { return (w <= p.hi) | (w >= p.lo); }

// Complex primitives. On the device every one is ONE or TWO packed f32x2 instructions (FADD2 /
// FMUL2 / FFMA2: both components per issue slot; the half swap of a multiplication by +-i is an
// operand modifier), on the host (CPU test harness) plain scalar code.
//   addrot(b, d) = b + (sg i) d     subrot(b, d) = b - (sg i) d
//   w8p(a) = a (1 + sg i) / sqrt 2  w8m(a) = a (-1 + sg i) / sqrt 2
#if defined(__CUDA_ARCH__)
#define FFT_P2(op, r, a, b)                                                                               \
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t" op               \
      ".rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"                                               \
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y))
__device__ __forceinline__ float2 fft_fma2(float ax, float ay, float bx, float by, float cx, float cy) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
  return r;
}
__device__ __forceinline__ float2 fft_add(float2 a, float2 b) { float2 r; FFT_P2("add", r, a, b); return r; }
__device__ __forceinline__ float2 fft_sub(float2 a, float2 b) { float2 r; FFT_P2("sub", r, a, b); return r; }
__device__ __forceinline__ float2 fft_mul2(float2 a, float2 b) { float2 r; FFT_P2("mul", r, a, b); return r; }
__device__ __forceinline__ float2 fft_addrot(float2 b, float2 d, float sg) { return fft_fma2(-sg, sg, d.y, d.x, b.x, b.y); }
__device__ __forceinline__ float2 fft_subrot(float2 b, float2 d, float sg) { return fft_fma2(sg, -sg, d.y, d.x, b.x, b.y); }
__device__ __forceinline__ float2 fft_w8p(float2 a, float sg) {
  constexpr float h = 0.70710678118654752f;
  return fft_mul2(make_float2(h, h), fft_fma2(-sg, sg, a.y, a.x, a.x, a.y));
}
__device__ __forceinline__ float2 fft_w8m(float2 a, float sg) {
  constexpr float h = 0.70710678118654752f;
  return fft_mul2(make_float2(h, h), fft_fma2(-sg, sg, a.y, a.x, -a.x, -a.y));
}
#else
FFT_HD float2 fft_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFT_HD float2 fft_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFT_HD float2 fft_addrot(float2 b, float2 d, float sg) { return make_float2(b.x - sg * d.y, b.y + sg * d.x); }
FFT_HD float2 fft_subrot(float2 b, float2 d, float sg) { return make_float2(b.x + sg * d.y, b.y - sg * d.x); }
FFT_HD float2 fft_w8p(float2 a, float sg) {
  constexpr float h = 0.70710678118654752f;
  return make_float2(h * (a.x - sg * a.y), h * (a.y + sg * a.x));
}
FFT_HD float2 fft_w8m(float2 a, float sg) {
  constexpr float h = 0.70710678118654752f;
  return make_float2(h * (-a.x - sg * a.y), h * (-a.y + sg * a.x));
}
#endif
FFT_HD float2 fft_cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// the 7 twiddles w^q of a radix-8 butterfly from the three stored ones (rows q = 1, 2, 4 of a table
// with L entries per row): 3 loads + 4 products instead of 7 loads -- the passes are bound by L1 /
// shared-memory wavefronts, not by issue slots
FFT_HD void fft_twiddles8(const float2* tab, int L, int j, float2* w) {
  w[1] = FFT_LDG(tab + j);
  w[2] = FFT_LDG(tab + L + j);
  w[4] = FFT_LDG(tab + 2 * L + j);
  w[3] = fft_cmul(w[1], w[2]);
  w[5] = fft_cmul(w[1], w[4]);
  w[6] = fft_cmul(w[2], w[4]);
  w[7] = fft_cmul(w[3], w[4]);
}

// r-point DFT X_s = sum_q a_q exp(sg * 2 pi i q s / r), in place, natural order in and out
FFT_HD void fft_dft2(float2& a0, float2& a1) {
  const float2 t = a0;
  a0 = fft_add(t, a1);
  a1 = fft_sub(t, a1);
}
FFT_HD void fft_dft4(float2& a0, float2& a1, float2& a2, float2& a3, float sg) {
  const float2 b0 = fft_add(a0, a2), b1 = fft_sub(a0, a2), b2 = fft_add(a1, a3), d = fft_sub(a1, a3);
  a0 = fft_add(b0, b2);
  a1 = fft_addrot(b1, d, sg);
  a2 = fft_sub(b0, b2);
  a3 = fft_subrot(b1, d, sg);
}
FFT_HD void fft_dft8(float2* a, float sg) {
  fft_dft4(a[0], a[2], a[4], a[6], sg);   // even inputs -> E_0..3 in a[0], a[2], a[4], a[6]
  fft_dft4(a[1], a[3], a[5], a[7], sg);   // odd inputs  -> O_0..3 in a[1], a[3], a[5], a[7]
  const float2 o0 = a[1], o1 = fft_w8p(a[3], sg), o2 = a[5], o3 = fft_w8m(a[7], sg);
  const float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
  a[0] = fft_add(e0, o0);
  a[1] = fft_add(e1, o1);
  a[2] = fft_addrot(e2, o2, sg);
  a[3] = fft_add(e3, o3);
  a[4] = fft_sub(e0, o0);
  a[5] = fft_sub(e1, o1);
  a[6] = fft_subrot(e2, o2, sg);
  a[7] = fft_sub(e3, o3);
}

// Pass structure of a length-2^LOGN transform: NP8 radix-8 passes (sub-transform lengths L = 1, 8,
// 64, ...), then one radix-2^RL pass if LOGN is not a multiple of 3. Pass 0 and the last pass work
// on ITEMS (t in [0, n/8): samples / results t + q n/8); the NM passes in between on butterflies u.
template <int LOGN>
struct FftAlg {
  static constexpr int n = 1 << LOGN;
  static constexpr int RL = LOGN % 3;
  static constexpr int NP8 = LOGN / 3;
  static constexpr int NM = NP8 - 1 - (RL == 0 ? 1 : 0);
  static constexpr int T = n / 8;
  static constexpr int kLastR = RL == 0 ? 8 : (1 << RL);
  static constexpr int kLastL = n / kLastR;
  // twiddle table: middle pass k (L = 8^k) keeps rows q = 1, 2, 4 of exp(sg 2 pi i j q / (8 L)), j < L, at
  // tw_mid(k) + row * L + j; the last pass rows q = 1, 2, 4 (radix 8) / 1, 2 (radix 4) / 1 (radix 2) of
  // exp(sg 2 pi i j q / n), j < kLastL, at kTwLast + row * kLastL + j
  static constexpr int tw_mid(int k) {
    int o = 0;
    for (int i = 1; i < k; ++i) o += 3 << (3 * i);
    return o;
  }
  static constexpr int kTwLast = tw_mid(NM + 1);
  static constexpr int kTwLastRows = RL == 0 ? 3 : RL;
  static constexpr int kTwCount = kTwLast + kTwLastRows * kLastL;

  // where sample i of the natural order sits before pass 0 (mixed-radix digit reversal: the digit
  // of the last pass, the lowest of i, is the highest of the position)
  static FFT_HD int pos(int i) {
    int p = 0, rem = n;
    if (RL) {
      rem >>= RL;
      p += (i & ((1 << RL) - 1)) * rem;
      i >>= RL;
    }
#pragma unroll
    for (int k = 0; k < NP8; ++k) {
      rem >>= 3;
      p += (i & 7) * rem;
      i >>= 3;
    }
    return p;
  }
};

inline int fft_tw_count(int logn) {
  switch (logn) {
    case 6: return FftAlg<6>::kTwCount;
    case 7: return FftAlg<7>::kTwCount;
    case 8: return FftAlg<8>::kTwCount;
    case 9: return FftAlg<9>::kTwCount;
    case 10: return FftAlg<10>::kTwCount;
    default: return 0;
  }
}
template <int LOGN>
inline void fft_fill_twiddles_t(int sign, float2* out) {
  using A = FftAlg<LOGN>;
  const double pi2 = 6.283185307179586476925286766559 * sign;
  auto w = [&](double num, double den) {
    const double a = pi2 * num / den;
    return make_float2(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
  };
  for (int k = 1; k <= A::NM; ++k) {
    const int L = 1 << (3 * k);
    for (int r = 0; r < 3; ++r)
      for (int j = 0; j < L; ++j) out[A::tw_mid(k) + r * L + j] = w(static_cast<double>(j) * (1 << r), 8.0 * L);
  }
  for (int r = 0; r < A::kTwLastRows; ++r)
    for (int j = 0; j < A::kLastL; ++j) out[A::kTwLast + r * A::kLastL + j] = w(static_cast<double>(j) * (1 << r), A::n);
}
inline void fft_fill_twiddles(int logn, int sign, float2* out) {
  switch (logn) {
    case 6: fft_fill_twiddles_t<6>(sign, out); break;
    case 7: fft_fill_twiddles_t<7>(sign, out); break;
    case 8: fft_fill_twiddles_t<8>(sign, out); break;
    case 9: fft_fill_twiddles_t<9>(sign, out); break;
    case 10: fft_fill_twiddles_t<10>(sign, out); break;
    default: break;
  }
}

// The three kinds of pass over shared memory s; LY::addr(line, p) places position p of a line.
template <int LOGN, class LY>
struct FftPhases {
  using A = FftAlg<LOGN>;
  // pass 0: a[q] = sample t + q n/8 of the line
  static FFT_HD void first(float2* s, int line, int t, float2* a, float sg) {
    fft_dft8(a, sg);
    const int p0 = A::pos(t);
#pragma unroll
    for (int q = 0; q < 8; ++q) s[LY::addr(line, p0 + q)] = a[q];
  }
  // radix-8 pass over sub-transforms of length L = 8^K, butterfly u in [0, n/8)
  template <int K>
  static FFT_HD void middle(float2* s, const float2* tw, int line, int u, float sg) {
    constexpr int LOGL = 3 * K, L = 1 << LOGL;
    const int j = u & (L - 1);
    const int base = ((u >> LOGL) << (LOGL + 3)) + j;
    float2 a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = s[LY::addr(line, base + q * L)];
    float2 w[8];
    fft_twiddles8(tw + A::tw_mid(K), L, j, w);
#pragma unroll
    for (int q = 1; q < 8; ++q) a[q] = fft_cmul(a[q], w[q]);
    fft_dft8(a, sg);
#pragma unroll
    for (int q = 0; q < 8; ++q) s[LY::addr(line, base + q * L)] = a[q];
  }
  // last pass: returns a[e] = result t + e n/8 of the line
  static FFT_HD void last(const float2* s, const float2* tw, int line, int t, float2* a, float sg) {
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = s[LY::addr(line, t + e * A::T)];
    constexpr int L = A::kLastL;
    if (A::RL == 0) {
      float2 w[8];
      fft_twiddles8(tw + A::kTwLast, L, t, w);
#pragma unroll
      for (int q = 1; q < 8; ++q) a[q] = fft_cmul(a[q], w[q]);
      fft_dft8(a, sg);
    } else if (A::RL == 1) {   // four radix-2 butterflies j = t + m n/8: (a[m], a[m + 4])
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        a[m + 4] = fft_cmul(a[m + 4], FFT_LDG(tw + A::kTwLast + t + m * A::T));
        fft_dft2(a[m], a[m + 4]);
      }
    } else {                   // two radix-4 butterflies j = t + m n/8: (a[m], a[m + 2], a[m + 4], a[m + 6])
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int j = t + m * A::T;
        const float2 w1 = FFT_LDG(tw + A::kTwLast + j), w2 = FFT_LDG(tw + A::kTwLast + L + j);
        a[m + 2] = fft_cmul(a[m + 2], w1);
        a[m + 4] = fft_cmul(a[m + 4], w2);
        a[m + 6] = fft_cmul(a[m + 6], fft_cmul(w1, w2));
        fft_dft4(a[m], a[m + 2], a[m + 4], a[m + 6], sg);
      }
    }
  }
};

template <int LOGW>
struct FftColLayout {
  static FFT_HD int addr(int c, int p) { return (p << LOGW) + c; }
};
template <int LOGN>
struct FftRowLayout {
  static FFT_HD int addr(int row, int p) { return row * fft_row_pitch(LOGN) + p + (p >> 6); }
};

// ---------------------------------------------------------------------------------------------
// Strided pass: FFT along an axis of stride `stride` for bundles of W adjacent x columns.
// ---------------------------------------------------------------------------------------------
struct FftColGeom {
  int n, N;                 // fine size / modes of the transformed axis
  long long stride;         // its element stride in the fine grid
  int n0, N0;               // fine size / modes of the x axis (bundles are groups of W populated x)
  int outer_count;          // bundles along the remaining axis (grid y)
  int outer_N, outer_n;     // outer_pop: grid y is a MODE index of that axis, else a fine index
  long long outer_stride;   // its element stride in the fine grid
  int outer_pop;
  int in_pop, out_pop;      // kind 0: read only the populated entries / write only the populated entries
  long long nftot, ntot;    // fine cells / modes per transform
  long long f_stride, f_outer_stride;   // strides of the two axes in the mode array
};

enum { kFftPlain = 0, kFftFromModes = 1, kFftToModes = 2 };

// The pass that touches the mode array multiplies by 1 / (p_axis p_outer p_x): the three tables
// hold the reciprocals of the deconvolution factors (rounded once from double, plan.cu), the x
// and outer ones are folded into one factor per thread.
FFT_HD float fft_thread_factor(const FftColGeom& g, int k0, int o, const float* ro, const float* rx) {
  float r = rx[k0 < 0 ? -k0 : k0];
  if (ro != nullptr) {
    const int ko = o - g.outer_N / 2;
    r *= ro[ko < 0 ? -ko : ko];
  }
  return r;
}

struct FftColCtx {   // what every thread of a CTA derives from its block index
  int k0g;           // signed x mode of the bundle's first column
  const float2* src;
  float2* dst;
  const float2* fsrc;
  float2* fdst;
};
FFT_HD FftColCtx fft_col_ctx(const FftColGeom& g, int logw, int gx, int o, int tr, float2* fw, float2* f) {
  FftColCtx c;
  const int i0g = gx << logw;
  c.k0g = i0g - g.N0 / 2;
  const int w0g = c.k0g >= 0 ? c.k0g : g.n0 + c.k0g;
  const int wo = g.outer_pop ? fft_w_of_mode(o, g.outer_N, g.outer_n) : o;
  c.dst = fw + static_cast<long long>(tr) * g.nftot + wo * g.outer_stride + w0g;
  c.src = c.dst;
  c.fdst = f + static_cast<long long>(tr) * g.ntot + o * g.f_outer_stride + i0g;
  c.fsrc = c.fdst;
  return c;
}

template <int LOGN, int KIND>
FFT_HD void fft_col_first(const FftColGeom& g, const FftColCtx& cx, int o, int tid, int nthr, const float* ra,
                          const float* ro, const float* rx, float2* s, float sg) {
  using A = FftAlg<LOGN>;
  constexpr int LOGW = fft_logw(LOGN), W = 1 << LOGW;
  using PH = FftPhases<LOGN, FftColLayout<LOGW>>;
  const int c = tid & (W - 1);
  const FftPop pop = fft_pop(g.N, g.n, KIND == kFftFromModes || g.in_pop);
  const float rt = KIND == kFftFromModes ? fft_thread_factor(g, cx.k0g + c, o, ro, rx) : 0.f;
  const int half = g.N / 2;
  constexpr int kItems = (A::T * W + kFftColThreads - 1) / kFftColThreads;
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    const int t = (tid >> LOGW) + it * (nthr >> LOGW);
    if (A::T * W % kFftColThreads != 0 && t >= A::T) break;
    float2 a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int w = t + q * A::T;
      const bool ok = fft_is_pop(pop, w);
      if (KIND == kFftFromModes) {
        const int k = w <= pop.hi ? w : w - g.n;   // signed mode of a populated w
        const float2 v = ok ? cx.fsrc[(k + half) * g.f_stride + c] : make_float2(0.f, 0.f);
        const float r = ok ? rt * FFT_LDG(ra + (k < 0 ? -k : k)) : 0.f;
        a[q] = make_float2(v.x * r, v.y * r);
      } else {
        a[q] = ok ? cx.src[w * g.stride + c] : make_float2(0.f, 0.f);
      }
    }
    PH::first(s, c, t, a, sg);
  }
}

template <int LOGN, int K>
FFT_HD void fft_col_middle(int tid, int nthr, float2* s, const float2* tw, float sg) {
  using A = FftAlg<LOGN>;
  constexpr int LOGW = fft_logw(LOGN), W = 1 << LOGW;
  using PH = FftPhases<LOGN, FftColLayout<LOGW>>;
  const int c = tid & (W - 1);
  for (int u = tid >> LOGW; u < A::T; u += nthr >> LOGW) PH::template middle<K>(s, tw, c, u, sg);
}

template <int LOGN, int KIND>
FFT_HD void fft_col_last(const FftColGeom& g, const FftColCtx& cx, int o, int tid, int nthr, const float* ra,
                         const float* ro, const float* rx, const float2* s, const float2* tw, float sg) {
  using A = FftAlg<LOGN>;
  constexpr int LOGW = fft_logw(LOGN), W = 1 << LOGW;
  using PH = FftPhases<LOGN, FftColLayout<LOGW>>;
  const int c = tid & (W - 1);
  const FftPop pop = fft_pop(g.N, g.n, KIND == kFftToModes || g.out_pop);
  const float rt = KIND == kFftToModes ? fft_thread_factor(g, cx.k0g + c, o, ro, rx) : 0.f;
  const int half = g.N / 2;
  for (int t = tid >> LOGW; t < A::T; t += nthr >> LOGW) {   // not unrolled: two items in flight spill at 40 registers
    float2 a[8];
    PH::last(s, tw, c, t, a, sg);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int w = t + e * A::T;
      if (fft_is_pop(pop, w)) {
        if (KIND == kFftToModes) {
          const int k = w <= pop.hi ? w : w - g.n;
          const float r = rt * FFT_LDG(ra + (k < 0 ? -k : k));
          cx.fdst[(k + half) * g.f_stride + c] = make_float2(a[e].x * r, a[e].y * r);
        } else {
          cx.dst[w * g.stride + c] = a[e];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// x pass: 2048 / n rows per CTA, one item per thread.
// ---------------------------------------------------------------------------------------------
struct FftRowGeom {
  int n0, N0;
  int in_pop, out_pop;
  long long nftot;
};

template <int LOGN>
FFT_HD void fft_row_first(const FftRowGeom& g, long long row0, int tr, int tid, const float2* fw, float2* s, float sg) {
  using A = FftAlg<LOGN>;
  using PH = FftPhases<LOGN, FftRowLayout<LOGN>>;
  const int row = tid / A::T, t = tid % A::T;
  const float2* src = fw + static_cast<long long>(tr) * g.nftot + (row0 + row) * A::n;
  const FftPop pop = fft_pop(g.N0, g.n0, g.in_pop);
  float2 a[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int w = t + q * A::T;
    a[q] = fft_is_pop(pop, w) ? src[w] : make_float2(0.f, 0.f);
  }
  PH::first(s, row, t, a, sg);
}

// thread -> (row, butterfly) of a middle pass, chosen so that a half warp's 8-byte accesses fall
// into 16 distinct bank pairs (row pitch = 8 mod 16 cells, one pad cell per 64)
template <int LOGN, int K>
FFT_HD void fft_row_middle(int tid, float2* s, const float2* tw, float sg) {
  using A = FftAlg<LOGN>;
  using PH = FftPhases<LOGN, FftRowLayout<LOGN>>;
  int row, u;
  if (LOGN == 9 && K == 1) {
    row = (tid >> 3) & 3;
    u = ((tid >> 5) << 3) + (tid & 7);
  } else if (LOGN == 10 && K == 1) {
    row = tid >> 7;
    u = (((((tid >> 3) & 1) << 3) + ((tid >> 4) & 7)) << 3) + (tid & 7);
  } else {
    row = tid / A::T;
    u = tid % A::T;
  }
  PH::template middle<K>(s, tw, row, u, sg);
}

template <int LOGN>
FFT_HD void fft_row_last(const FftRowGeom& g, long long row0, int tr, int tid, float2* fw, const float2* s,
                         const float2* tw, float sg) {
  using A = FftAlg<LOGN>;
  using PH = FftPhases<LOGN, FftRowLayout<LOGN>>;
  const int row = tid / A::T, t = tid % A::T;
  float2* dst = fw + static_cast<long long>(tr) * g.nftot + (row0 + row) * A::n;
  const FftPop pop = fft_pop(g.N0, g.n0, g.out_pop);
  float2 a[8];
  PH::last(s, tw, row, t, a, sg);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int w = t + e * A::T;
    if (fft_is_pop(pop, w)) dst[w] = a[e];
  }
}

// ---------------------------------------------------------------------------------------------
// The sequence of passes of one transform batch. `ex.col(axis, kind, geom, axis_a, axis_o)` runs a
// strided pass along `axis` (axis_a / axis_o: which axes' deconvolution factors a from-modes /
// to-modes pass divides by, -1 none), `ex.row(geom, rows)` the x pass. plan.cu launches the
// kernels from it; the CPU test harness runs the same functions in loops. The factor tables the
// from-modes / to-modes pass uses are the RECIPROCALS 1 / phihat_d[|k|].
// ---------------------------------------------------------------------------------------------
template <class Exec>
inline void fft_pruned_sequence(int type, int rank, const int* n, const int* N, Exec& ex) {
  const long long n0 = n[0], n1 = n[1], n2 = rank > 2 ? n[2] : 1;
  const long long nftot = n0 * n1 * n2;
  const long long ntot = static_cast<long long>(N[0]) * N[1] * (rank > 2 ? N[2] : 1);
  FftColGeom g{};
  g.n0 = n[0];
  g.N0 = N[0];
  g.nftot = nftot;
  g.ntot = ntot;
  FftRowGeom rg{n[0], N[0], 0, 0, nftot};
  auto modes_pass = [&](int kind) {   // slowest axis <-> the mode array
    if (rank == 3) {
      g.n = n[2]; g.N = N[2]; g.stride = n0 * n1;
      g.outer_count = N[1]; g.outer_N = N[1]; g.outer_n = n[1]; g.outer_stride = n0; g.outer_pop = 1;
      g.f_stride = static_cast<long long>(N[0]) * N[1]; g.f_outer_stride = N[0];
      g.in_pop = g.out_pop = 0;
      ex.col(2, kind, g, 2, 1);
    } else {
      g.n = n[1]; g.N = N[1]; g.stride = n0;
      g.outer_count = 1; g.outer_N = 1; g.outer_n = 1; g.outer_stride = 0; g.outer_pop = 0;
      g.f_stride = N[0]; g.f_outer_stride = 0;
      g.in_pop = g.out_pop = 0;
      ex.col(1, kind, g, 1, -1);
    }
  };
  auto y_pass_3d = [&](int in_pop, int out_pop) {
    g.n = n[1]; g.N = N[1]; g.stride = n0;
    g.outer_count = n[2]; g.outer_N = N[2]; g.outer_n = n[2]; g.outer_stride = n0 * n1; g.outer_pop = 0;
    g.f_stride = g.f_outer_stride = 0;
    g.in_pop = in_pop; g.out_pop = out_pop;
    ex.col(1, kFftPlain, g, -1, -1);
  };
  if (type == 2) {
    modes_pass(kFftFromModes);
    if (rank == 3) y_pass_3d(1, 0);
    rg.in_pop = 1;
    ex.row(rg, n1 * n2);
  } else {
    rg.out_pop = 1;
    ex.row(rg, n1 * n2);
    if (rank == 3) y_pass_3d(0, 1);
    modes_pass(kFftToModes);
  }
}

// eligibility of a plan: complex64, rank 2 or 3, every fine size a power of two in [64, 1024], the
// x modes a multiple of 32 (groups of 16 populated x never straddle k = 0) and not wider than the
// fine grid
template <class I>
inline bool fft_pruned_ok(int rank, const int* n, const I* N) {
  if (rank != 2 && rank != 3) return false;
  for (int d = 0; d < rank; ++d) {
    if (n[d] < (1 << kFftMinLog) || n[d] > (1 << kFftMaxLog) || (n[d] & (n[d] - 1)) != 0) return false;
    if (N[d] < 1 || N[d] > n[d]) return false;
  }
  return N[0] % 32 == 0;
}
inline int fft_log2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

#if defined(__CUDACC__)
// grid (N0 / W, outer_count, transforms); dynamic shared memory fft_col_smem_bytes(LOGN)
template <int LOGN, int KIND>
__global__ void __launch_bounds__(kFftColThreads, 3)
fft_col_kernel(FftColGeom g, float sg, const float2* __restrict__ tw, float2* fw, float2* f,
               const float* __restrict__ ra, const float* __restrict__ ro, const float* __restrict__ rx) {
  using A = FftAlg<LOGN>;
  extern __shared__ __align__(16) float2 fft_smem[];
  const int tid = threadIdx.x;
  const int o = blockIdx.y;
  const FftColCtx cx = fft_col_ctx(g, fft_logw(LOGN), blockIdx.x, o, blockIdx.z, fw, f);
  fft_col_first<LOGN, KIND>(g, cx, o, tid, kFftColThreads, ra, ro, rx, fft_smem, sg);
  __syncthreads();
  if (A::NM >= 1) {
    fft_col_middle<LOGN, 1>(tid, kFftColThreads, fft_smem, tw, sg);
    __syncthreads();
  }
  if (A::NM >= 2) {
    fft_col_middle<LOGN, (A::NM >= 2 ? 2 : 1)>(tid, kFftColThreads, fft_smem, tw, sg);
    __syncthreads();
  }
  fft_col_last<LOGN, KIND>(g, cx, o, tid, kFftColThreads, ra, ro, rx, fft_smem, tw, sg);
}

// grid (rows / rows_per_cta, transforms); dynamic shared memory fft_row_smem_bytes(LOGN)
template <int LOGN>
__global__ void __launch_bounds__(kFftThreads)
fft_row_kernel(FftRowGeom g, float sg, const float2* __restrict__ tw, float2* fw) {
  using A = FftAlg<LOGN>;
  extern __shared__ __align__(16) float2 fft_smem[];
  const int tid = threadIdx.x;
  const long long row0 = static_cast<long long>(blockIdx.x) * fft_rows_per_cta(LOGN);
  fft_row_first<LOGN>(g, row0, blockIdx.y, tid, fw, fft_smem, sg);
  __syncthreads();
  if (A::NM >= 1) {
    fft_row_middle<LOGN, 1>(tid, fft_smem, tw, sg);
    __syncthreads();
  }
  if (A::NM >= 2) {
    fft_row_middle<LOGN, (A::NM >= 2 ? 2 : 1)>(tid, fft_smem, tw, sg);
    __syncthreads();
  }
  fft_row_last<LOGN>(g, row0, blockIdx.y, tid, fw, fft_smem, tw, sg);
}
#endif  // __CUDACC__

}  // namespace b200
