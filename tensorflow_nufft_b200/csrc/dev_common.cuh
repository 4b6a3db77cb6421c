// dev_common.cuh -- shared device/host helpers for the B200 NUFFT engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace b200 {

constexpr int kWarp = 32;
constexpr int kNumSMsB200 = 148;

// Complex value type per real type.
template <typename F> struct CplxT;
template <> struct CplxT<float> { using type = float2; };
template <> struct CplxT<double> { using type = double2; };
template <typename F> using Cplx = typename CplxT<F>::type;

template <typename F> __host__ __device__ inline Cplx<F> make_cplx(F re, F im);
template <> __host__ __device__ inline float2 make_cplx<float>(float re, float im) { return make_float2(re, im); }
template <> __host__ __device__ inline double2 make_cplx<double>(double re, double im) { return make_double2(re, im); }

// Round-to-nearest primitives that the compiler may not contract into FMAs: the folded
// coordinates (hence bins and kernel offsets) must match the reference's bits.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// Native vector reductions to global memory (REDG.E.ADD.F32x2 / F32x4 on sm_90+).
__device__ __forceinline__ void red_add(float2* p, float2 v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float4* p, float4 v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double2* p, double2 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace b200
