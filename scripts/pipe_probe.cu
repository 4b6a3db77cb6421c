// pipe_probe.cu -- micro-benchmarks behind the round-2 spreader design (B200, sm_100a):
//   * shared-memory load cost (LSU wavefronts) of the access patterns a spreader's inner loop uses:
//     32 / 64 / 128-bit loads that are warp-uniform, uniform per 8-lane group, or 8 distinct rows;
//   * FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate;
//   * SHFL rate (does broadcasting through shuffles relieve the LSU pipe?).
// Every kernel is timed with clock64() inside one persistent CTA per SM (8 warps), so the result is
// "warp-instructions per clock per SM", independent of the SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_probe scripts/pipe_probe.cu && build/pipe_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int kIters = 4096;
constexpr int kUnroll = 8;

enum Pattern { kUniform = 0, kPerGroup8 = 1, kRows8 = 2, kDistinct = 3 };

__device__ __forceinline__ unsigned lane_offset_bytes(int pattern, int width_bytes, int lane) {
  switch (pattern) {
    case kUniform: return 0u;
    case kPerGroup8: return (lane >> 3) * width_bytes;              // 4 distinct addresses, adjacent
    case kRows8: return (lane & 7) * 208u + (lane >> 3) * 3328u;    // 8 rows of pitch 208 B, 4 tiles 3328 B apart
    default: return lane * width_bytes;                             // fully distinct, contiguous
  }
}

template <int WIDTH>   // bytes per lane: 4, 8, 16
__global__ void __launch_bounds__(256) lds_kernel(int pattern, float* out, long long* cycles) {
  extern __shared__ __align__(128) float smem[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) smem[i] = 1.0f + i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(smem)) + lane_offset_bytes(pattern, WIDTH, lane);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const unsigned a = base + u * 16384u / kUnroll * 0u + ((it + u) & 7) * 16u;   // small moving offset, same pattern
      if (WIDTH == 4) {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
        acc += x;
      } else if (WIDTH == 8) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
        acc += x + y;
      } else {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
        acc += x + y + z + w;
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) ffma_kernel(float* out, long long* cycles, float s) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  const float b = s + 1.0f, c = s * 0.5f;
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float* out, long long* cycles, float s) {
  unsigned long long a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x = threadIdx.x + i, y = threadIdx.x - i;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(x), "f"(y));
  }
  unsigned long long b, c;
  const float b0 = s + 1.0f, c0 = s * 0.5f;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b0));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c0));
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x, y;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i]));
    r += x + y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) shfl_kernel(float* out, long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], (it + i) & 31);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// Mixed loop shaped like the planned 2D spreader inner loop: per "point" 4 uniform LDS.64 (wx), one
// LDS.32 (wy, 8 rows), two per-group LDS.64 (strengths), one uniform LDS.32 (header), 32 FFMA.
__global__ void __launch_bounds__(256) mix_kernel(float* out, long long* cycles, int packed) {
  extern __shared__ __align__(128) float smem[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) smem[i] = 1e-3f * (i & 255);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(smem)) + warp * 4096u;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    const unsigned rec = sbase + (it & 31) * 112u;
    float wx[8], wy, c[4], hdr;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(wx[2 * k]), "=f"(wx[2 * k + 1]) : "r"(rec + 8u * k));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wy) : "r"(rec + 32u + 4u * (lane & 7)));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(c[0]), "=f"(c[1]) : "r"(rec + 64u + 16u * (lane >> 3)));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(c[2]), "=f"(c[3]) : "r"(rec + 72u + 16u * (lane >> 3)));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hdr) : "r"(rec + 96u));
    const float c0 = c[0] * wy, c1 = c[1] * wy, c2 = c[2] * wy, c3 = c[3] * wy;
    if (!packed) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[4 * k + 0] = fmaf(c0, wx[k], acc[4 * k + 0]);
        acc[4 * k + 1] = fmaf(c1, wx[k], acc[4 * k + 1]);
        acc[4 * k + 2] = fmaf(c2, wx[k], acc[4 * k + 2]);
        acc[4 * k + 3] = fmaf(c3, wx[k], acc[4 * k + 3]);
      }
    } else {
      // accumulators paired over adjacent cells: (cell 2k, cell 2k+1) += splat(c) * (wx[2k], wx[2k+1])
      unsigned long long s0, s1, s2, s3;
      asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s0) : "f"(c0));
      asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s1) : "f"(c1));
      asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s2) : "f"(c2));
      asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s3) : "f"(c3));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned long long w2;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w2) : "f"(wx[2 * k]), "f"(wx[2 * k + 1]));
        unsigned long long* a2 = reinterpret_cast<unsigned long long*>(&acc[8 * k]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[0]) : "l"(s0), "l"(w2));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[1]) : "l"(s1), "l"(w2));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[2]) : "l"(s2), "l"(w2));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[3]) : "l"(s3), "l"(w2));
      }
    }
    if (__float_as_int(hdr) == 0x7fffffff) acc[0] += 1.f;   // keeps the header load alive
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static double mean_cycles(long long* d_cycles, int blocks) {
  long long* h = static_cast<long long*>(malloc(sizeof(long long) * blocks));
  cudaMemcpy(h, d_cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < blocks; ++i) s += h[i];
  free(h);
  return s / blocks;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms, threads = 256, warps = threads / 32;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  const char* pnames[] = {"uniform", "per8group", "rows8x208B", "distinct"};
  cudaFuncSetAttribute(lds_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(lds_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int rep = 0; rep < 2; ++rep) {
    for (int p = 0; p < 4; ++p) {
      for (int w = 4; w <= 16; w *= 2) {
        if (w == 4) lds_kernel<4><<<blocks, threads, 65536>>>(p, out, cyc);
        if (w == 8) lds_kernel<8><<<blocks, threads, 65536>>>(p, out, cyc);
        if (w == 16) lds_kernel<16><<<blocks, threads, 65536>>>(p, out, cyc);
        cudaDeviceSynchronize();
        const double c = mean_cycles(cyc, blocks);
        const double instr = static_cast<double>(warps) * kIters * kUnroll;
        if (rep) printf("{\"probe\": \"lds\", \"bits\": %d, \"pattern\": \"%s\", \"clk_per_warp_instr_per_sm\": %.3f}\n",
                        w * 8, pnames[p], c / instr);
      }
    }
    ffma_kernel<<<blocks, threads>>>(out, cyc, 0.001f);
    cudaDeviceSynchronize();
    double c = mean_cycles(cyc, blocks);
    if (rep) printf("{\"probe\": \"ffma\", \"warp_instr_per_clk_per_sm\": %.3f}\n", warps * 16.0 * kIters / c);
    ffma2_kernel<<<blocks, threads>>>(out, cyc, 0.001f);
    cudaDeviceSynchronize();
    c = mean_cycles(cyc, blocks);
    if (rep) printf("{\"probe\": \"ffma2\", \"warp_instr_per_clk_per_sm\": %.3f, \"fma_lanes_per_clk_per_sm\": %.1f}\n",
                    warps * 16.0 * kIters / c, warps * 16.0 * kIters / c * 64);
    shfl_kernel<<<blocks, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    c = mean_cycles(cyc, blocks);
    if (rep) printf("{\"probe\": \"shfl\", \"clk_per_warp_instr_per_sm\": %.3f}\n", c / (warps * 8.0 * kIters));
    for (int packed = 0; packed < 2; ++packed) {
      mix_kernel<<<blocks, threads, 65536>>>(out, cyc, packed);
      cudaDeviceSynchronize();
      c = mean_cycles(cyc, blocks);
      if (rep) printf("{\"probe\": \"mix_point_loop\", \"packed_ffma2\": %d, \"warps_per_sm\": %d, \"clk_per_point_per_warp\": %.2f, "
                      "\"clk_per_point_per_sm\": %.2f}\n", packed, warps, c / kIters, c / kIters / warps);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
