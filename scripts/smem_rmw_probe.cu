// smem_rmw_probe.cu -- micro-benchmark behind the spreader's central design choice.
// Measures, on one B200, the throughput of accumulating complex64 values into a shared-memory tile
//   (A) with float atomicAdd, one thread per point walking its 7x7 stencil (the reference's
//       SpreadSubproblem scheme; on sm_100 a shared float atomicAdd is a CAS loop: ATOMS.CAST.SPIN),
//   (B) with lanes laid over the stencil and plain 128-bit load / add / store (this engine's tile
//       kernels: no atomics, conflict-free),
// in "cell updates per second" (one update = one complex64 cell += value). Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/smem_rmw_probe scripts/smem_rmw_probe.cu && build/smem_rmw_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int TX = 40, TY = 40, NS = 7, PTS = 1024, REPS = 64;

__global__ void atomic_kernel(const int2* __restrict__ pos, float2* __restrict__ out) {
  __shared__ float2 tile[TX * TY];
  for (int i = threadIdx.x; i < TX * TY; i += blockDim.x) tile[i] = make_float2(0.f, 0.f);
  __syncthreads();
  for (int rep = 0; rep < REPS; ++rep) {
    for (int p = threadIdx.x; p < PTS; p += blockDim.x) {
      const int2 o = pos[(blockIdx.x * PTS + p) % (PTS * 64)];
      for (int dy = 0; dy < NS; ++dy)
        for (int dx = 0; dx < NS; ++dx) {
          float2* c = &tile[(o.y + dy) * TX + o.x + dx];
          atomicAdd(&c->x, 1.0f);
          atomicAdd(&c->y, 0.5f);
        }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TX * TY; i += blockDim.x) out[blockIdx.x * TX * TY + i] = tile[i];
}

__global__ void rmw_kernel(const int2* __restrict__ pos, float2* __restrict__ out) {
  extern __shared__ float4 tile4[];   // one private tile per warp
  float2* tile = reinterpret_cast<float2*>(tile4) + (threadIdx.x >> 5) * TX * TY;
  const int lane = threadIdx.x & 31;
  for (int i = lane; i < TX * TY; i += 32) tile[i] = make_float2(0.f, 0.f);
  __syncwarp();
  const int q = lane & 3, r = lane >> 2;
  for (int rep = 0; rep < REPS; ++rep) {
    for (int p = 0; p < PTS / 8; ++p) {     // 8 warps share the CTA's 1024 points
      const int2 o = pos[(blockIdx.x * PTS + (threadIdx.x >> 5) * (PTS / 8) + p) % (PTS * 64)];
      if (r < NS) {
        float4* ptr = reinterpret_cast<float4*>(tile + (o.y + r) * TX + (o.x & ~1) + 2 * q);
        float4 v = *ptr;
        v.x += 1.0f; v.y += 0.5f; v.z += 1.0f; v.w += 0.5f;
        *ptr = v;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  for (int i = lane; i < TX * TY; i += 32) out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * TX * TY + i] = tile[i];
}

int main() {
  const int nblk = 148 * 4;
  int2* h = (int2*)malloc(sizeof(int2) * PTS * 64);
  srand(1);
  for (int i = 0; i < PTS * 64; ++i) h[i] = make_int2(rand() % (TX - 8), rand() % (TY - NS));
  int2* d; float2* out;
  cudaMalloc(&d, sizeof(int2) * PTS * 64);
  cudaMalloc(&out, sizeof(float2) * TX * TY * nblk * 8);
  cudaMemcpy(d, h, sizeof(int2) * PTS * 64, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(rmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TX * TY * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0); atomic_kernel<<<nblk, 256>>>(d, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double upd = (double)nblk * PTS * REPS * NS * NS;
  printf("{\"variant\": \"shared float atomicAdd, thread per point (CAS loop)\", \"ms\": %.3f, \"G_cell_updates_per_s\": %.1f}\n", ms, upd / ms / 1e6);
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0); rmw_kernel<<<nblk, 256, 8 * TX * TY * 8>>>(d, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("{\"variant\": \"lanes over stencil, 128-bit load/add/store, private tile per warp\", \"ms\": %.3f, \"G_cell_updates_per_s\": %.1f}\n", ms, upd / ms / 1e6);
  printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
