// Microbenchmark: issue-rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  u64 rd; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, float s, int iters) {
  float a[8], b[8];
#pragma unroll
  for (int q = 0; q < 8; q++) { a[q] = threadIdx.x * 1e-3f + q; b[q] = s + q; }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 8; q++) a[q] = fmaf(a[q], b[q], s);          // 8 independent scalar FFMA (3-register form)
    } else {
#pragma unroll
      for (int q = 0; q < 8; q += 2) {                                   // 4 packed FFMA2 = the same 8 FMAs
        float2 r = fma2(make_float2(a[q], a[q + 1]), make_float2(b[q], b[q + 1]), make_float2(s, s));
        a[q] = r.x; a[q + 1] = r.y;
      }
    }
  }
  float t = 0; for (int q = 0; q < 8; q++) t += a[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, 0.999f, iters); else k<1><<<148 * 8, 256>>>(d, 0.999f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * 8.0 * iters;
      if (rep == 2) printf("%s: %.3f ms, %.2f TFLOP/s (FMA = 2 flop)\n", mode ? "FFMA2 (packed)" : "FFMA  (scalar)", ms, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
