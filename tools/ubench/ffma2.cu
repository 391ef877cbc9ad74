// Microbenchmark: issue / pipe throughput of FFMA vs FFMA2 (plain, swapped, broadcast operands) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 f2(u64 v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ u64 u2(float2 v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(u2(a)), "l"(u2(b)), "l"(u2(c)));
  return f2(r);
}
template <int MODE>
__global__ void k(float2* out, float2 w, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {          // scalar: 2 FFMA per complex value
        acc[i].x = fmaf(acc[i].x, w.x, w.y);
        acc[i].y = fmaf(acc[i].y, w.x, w.y);
      } else if (MODE == 1) {   // FFMA2 plain
        acc[i] = ffma2(acc[i], w, w);
      } else if (MODE == 2) {   // FFMA2 with swapped + sign-pattern operand
        acc[i] = ffma2(make_float2(acc[i].y, -acc[i].x), w, w);
      } else if (MODE == 3) {   // FFMA2 with broadcast operand
        acc[i] = ffma2(make_float2(acc[(i + 1) & 7].x, acc[(i + 1) & 7].x), w, acc[i]);
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float2* out) {
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, make_float2(0.999f, 0.001f), 16);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, make_float2(0.999f, 0.001f), iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cplx = (double)blocks * threads * iters * 8;      // complex-value FMAs (2 real FMAs each)
  printf("%-28s %.3f ms  %.2f T real-FMA/s  (%.1f TFLOP/s)\n", name, ms, 2 * cplx / ms * 1e-9, 4 * cplx / ms * 1e-9);
}
int main() {
  float2* out; cudaMalloc(&out, 148 * 4 * 512 * sizeof(float2));
  run<0>("FFMA x2 (scalar)", out);
  run<1>("FFMA2 plain", out);
  run<2>("FFMA2 swap+sign operand", out);
  run<3>("FFMA2 broadcast operand", out);
  return 0;
}
