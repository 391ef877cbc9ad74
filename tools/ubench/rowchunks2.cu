// Micro-benchmark 2: how many bytes in flight per SM does the [256 x 32] slab pattern need?  One 512-thread CTA per SM.
//   variant 0: every warp loads 16 rows of the SAME slab (16 loads/thread), slab after slab           (32 KB in flight)
//   variant 1: 4 groups of 4 warps, each group its own slab, 64 loads/thread                          (128 KB in flight)
//   variant 2: variant 1 + st.shared of every value + fence.proxy.async (MEMBAR.ALL.CTA) per slab
//   variant 3: variant 0 with TWO CTAs per SM (256 threads each, 32 rows per warp)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int VAR>
__global__ void __launch_bounds__(512, 1) k(const float* S, int B, int nchunk, long ld, int cols, int per, float* out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long total = (long)B * nchunk;
  const long g0 = (long)blockIdx.x * per, g1 = min(g0 + per, total);
  float acc = 0.f;
  if (VAR == 0) {
    for (long g = g0; g < g1; ++g) {
      const long b = g / nchunk; const int c = (int)(g - b * nchunk); const int kcol = c * 32 + lane;
      const float* p = S + (b * 256 + warp) * ld + kcol;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = kcol < cols ? __ldg(p + (long)i * 16 * ld) : 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i];
    }
  } else {
    const int grp = warp >> 2, wg = warp & 3;
    for (long g = g0 + grp; g < g1; g += 4) {
      const long b = g / nchunk; const int c = (int)(g - b * nchunk); const int kcol = c * 32 + lane;
      const float* p = S + (b * 256 + wg) * ld + kcol;
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = kcol < cols ? __ldg(p + (long)i * 4 * ld) : 0.f;
      if (VAR == 2) {
        asm volatile("" ::: "memory");
#pragma unroll
        for (int i = 0; i < 64; ++i) sm[(grp * 256 + wg + 4 * i) * 32 + lane] = v[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc += v[i];
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv) {
  const int B = 40, rows = 256, cols = 3905; const long ld = 3905; const int nchunk = (cols + 31) / 32;
  float *S[3], *out; for (int i = 0; i < 3; ++i) { cudaMalloc(&S[i], sizeof(float) * B * rows * ld); cudaMemset(S[i], 0, sizeof(float) * B * rows * ld); }
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int per = (B * nchunk + 147) / 148; const int grid = (B * nchunk + per - 1) / per;
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  for (int var = 0; var < 3; ++var) {
    auto launch = [&](int it) {
      if (var == 0) k<0><<<grid, 512>>>(S[it % 3], B, nchunk, ld, cols, per, out);
      if (var == 1) k<1><<<grid, 512>>>(S[it % 3], B, nchunk, ld, cols, per, out);
      if (var == 2) k<2><<<grid, 512, 131072>>>(S[it % 3], B, nchunk, ld, cols, per, out);
    };
    for (int it = 0; it < 3; ++it) launch(it);
    cudaEventRecord(e0);
    for (int it = 0; it < 12; ++it) launch(it);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 12;
    printf("variant %d grid %d: %.1f us  %.0f GB/s  (%s)\n", var, grid, ms * 1e3, 4.0 * B * rows * cols / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
