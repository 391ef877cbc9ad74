// Micro-benchmark: streaming kernels with the read/write mix of the pipeline's passes (40 M floats):
//   r1w0 (Gram), r1w1 (copy, STFT-like), r1w2 (rank-1 projection: read the image, write S and D)
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void k(const float4* x, float4* y1, float4* y2, long n4, float* sink) {
  float acc = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = x[i];
    if (W >= 1) y1[i] = make_float4(v.x * 2.f, v.y * 2.f, v.z * 2.f, v.w * 2.f);
    if (W >= 2) y2[i] = make_float4(v.x + 1.f, v.y + 1.f, v.z + 1.f, v.w + 1.f);
    if (W == 0) acc += v.x + v.y + v.z + v.w;
  }
  if (W == 0 && acc == 123.f) *sink = acc;
}
int main() {
  const long n = 40L * 256 * 3905, n4 = n / 4;
  float4 *x[3], *y1[3], *y2[3]; float* sink; cudaMalloc(&sink, 4);
  for (int i = 0; i < 3; ++i) { cudaMalloc(&x[i], n * 4); cudaMalloc(&y1[i], n * 4); cudaMalloc(&y2[i], n * 4); cudaMemset(x[i], 0, n * 4); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) for (int bps = 2; bps <= 8; bps *= 2) {
    const int grid = 148 * bps;
    auto go = [&](int it) {
      if (w == 0) k<0><<<grid, 512>>>(x[it % 3], y1[it % 3], y2[it % 3], n4, sink);
      if (w == 1) k<1><<<grid, 512>>>(x[it % 3], y1[it % 3], y2[it % 3], n4, sink);
      if (w == 2) k<2><<<grid, 512>>>(x[it % 3], y1[it % 3], y2[it % 3], n4, sink);
    };
    for (int it = 0; it < 3; ++it) go(it);
    cudaEventRecord(e0);
    for (int it = 0; it < 12; ++it) go(it);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 12;
    printf("read 1 write %d, %d CTAs/SM: %.1f us  %.0f GB/s\n", w, bps, ms * 1e3, (1 + w) * 4.0 * n / ms / 1e6);
  }
  return 0;
}
