// Micro-benchmark: read bandwidth of the Gram producer's access pattern -- [256 rows x 32 floats] slabs of a
// [B][256][ld] image -- under different CTA -> chunk mappings.  Not part of the library.
//   mode 0: CTA i owns the contiguous global chunk range [i*per, (i+1)*per)          (gram_tc_kernel today)
//   mode 1: the CTAs that share a matrix interleave its chunks (c = k, k+ns, ...)
//   mode 2: grid-stride over all chunks (adjacent CTAs read adjacent chunks; not usable for a Gram accumulation)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(const float* S, int B, int nchunk, long ld, int cols, int mode, int per, int ns, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long total = (long)B * nchunk;
  float acc = 0.f;
  auto slab = [&](long g) {
    const long b = g / nchunk; const int c = (int)(g - b * nchunk);
    const int kcol = c * 32 + lane;
    const float* p = S + (b * 256 + warp) * ld + kcol;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = kcol < cols ? __ldg(p + (long)i * 16 * ld) : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += v[i];
  };
  if (mode == 0) {
    const long g0 = (long)blockIdx.x * per, g1 = min(g0 + per, total);
    for (long g = g0; g < g1; ++g) slab(g);
  } else if (mode == 1) {
    const int b = blockIdx.x / ns, kk = blockIdx.x % ns;
    if (b < B) for (int c = kk; c < nchunk; c += ns) slab((long)b * nchunk + c);
  } else {
    for (long g = blockIdx.x; g < total; g += gridDim.x) slab(g);
  }
  if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv) {
  const int B = 40, rows = 256, cols = 3905; const long ld = argc > 1 ? atol(argv[1]) : 3905; const int nchunk = (cols + 31) / 32;
  float *S[3], *out; for (int i = 0; i < 3; ++i) { cudaMalloc(&S[i], sizeof(float) * B * rows * ld); cudaMemset(S[i], 0, sizeof(float) * B * rows * ld); }
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) for (int ns = 3; ns <= 4; ++ns) {
    if (mode != 1 && ns == 4) continue;
    const int per = (B * nchunk + 147) / 148;
    const int grid = mode == 0 ? (B * nchunk + per - 1) / per : (mode == 1 ? B * ns : 148);
    for (int it = 0; it < 3; ++it) k<<<grid, 512>>>(S[it % 3], B, nchunk, ld, cols, mode, per, ns, out);
    cudaEventRecord(e0);
    for (int it = 0; it < 12; ++it) k<<<grid, 512>>>(S[it % 3], B, nchunk, ld, cols, mode, per, ns, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 12;
    printf("ld %ld mode %d ns %d grid %d: %.1f us  %.0f GB/s\n", ld, mode, ns, grid, ms * 1e3, 4.0 * B * rows * cols / ms / 1e6);
  }
  return 0;
}
