// K2b: all-pairs segment-averaged cross-power spectrum,
//   P[i][j][f] = mean_t conj(X_i[t][f]) X_j[t][f] * scale   (x2 on bins 1..nfreq-2),
// i.e. scipy.signal.csd(x_i, x_j, average='mean') for every channel pair -- the arithmetic the
// north-star assigns to `ae_co2` (interferometer/crosspowerspec.py:39).
//
// Input: the unscaled one-sided spectra X[C][nseg][ldf] produced by the STFT kernel
// (STFT_MODE_SPECTRA).  A CTA owns 32 consecutive frequencies (lane = frequency), a block of 8 rows i
// (warp = row) and one chunk of segments; per segment the C spectra of the 32 bins are staged in
// shared memory and every warp accumulates its row's C products in registers.  Chunk partials are
// reduced by a second, deterministic kernel that also applies the scale.
#include "kernels.h"

namespace specgpu {

constexpr int kCsdThreads = 256, kCsdRows = 8, kCsdStage = 4;  // segments staged per barrier

template <int CMAX>
__global__ void __launch_bounds__(kCsdThreads) csd_pairs_kernel(const float2* X, int C, int64_t nseg, int64_t ldf,
                                                                int nfreq, int64_t i0, int ni, int64_t seg_per_chunk,
                                                                float2* partial) {
  SPECGPU_DYN_SMEM(smem);
  float2* sx = reinterpret_cast<float2*>(smem);  // [kCsdStage][C][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f0 = blockIdx.x * 32;
  const int f = f0 + lane;
  const int chunk = blockIdx.y;
  const int rb = blockIdx.z;
  const int irow = rb * kCsdRows + warp;           // row inside [0, ni)
  const bool row_ok = irow < ni;
  const int64_t ig = i0 + irow;                    // global channel index of the row
  const int64_t t0 = (int64_t)chunk * seg_per_chunk;
  const int64_t t1 = (t0 + seg_per_chunk < nseg) ? t0 + seg_per_chunk : nseg;

  float2 acc[CMAX];
#pragma unroll
  for (int j = 0; j < CMAX; ++j) acc[j] = make_float2(0.f, 0.f);

  for (int64_t t = t0; t < t1; t += kCsdStage) {
    __syncthreads();
    for (int i = tid; i < kCsdStage * C * 32; i += kCsdThreads) {
      const int l = i & 31, c = (i >> 5) % C, s = (i >> 5) / C;
      float2 v = make_float2(0.f, 0.f);
      if (t + s < t1 && f0 + l < nfreq) v = X[((int64_t)c * nseg + t + s) * ldf + f0 + l];
      sx[i] = v;
    }
    __syncthreads();
    if (row_ok) {
#pragma unroll
      for (int s = 0; s < kCsdStage; ++s) {
        const float2 xi = sx[(s * C + (int)ig) * 32 + lane];
#pragma unroll
        for (int j = 0; j < CMAX; ++j) {
          if (j < C) {
            const float2 xj = sx[(s * C + j) * 32 + lane];
            // conj(xi) * xj
            acc[j].x += xi.x * xj.x + xi.y * xj.y;
            acc[j].y += xi.x * xj.y - xi.y * xj.x;
          }
        }
      }
    }
  }
  if (row_ok && f < nfreq) {
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
      if (j < C) partial[(((int64_t)chunk * ni + irow) * C + j) * nfreq + f] = acc[j];
  }
}

__global__ void csd_reduce_kernel(const float2* partial, int nchunk, int64_t count, int nfreq, float scale, int accumulate,
                                  float2* P) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float2 s = make_float2(0.f, 0.f);
    for (int c = 0; c < nchunk; ++c) {
      const float2 v = partial[(int64_t)c * count + i];
      s.x += v.x;
      s.y += v.y;
    }
    const int f = (int)(i % nfreq);
    const float sc = (f == 0 || f == nfreq - 1) ? scale : 2.0f * scale;
    float2 r = make_float2(s.x * sc, s.y * sc);
    if (accumulate) {            // a later block of segments of the same average (channel-sharded, chunked exchange)
      const float2 p = P[i];
      r.x += p.x;
      r.y += p.y;
    }
    P[i] = r;
  }
}

static int csd_num_chunks(int64_t nseg, int64_t ni, int nfreq) {
  const int64_t ctas = ceil_div(nfreq, 32) * ceil_div(ni, kCsdRows);
  int64_t want = ceil_div(148 * 4, ctas);
  int64_t maxc = ceil_div(nseg, 16);
  return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxc), 64));
}

size_t csd_pairs_workspace_bytes(int64_t C, int64_t ni, int nfreq, int64_t nseg) {
  return (size_t)csd_num_chunks(nseg, ni, nfreq) * ni * C * nfreq * sizeof(float2) + 256;
}

template <int CMAX>
static int launch_pairs_t(const float2* X, int C, int64_t nseg, int64_t ldf, int nfreq, int64_t i0, int ni, int nchunk,
                          float2* partial, cudaStream_t stream) {
  const size_t smem = (size_t)kCsdStage * C * 32 * sizeof(float2);
  auto kern = csd_pairs_kernel<CMAX>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t spc = ceil_div(ceil_div(nseg, nchunk), kCsdStage) * kCsdStage;
  SPECGPU_LAUNCH(kern, dim3((unsigned)ceil_div(nfreq, 32), (unsigned)nchunk, (unsigned)ceil_div(ni, kCsdRows)), kCsdThreads,
                 smem, stream, X, C, nseg, ldf, nfreq, i0, ni, spc, partial);
  return (int)cudaGetLastError();
}

int launch_csd_pairs(const float* X, int64_t C, int64_t nseg, int64_t nseg_total, int64_t ldf, int nfreq, int64_t i0,
                     int64_t ni, float scale, int accumulate, float* partial_ws, float* P, cudaStream_t stream) {
  if (C == 0 || ni == 0 || nfreq == 0) return 0;
  if (C > 64) return -1;
  const int nchunk = csd_num_chunks(nseg, ni, nfreq);
  const float2* X2 = reinterpret_cast<const float2*>(X);
  float2* part = reinterpret_cast<float2*>(partial_ws);
  int e;
  if (C <= 4) e = launch_pairs_t<4>(X2, (int)C, nseg, ldf, nfreq, i0, (int)ni, nchunk, part, stream);
  else if (C <= 8) e = launch_pairs_t<8>(X2, (int)C, nseg, ldf, nfreq, i0, (int)ni, nchunk, part, stream);
  else if (C <= 16) e = launch_pairs_t<16>(X2, (int)C, nseg, ldf, nfreq, i0, (int)ni, nchunk, part, stream);
  else if (C <= 32) e = launch_pairs_t<32>(X2, (int)C, nseg, ldf, nfreq, i0, (int)ni, nchunk, part, stream);
  else e = launch_pairs_t<64>(X2, (int)C, nseg, ldf, nfreq, i0, (int)ni, nchunk, part, stream);
  if (e) return e;
  const int64_t count = ni * C * nfreq;
  SPECGPU_LAUNCH(csd_reduce_kernel, (unsigned)std::min<int64_t>(ceil_div(count, 256), 148 * 8), 256, 0, stream,
                 (const float2*)part, nchunk, count, nfreq, scale / (float)nseg_total, accumulate, reinterpret_cast<float2*>(P));
  return (int)cudaGetLastError();
}

}  // namespace specgpu
