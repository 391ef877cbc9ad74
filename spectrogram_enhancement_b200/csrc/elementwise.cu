// Array helpers of the reference: rescale / norm (spec_denoising/pipeline_data.py:38-44) and the VAE
// tile cut patch / unpatch (VAE/manual_scan.py:28-49).  All are single streaming passes over HBM.
#include "kernels.h"

namespace specgpu {

constexpr int kEwThreads = 256;

// ---- global min / max per batch item ---------------------------------------------------------------
__global__ void minmax_reduce_kernel(const float* src, int64_t rows, int64_t cols, int64_t ld, unsigned* mm) {
  const int64_t b = blockIdx.y;
  const int64_t total = rows * cols;
  float vmin = INFINITY, vmax = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float v = src[(b * rows + r) * ld + c];
    vmin = fminf(vmin, v);
    vmax = fmaxf(vmax, v);
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  __shared__ float s_min[kEwThreads / 32], s_max[kEwThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_min[warp] = vmin;
    s_max[warp] = vmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kEwThreads / 32; ++w) {
      vmin = fminf(vmin, s_min[w]);
      vmax = fmaxf(vmax, s_max[w]);
    }
    atomicMin(mm + 2 * b, float_to_ordered(vmin));
    atomicMax(mm + 2 * b + 1, float_to_ordered(vmax));
  }
}

__global__ void rescale_apply_kernel(const float* src, int64_t rows, int64_t cols, int64_t ld, const unsigned* mm,
                                     float* dst) {
  const int64_t b = blockIdx.y;
  const float mn = ordered_to_float(mm[2 * b]);
  const float den = ordered_to_float(mm[2 * b + 1]) - mn;
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t o = (b * rows + r) * ld + c;
    dst[o] = __fdiv_rn(src[o] - mn, den);
  }
}

// ---- mean / population std per batch item (double accumulation) --------------------------------------
__global__ void moments_kernel(const float* src, int64_t rows, int64_t cols, int64_t ld, double* sums) {
  const int64_t b = blockIdx.y;
  const int64_t total = rows * cols;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const double v = (double)src[(b * rows + r) * ld + c];
    s1 += v;
    s2 += v * v;
  }
  __shared__ double sh1[kEwThreads], sh2[kEwThreads];
  sh1[threadIdx.x] = s1;
  sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = kEwThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh1[threadIdx.x] += sh1[threadIdx.x + o];
      sh2[threadIdx.x] += sh2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // one slot per block: summed (in block order, deterministically) by norm_apply_kernel
    sums[(b * gridDim.x + blockIdx.x) * 2] = sh1[0];
    sums[(b * gridDim.x + blockIdx.x) * 2 + 1] = sh2[0];
  }
}

__global__ void norm_apply_kernel(const float* src, int64_t rows, int64_t cols, int64_t ld, const double* sums,
                                  int nparts, float* dst) {
  const int64_t b = blockIdx.y;
  double s1 = 0.0, s2 = 0.0;
  for (int p = 0; p < nparts; ++p) {
    s1 += sums[(b * nparts + p) * 2];
    s2 += sums[(b * nparts + p) * 2 + 1];
  }
  const int64_t total = rows * cols;
  const double mean = s1 / (double)total;
  double var = s2 / (double)total - mean * mean;
  if (var < 0.0) var = 0.0;
  const float mn = (float)mean;
  const float sd = (float)sqrt(var);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t o = (b * rows + r) * ld + c;
    dst[o] = __fdiv_rn(src[o] - mn, sd);
  }
}

// ---- VAE tiles ---------------------------------------------------------------------------------------
// out[(i*ntiles + x)][r][c] = src[i][r][x*tile_w + c].  One CTA per (image i, row r): the row is read once,
// contiguously, and leaves as ntiles runs of tile_w contiguous elements.
template <class OutT>
__global__ void patch_kernel(const float* src, int64_t rows, int64_t ld, int tile_w, int ntiles, OutT* out) {
  const int64_t i = blockIdx.y;
  const int64_t r = blockIdx.x;
  const float* row = src + (i * rows + r) * ld;
  OutT* obase = out + (i * ntiles * rows + r) * tile_w;           // tile x of this row starts at obase + x*rows*tile_w
  const int64_t tstride = rows * tile_w;
  const unsigned width = (unsigned)tile_w * (unsigned)ntiles;
  for (unsigned col = threadIdx.x; col < width; col += blockDim.x) {
    const unsigned x = col / (unsigned)tile_w, c = col - x * (unsigned)tile_w;
    obase[x * tstride + c] = (OutT)__ldg(row + col);
  }
}

// dst[i][r][x*tile_w + c] = tiles[(i*ntiles + x)][r][c]
template <class InT, class OutT>
__global__ void unpatch_kernel(const InT* tiles, int64_t rows, int tile_w, int ntiles, OutT* dst, int64_t ld) {
  const int64_t i = blockIdx.y;
  const int64_t r = blockIdx.x;
  OutT* row = dst + (i * rows + r) * ld;
  const InT* ibase = tiles + (i * ntiles * rows + r) * tile_w;
  const int64_t tstride = rows * tile_w;
  const unsigned width = (unsigned)tile_w * (unsigned)ntiles;
  for (unsigned col = threadIdx.x; col < width; col += blockDim.x) {
    const unsigned x = col / (unsigned)tile_w, c = col - x * (unsigned)tile_w;
    row[col] = (OutT)ibase[x * tstride + c];
  }
}

// hacked = svd.copy(); hacked[hacked < 0] = 0   (denoising_by_svd.ipynb:280-281); NaN stays NaN
__global__ void clip_neg_kernel(const float* src, int64_t n, float* dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    dst[i] = (v < 0.f) ? 0.f : v;
  }
}

static unsigned stream_blocks(int64_t total) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, kEwThreads * 4), 148 * 8));
}

// mm[b] = {0xffffffff, 0}: identity of the ordered-uint (min, max) pair of rescale()
__global__ void rescale_minmax_init_kernel(unsigned* mm, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    mm[2 * i] = 0xffffffffu;
    mm[2 * i + 1] = 0u;
  }
}

int launch_rescale(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst, unsigned* mm_ws,
                   cudaStream_t stream) {
  if (B == 0 || rows * cols == 0) return 0;
  SPECGPU_LAUNCH(rescale_minmax_init_kernel, (unsigned)ceil_div(B, 128), 128, 0, stream, mm_ws, B);
  const unsigned gx = stream_blocks(rows * cols);
  SPECGPU_LAUNCH(minmax_reduce_kernel, dim3(gx, (unsigned)B), kEwThreads, 0, stream, src, rows, cols, ld, mm_ws);
  SPECGPU_LAUNCH(rescale_apply_kernel, dim3(gx, (unsigned)B), kEwThreads, 0, stream, src, rows, cols, ld,
                 (const unsigned*)mm_ws, dst);
  return (int)cudaGetLastError();
}

int launch_clip_neg(const float* src, int64_t n, float* dst, cudaStream_t stream) {
  if (n == 0) return 0;
  SPECGPU_LAUNCH(clip_neg_kernel, stream_blocks(n), kEwThreads, 0, stream, src, n, dst);
  return (int)cudaGetLastError();
}

int norm_num_parts(int64_t rows, int64_t cols) { return (int)std::min<unsigned>(stream_blocks(rows * cols), 256u); }

int launch_norm(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* dst, double* sums_ws,
                cudaStream_t stream) {
  if (B == 0 || rows * cols == 0) return 0;
  const unsigned parts = (unsigned)norm_num_parts(rows, cols);
  SPECGPU_LAUNCH(moments_kernel, dim3(parts, (unsigned)B), kEwThreads, 0, stream, src, rows, cols, ld, sums_ws);
  SPECGPU_LAUNCH(norm_apply_kernel, dim3(stream_blocks(rows * cols), (unsigned)B), kEwThreads, 0, stream, src, rows, cols,
                 ld, (const double*)sums_ws, (int)parts, dst);
  return (int)cudaGetLastError();
}

int launch_patch(const float* src, int64_t n, int64_t rows, int64_t ld, int tile_w, int ntiles, void* out, int out_f64,
                 cudaStream_t stream) {
  if (n == 0 || rows == 0 || ntiles == 0) return 0;
  const dim3 grid((unsigned)rows, (unsigned)n);
  if (out_f64) SPECGPU_LAUNCH(patch_kernel<double>, grid, kEwThreads, 0, stream, src, rows, ld, tile_w, ntiles, (double*)out);
  else SPECGPU_LAUNCH(patch_kernel<float>, grid, kEwThreads, 0, stream, src, rows, ld, tile_w, ntiles, (float*)out);
  return (int)cudaGetLastError();
}

int launch_unpatch(const void* tiles, int in_f64, int64_t n, int64_t rows, int tile_w, int ntiles, void* dst,
                   int out_f64, int64_t ld, cudaStream_t stream) {
  if (n == 0 || rows == 0 || ntiles == 0) return 0;
  const dim3 grid((unsigned)rows, (unsigned)n);
  if (in_f64 && out_f64)
    SPECGPU_LAUNCH((unpatch_kernel<double, double>), grid, kEwThreads, 0, stream, (const double*)tiles, rows, tile_w, ntiles, (double*)dst, ld);
  else if (in_f64)
    SPECGPU_LAUNCH((unpatch_kernel<double, float>), grid, kEwThreads, 0, stream, (const double*)tiles, rows, tile_w, ntiles, (float*)dst, ld);
  else if (out_f64)
    SPECGPU_LAUNCH((unpatch_kernel<float, double>), grid, kEwThreads, 0, stream, (const float*)tiles, rows, tile_w, ntiles, (double*)dst, ld);
  else
    SPECGPU_LAUNCH((unpatch_kernel<float, float>), grid, kEwThreads, 0, stream, (const float*)tiles, rows, tile_w, ntiles, (float*)dst, ld);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
