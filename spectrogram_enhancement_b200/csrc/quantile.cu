// K4: per-column quantile threshold (quantfilt, spec_denoising/pipeline_data.py:46-49).
//
// np.quantile(src, thr, axis=0) with numpy's default 'linear' method needs the order statistics
// lo = floor((rows-1)*thr) and lo+1 of every column.  A CTA stages a [rows x 32 columns] tile in shared
// memory (row-contiguous loads), each warp fully sorts four of the columns with a register/shuffle
// bitonic network (element e = r*32 + lane, E = rows_pad/32 registers per lane), interpolates with
// numpy's float32 arithmetic (separate multiply and add, the `t >= 0.5` branch of numpy's _lerp) and the
// tile is written back thresholded.  The comparison src < q is an integer-valued output: it is
// bit-exact given the same input.
#include "kernels.h"

namespace specgpu {

constexpr int kQThreads = 256;
constexpr int kQCols = 32;
constexpr int kQPitch = kQCols + 1;

template <int E>
__device__ __forceinline__ void bitonic_sort_warp(float (&v)[E], int lane) {
  constexpr int NTOT = 32 * E;
#pragma unroll
  for (int k = 2; k <= NTOT; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < 32) {
        const bool upper = (lane & j) != 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int i = r * 32 + lane;
          const bool asc = (i & k) == 0;
          const float o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const float lo = fminf(v[r], o), hi = fmaxf(v[r], o);
          v[r] = (upper == asc) ? hi : lo;
        }
      } else {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & jr) == 0) {
            const int i = r * 32 + lane;
            const bool asc = (i & k) == 0;
            const float a = v[r], b = v[r | jr];
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            v[r] = asc ? lo : hi;
            v[r | jr] = asc ? hi : lo;
          }
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ float pick_sorted(const float (&v)[E], int idx) {
  float x = 0.f;
#pragma unroll
  for (int r = 0; r < E; ++r)
    if (r == (idx >> 5)) x = v[r];
  return __shfl_sync(0xffffffffu, x, idx & 31);
}

template <int E>
__global__ void __launch_bounds__(kQThreads) quantfilt_kernel(const float* src, int rows, int64_t cols, int64_t ld, int lo,
                                                              float g, float* dst, float* thr_out, uint8_t* mask) {
  SPECGPU_DYN_SMEM(smem);
  float* tile = reinterpret_cast<float*>(smem);               // [rows][kQPitch]
  float* s_thr = tile + (size_t)rows * kQPitch;                // [kQCols]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.y;
  const int64_t c0 = (int64_t)blockIdx.x * kQCols;
  const int ncol = (int)((cols - c0 < kQCols) ? (cols - c0) : kQCols);
  const float* sb = src + b * rows * ld;

  for (int r = warp; r < rows; r += kQThreads / 32)
    tile[r * kQPitch + lane] = (lane < ncol) ? sb[(int64_t)r * ld + c0 + lane] : 0.f;
  __syncthreads();

  for (int c = warp; c < ncol; c += kQThreads / 32) {
    float v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int i = r * 32 + lane;
      v[r] = (i < rows) ? tile[i * kQPitch + c] : INFINITY;
    }
    bitonic_sort_warp<E>(v, lane);
    const float a = pick_sorted<E>(v, lo);
    const float bb = pick_sorted<E>(v, (lo + 1 < rows) ? lo + 1 : rows - 1);
    // numpy _lerp in float32: a + (b-a)*g, replaced by b - (b-a)*(1-g) where g >= 0.5
    const float d = __fsub_rn(bb, a);
    float q;
    if (lo >= rows - 1) q = a;
    else if (g >= 0.5f) q = __fsub_rn(bb, __fmul_rn(d, __fsub_rn(1.0f, g)));
    else q = __fadd_rn(a, __fmul_rn(d, g));
    if (lane == 0) s_thr[c] = q;
  }
  __syncthreads();

  if (thr_out != nullptr && tid < ncol) thr_out[b * cols + c0 + tid] = s_thr[tid];
  const float q = (lane < ncol) ? s_thr[lane] : 0.f;
  for (int r = warp; r < rows; r += kQThreads / 32) {
    if (lane < ncol) {
      const float x = tile[r * kQPitch + lane];
      const bool below = x < q;
      const int64_t o = (b * rows + r) * ld + c0 + lane;
      if (dst != nullptr) dst[o] = below ? 0.f : x;
      if (mask != nullptr) mask[o] = below ? 0 : 1;
    }
  }
}

template <int E>
static int launch_q(const float* src, int64_t B, int rows, int64_t cols, int64_t ld, int lo, float g, float* dst,
                    float* thr_out, uint8_t* mask, cudaStream_t stream) {
  const size_t smem = ((size_t)rows * kQPitch + kQCols) * sizeof(float);
  auto kern = quantfilt_kernel<E>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  SPECGPU_LAUNCH(kern, dim3((unsigned)ceil_div(cols, kQCols), (unsigned)B), kQThreads, smem, stream, src, rows, cols, ld,
                 lo, g, dst, thr_out, mask);
  return (int)cudaGetLastError();
}

int launch_quantfilt(const float* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, int lo, float g, float* dst,
                     float* thr_out, uint8_t* mask, cudaStream_t stream) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  const int r = (int)rows;
  if (r <= 32) return launch_q<1>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  if (r <= 64) return launch_q<2>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  if (r <= 128) return launch_q<4>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  if (r <= 256) return launch_q<8>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  if (r <= 512) return launch_q<16>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  if (r <= 1024) return launch_q<32>(src, B, r, cols, ld, lo, g, dst, thr_out, mask, stream);
  return -1;
}

}  // namespace specgpu
