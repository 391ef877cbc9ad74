// K5: the cv2 image chain of the reference (spec_denoising/pipeline_data.py:52-72):
//   gaussblr  uint8-quantise -> cv2.GaussianBlur(ksize=(kw, kh), sigma 0) -> rescale        (:52-55)
//   meansub   |x - mean over time of each frequency row| -> rescale                          (:58-61)
//   morph     uint8-quantise -> CLOSE rect 4x4 -> OPEN rect 3(time) x 1 -> rescale           (:64-72)
// The uint8 intermediates are integer outputs and are reproduced bit for bit:
//   * (rescale(x)*255).astype('uint8') is evaluated in the input's own dtype with IEEE division and truncation;
//   * GaussianBlur on CV_8U is OpenCV's fixed-point path: Q8.8 kernel taps (error-diffusion rounding, built on
//     the host, see specgpu.cu), horizontal pass into Q8.8, vertical pass into Q16.16, +0.5 and >> 16,
//     BORDER_REFLECT_101;
//   * morphology on a rectangle with OpenCV's anchor (k/2, k/2) and "ignore outside" borders.
// Float outputs are float64 like numpy's (uint8 / uint8 true division, float64 means).
#include "kernels.h"

namespace specgpu {

constexpr int kImgThreads = 256;
constexpr int kImgParts = 64;     // per-image partial min/max slots

template <class T>
struct MinMax {
  T mn, mx;
};

// ---- per-image min / max in two steps (no 64-bit float atomics needed): partials, then every consumer folds them ----
template <class T>
__global__ void img_minmax_kernel(const T* src, int64_t rows, int64_t cols, int64_t ld, T* part) {
  const int64_t b = blockIdx.y;
  const int64_t total = rows * cols;
  T vmin, vmax;
  if constexpr (sizeof(T) == 1) {
    vmin = (T)255;
    vmax = (T)0;
  } else {
    vmin = (T)INFINITY;
    vmax = (T)-INFINITY;
  }
  (void)total;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {          // a CTA walks whole rows: no division per element
    const T* row = src + (b * rows + r) * ld;
    for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) {
      const T v = row[c];
      vmin = v < vmin ? v : vmin;      // NaN never wins, like np.min on finite data
      vmax = v > vmax ? v : vmax;
    }
  }
  __shared__ T s_min[kImgThreads], s_max[kImgThreads];
  s_min[threadIdx.x] = vmin;
  s_max[threadIdx.x] = vmax;
  __syncthreads();
  for (int o = kImgThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const T a = s_min[threadIdx.x + o], c = s_max[threadIdx.x + o];
      if (a < s_min[threadIdx.x]) s_min[threadIdx.x] = a;
      if (c > s_max[threadIdx.x]) s_max[threadIdx.x] = c;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[(b * kImgParts + blockIdx.x) * 2] = s_min[0];
    part[(b * kImgParts + blockIdx.x) * 2 + 1] = s_max[0];
  }
}

// Fold the kImgParts partial (min, max) pairs of image b: warp 0 reads two per lane and reduces with shuffles, the
// result is broadcast through shared memory.  Every thread of the CTA must call this (it contains a barrier).
template <class T>
__device__ __forceinline__ MinMax<T> fold_minmax(const T* part, int64_t b) {
  __shared__ double s_mm[2];     // wide enough for every T
  static_assert(kImgParts == 64, "two partials per lane");
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    T mn = part[(b * kImgParts + lane) * 2], mx = part[(b * kImgParts + lane) * 2 + 1];
    const T a = part[(b * kImgParts + 32 + lane) * 2], c = part[(b * kImgParts + 32 + lane) * 2 + 1];
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      // shuffles move 32- or 64-bit payloads: widen uint8 to int
      if constexpr (sizeof(T) == 1) {
        const int a2 = __shfl_xor_sync(0xffffffffu, (int)mn, o), c2 = __shfl_xor_sync(0xffffffffu, (int)mx, o);
        mn = (T)a2 < mn ? (T)a2 : mn;
        mx = (T)c2 > mx ? (T)c2 : mx;
      } else {
        const T a2 = __shfl_xor_sync(0xffffffffu, mn, o), c2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a2 < mn ? a2 : mn;
        mx = c2 > mx ? c2 : mx;
      }
    }
    if (lane == 0) {
      reinterpret_cast<T*>(s_mm)[0] = mn;
      reinterpret_cast<T*>(s_mm + 1)[0] = mx;
    }
  }
  __syncthreads();
  MinMax<T> m;
  m.mn = reinterpret_cast<const T*>(s_mm)[0];
  m.mx = reinterpret_cast<const T*>(s_mm + 1)[0];
  return m;
}

// (rescale(src) * 255).astype('uint8') in the dtype of src
template <class T>
__global__ void img_quantise_kernel(const T* src, int64_t rows, int64_t cols, int64_t ld, const T* part, uint8_t* dst) {
  const int64_t b = blockIdx.y;
  const int64_t r = blockIdx.x;
  const MinMax<T> m = fold_minmax(part, b);
  const T den = m.mx - m.mn;
  const T* row = src + (b * rows + r) * ld;
  uint8_t* out = dst + (b * rows + r) * cols;
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) {
    T q;
    if constexpr (sizeof(T) == 4) q = __fmul_rn(__fdiv_rn(__fsub_rn(row[c], m.mn), den), 255.0f);
    else q = __dmul_rn(__ddiv_rn(__dsub_rn(row[c], m.mn), den), 255.0);
    out[c] = (uint8_t)(int)q;          // truncation toward zero; inputs are in [0, 255]
  }
}

// (u - min) / (max - min) with numpy's uint8 arithmetic and float64 true division
__global__ void img_rescale_u8_kernel(const uint8_t* src, int64_t rows, int64_t cols, const uint8_t* part, double* dst,
                                      int64_t ldo) {
  const int64_t b = blockIdx.y;
  const int64_t r = blockIdx.x;
  const MinMax<uint8_t> m = fold_minmax(part, b);
  const double den = (double)(uint8_t)(m.mx - m.mn);
  const uint8_t* row = src + (b * rows + r) * cols;
  double* out = dst + (b * rows + r) * ldo;
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) out[c] = __ddiv_rn((double)(uint8_t)(row[c] - m.mn), den);
}

__global__ void img_rescale_f64_kernel(const double* src, int64_t rows, int64_t cols, int64_t ld, const double* part,
                                       double* dst, int64_t ldo) {
  const int64_t b = blockIdx.y;
  const int64_t r = blockIdx.x;
  const MinMax<double> m = fold_minmax(part, b);
  const double den = m.mx - m.mn;
  const double* row = src + (b * rows + r) * ld;
  double* out = dst + (b * rows + r) * ldo;
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) out[c] = __ddiv_rn(row[c] - m.mn, den);
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

// horizontal pass: inter[r][c] = sum_d src[r][reflect(c + d - kw/2)] * kx[d]   (Q8.8, <= 255 * 256)
__global__ void blur_h_kernel(const uint8_t* src, int64_t rows, int cols, const uint16_t* kx, int kw, uint16_t* inter) {
  SPECGPU_DYN_SMEM(smem);
  uint8_t* s_row = smem;                                   // [cols + kw - 1] with the border already reflected
  const int64_t row = blockIdx.x;                           // over B * rows
  const uint8_t* in = src + row * cols;
  const int half = kw / 2;
  for (int i = threadIdx.x; i < cols + kw - 1; i += blockDim.x) s_row[i] = in[reflect101(i - half, cols)];
  __syncthreads();
  uint16_t* out = inter + row * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    unsigned acc = 0;
    for (int d = 0; d < kw; ++d) acc += (unsigned)s_row[c + d] * (unsigned)kx[d];
    out[c] = (uint16_t)(acc > 65535u ? 65535u : acc);      // ufixedpoint16 saturates (cannot trigger: taps sum to 256)
  }
}

// vertical pass: dst[r][c] = (sum_e inter[reflect(r + e - kh/2)][c] * ky[e] + 2^15) >> 16
__global__ void blur_v_kernel(const uint16_t* inter, int rows, int cols, const uint16_t* ky, int kh, uint8_t* dst) {
  const int64_t b = blockIdx.z;
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const uint16_t* img = inter + b * (int64_t)rows * cols;
  unsigned acc = 0;
  for (int e = 0; e < kh; ++e) acc += (unsigned)img[(int64_t)reflect101(r + e - kh / 2, rows) * cols + c] * (unsigned)ky[e];
  const unsigned v = (acc + 32768u) >> 16;
  dst[(b * rows + r) * cols + c] = (uint8_t)(v > 255u ? 255u : v);
}

// rectangular dilation / erosion, anchor (ay, ax), pixels outside the image are ignored
__global__ void morph_rect_kernel(const uint8_t* src, int rows, int cols, int kh, int kw, int ay, int ax, int erode,
                                  uint8_t* dst) {
  const int64_t b = blockIdx.z;
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const uint8_t* img = src + b * (int64_t)rows * cols;
  int v = erode ? 255 : 0;
  for (int dy = 0; dy < kh; ++dy) {
    const int y = r + dy - ay;
    if (y < 0 || y >= rows) continue;
    for (int dx = 0; dx < kw; ++dx) {
      const int x = c + dx - ax;
      if (x < 0 || x >= cols) continue;
      const int p = img[(int64_t)y * cols + x];
      v = erode ? (p < v ? p : v) : (p > v ? p : v);
    }
  }
  dst[(b * rows + r) * cols + c] = (uint8_t)v;
}

// |x - mean over the row| (np.mean(axis=1) in float64), one CTA per row; fixed-order tree => deterministic
__global__ void meansub_abs_kernel(const double* src, int64_t rows, int64_t cols, int64_t ld, double* dst) {
  const int64_t row = blockIdx.x;       // over B * rows
  const double* in = src + row * ld;
  double s = 0.0;
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) s += in[c];
  __shared__ double sh[kImgThreads];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = kImgThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double mean = __ddiv_rn(sh[0], (double)cols);
  double* out = dst + row * cols;
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) out[c] = fabs(in[c] - mean);
}

// ---- launchers ----------------------------------------------------------------------------------------------------
size_t imgchain_workspace_bytes(int64_t B, int64_t rows, int64_t cols) {
  const size_t px = (size_t)B * rows * cols;
  // two uint8 planes, one uint16 plane, one float64 plane, partial min/max slots, kernel taps
  return 2 * (px + 256) + 2 * px + 256 + 8 * px + 256 + (size_t)B * kImgParts * 2 * 8 + 256 + 4096;
}

namespace {
struct ImgWs {
  uint8_t *u8a, *u8b;
  uint16_t* u16;
  double* f64;
  void* part;
  uint16_t* taps;
};
ImgWs img_carve(void* ws, int64_t B, int64_t rows, int64_t cols) {
  const size_t px = (size_t)B * rows * cols;
  char* p = static_cast<char*>(ws);
  auto take = [&](size_t bytes) {
    char* q = p;
    p += (bytes + 255) & ~(size_t)255;
    return q;
  };
  ImgWs w;
  w.f64 = reinterpret_cast<double*>(take(8 * px));
  w.u16 = reinterpret_cast<uint16_t*>(take(2 * px));
  w.u8a = reinterpret_cast<uint8_t*>(take(px));
  w.u8b = reinterpret_cast<uint8_t*>(take(px));
  w.part = take((size_t)B * kImgParts * 2 * 8);
  w.taps = reinterpret_cast<uint16_t*>(take(4096));
  return w;
}
template <class T>
void run_minmax(const T* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, T* part, cudaStream_t st) {
  SPECGPU_LAUNCH((img_minmax_kernel<T>), dim3(kImgParts, (unsigned)B), kImgThreads, 0, st, src, rows, cols, ld, part);
}
}  // namespace

// taps_host: kw + kh Q8.8 taps (kx then ky), built by the caller.
int launch_gaussblr(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, const uint16_t* taps_host,
                    int kw, int kh, void* ws, double* dst, int64_t ldo, uint8_t* u8_out, cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  cudaError_t e = cudaMemcpyAsync(w.taps, taps_host, (size_t)(kw + kh) * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  const dim3 rowgrid((unsigned)rows, (unsigned)B);
  if (in_f64) {
    run_minmax<double>((const double*)src, B, rows, cols, ld, (double*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<double>), rowgrid, kImgThreads, 0, st, (const double*)src, rows, cols, ld,
                   (const double*)w.part, w.u8a);
  } else {
    run_minmax<float>((const float*)src, B, rows, cols, ld, (float*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<float>), rowgrid, kImgThreads, 0, st, (const float*)src, rows, cols, ld,
                   (const float*)w.part, w.u8a);
  }
  SPECGPU_LAUNCH(blur_h_kernel, (unsigned)(B * rows), kImgThreads, (size_t)(cols + kw), st, (const uint8_t*)w.u8a, rows,
                 (int)cols, (const uint16_t*)w.taps, kw, w.u16);
  const dim3 pixgrid((unsigned)ceil_div(cols, kImgThreads), (unsigned)rows, (unsigned)B);
  uint8_t* blurred = u8_out ? u8_out : w.u8b;
  SPECGPU_LAUNCH(blur_v_kernel, pixgrid, kImgThreads, 0, st, (const uint16_t*)w.u16, (int)rows, (int)cols,
                 (const uint16_t*)(w.taps + kw), kh, blurred);
  run_minmax<uint8_t>(blurred, B, rows, cols, cols, (uint8_t*)w.part, st);
  SPECGPU_LAUNCH(img_rescale_u8_kernel, rowgrid, kImgThreads, 0, st, (const uint8_t*)blurred, rows, cols,
                 (const uint8_t*)w.part, dst, ldo);
  return (int)cudaGetLastError();
}

int launch_meansub(const double* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst, int64_t ldo,
                   cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  SPECGPU_LAUNCH(meansub_abs_kernel, (unsigned)(B * rows), kImgThreads, 0, st, src, rows, cols, ld, w.f64);
  run_minmax<double>(w.f64, B, rows, cols, cols, (double*)w.part, st);
  SPECGPU_LAUNCH(img_rescale_f64_kernel, dim3((unsigned)rows, (unsigned)B), kImgThreads, 0, st, (const double*)w.f64, rows,
                 cols, cols, (const double*)w.part, dst, ldo);
  return (int)cudaGetLastError();
}

int launch_morph(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst,
                 int64_t ldo, uint8_t* u8_out, cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  const dim3 rowgrid((unsigned)rows, (unsigned)B);
  if (in_f64) {
    run_minmax<double>((const double*)src, B, rows, cols, ld, (double*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<double>), rowgrid, kImgThreads, 0, st, (const double*)src, rows, cols, ld,
                   (const double*)w.part, w.u8a);
  } else {
    run_minmax<float>((const float*)src, B, rows, cols, ld, (float*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<float>), rowgrid, kImgThreads, 0, st, (const float*)src, rows, cols, ld,
                   (const float*)w.part, w.u8a);
  }
  const dim3 pixgrid((unsigned)ceil_div(cols, kImgThreads), (unsigned)rows, (unsigned)B);
  const int R = (int)rows, Cc = (int)cols;
  uint8_t* fin = u8_out ? u8_out : w.u8a;
  // MORPH_CLOSE with rect(4, 4): dilate then erode, anchor (2, 2)
  SPECGPU_LAUNCH(morph_rect_kernel, pixgrid, kImgThreads, 0, st, (const uint8_t*)w.u8a, R, Cc, 4, 4, 2, 2, 0, w.u8b);
  SPECGPU_LAUNCH(morph_rect_kernel, pixgrid, kImgThreads, 0, st, (const uint8_t*)w.u8b, R, Cc, 4, 4, 2, 2, 1, w.u8a);
  // MORPH_OPEN with getStructuringElement(RECT, (3, 1)) = 1 row x 3 columns: erode then dilate, anchor (0, 1)
  SPECGPU_LAUNCH(morph_rect_kernel, pixgrid, kImgThreads, 0, st, (const uint8_t*)w.u8a, R, Cc, 1, 3, 0, 1, 1, w.u8b);
  SPECGPU_LAUNCH(morph_rect_kernel, pixgrid, kImgThreads, 0, st, (const uint8_t*)w.u8b, R, Cc, 1, 3, 0, 1, 0, fin);
  run_minmax<uint8_t>(fin, B, rows, cols, cols, (uint8_t*)w.part, st);
  SPECGPU_LAUNCH(img_rescale_u8_kernel, rowgrid, kImgThreads, 0, st, (const uint8_t*)fin, rows, cols, (const uint8_t*)w.part,
                 dst, ldo);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
