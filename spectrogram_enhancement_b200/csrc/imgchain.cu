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
//
// Layout: the uint8 planes between the stages live in the workspace with rows padded to 16 bytes, so every stage moves
// 4 pixels per 32-bit word.  The blur is ONE kernel per tile (horizontal Q8.8 pass with IDP.4A on packed taps into
// shared memory, vertical pass out of it), the four morphology passes are ONE kernel per tile (separable min / max
// on 16-bit lanes in shared-memory planes, halo 4 + 2 rows and 8 + 8 columns); both fold the image min / max of their uint8
// result with integer atomics, and the final rescale is a 256-entry float64 table per CTA (uint8 has 256 quotients).
#include <type_traits>
#include <utility>
#include <vector>

#include "kernels.h"

namespace specgpu {

constexpr int kImgThreads = 256;   // the tile kernels (blur, morphology)
#if defined(SPECGPU_EMULATE)
constexpr int kRowThreads = 64;    // the CPU emulation pays per thread and barrier: narrower row CTAs, same code
#else
constexpr int kRowThreads = 256;   // the row-per-CTA kernels (min/max, quantise, rescale, meansub)
#endif
constexpr int kImgParts = 64;      // per-image partial min/max slots

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
    for (unsigned c0 = threadIdx.x; c0 < (unsigned)cols; c0 += 4 * kRowThreads) {     // four loads in flight per thread
      T v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned c = c0 + k * kRowThreads;
        v[k] = row[c < (unsigned)cols ? c : c0];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        vmin = v[k] < vmin ? v[k] : vmin;      // NaN never wins, like np.min on finite data
        vmax = v[k] > vmax ? v[k] : vmax;
      }
    }
  }
  __shared__ T s_min[kRowThreads], s_max[kRowThreads];
  s_min[threadIdx.x] = vmin;
  s_max[threadIdx.x] = vmax;
  __syncthreads();
  for (int o = kRowThreads / 2; o > 0; o >>= 1) {
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
  // images with fewer rows than slots launch fewer CTAs: the unused slots hold the identity
  if (blockIdx.x == 0)
    for (int q = gridDim.x + threadIdx.x; q < kImgParts; q += blockDim.x) {
      part[(b * kImgParts + q) * 2] = sizeof(T) == 1 ? (T)255 : (T)INFINITY;
      part[(b * kImgParts + q) * 2 + 1] = sizeof(T) == 1 ? (T)0 : (T)-INFINITY;
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

// t / den for many t and one den, correctly rounded: q = RN(t * y), r = t - q * den (exact in the FMA), q' = RN(q + r * y)
// with y = RN(1 / den) is the IEEE quotient whenever den is normal and its significand is not all ones (Markstein);
// `ok` is false otherwise and the caller divides.  Quotients so small that the residual underflows truncate to 0 in the
// uint8 quantisation either way.
struct FastDivF {
  float den, y;
  bool ok;
  __device__ __forceinline__ FastDivF(float d) : den(d) {
    y = __fdiv_rn(1.0f, d);
    const uint32_t u = __float_as_uint(d), e = (u >> 23) & 255u;
    ok = e > 1u && e < 254u && (u & 0x7fffffu) != 0x7fffffu;
  }
  __device__ __forceinline__ float div(float t) const {
    if (!ok) return __fdiv_rn(t, den);
    const float q = __fmul_rn(t, y);
    return __fmaf_rn(__fmaf_rn(-q, den, t), y, q);
  }
};
struct FastDivD {
  double den, y;
  bool ok;
  __device__ __forceinline__ FastDivD(double d) : den(d) {
    y = __ddiv_rn(1.0, d);
    const uint64_t u = (uint64_t)__double_as_longlong(d), e = (u >> 52) & 2047u;
    ok = e > 1u && e < 2046u && (u & 0xfffffffffffffull) != 0xfffffffffffffull;
  }
  __device__ __forceinline__ double div(double t) const {
    if (!ok) return __ddiv_rn(t, den);
    const double q = __dmul_rn(t, y);
    return __fma_rn(__fma_rn(-q, den, t), y, q);
  }
};

// pitch (bytes) of the workspace uint8 planes
static inline int64_t img_pitch(int64_t cols) { return (cols + 15) & ~(int64_t)15; }

// (rescale(src) * 255).astype('uint8') in the dtype of src, into a pitched plane.  Block (0, b) also resets the
// {min, max} slot the uint8 producers of image b fold into.
template <class T>
__global__ void img_quantise_kernel(const T* src, int64_t rows, int64_t cols, int64_t ld, const T* part, uint8_t* dst,
                                    int64_t pitch, unsigned* mm8, int rpc) {
  const int64_t b = blockIdx.y;
  const MinMax<T> m = fold_minmax(part, b);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    mm8[2 * b] = 255u;
    mm8[2 * b + 1] = 0u;
  }
  const T den = m.mx - m.mn;
  typename std::conditional<sizeof(T) == 4, FastDivF, FastDivD>::type fd(den);
  const int64_t rend = min((int64_t)(blockIdx.x + 1) * rpc, rows);
  for (int64_t r = (int64_t)blockIdx.x * rpc; r < rend; ++r) {      // rpc rows per CTA amortise the fold above
    const T* row = src + (b * rows + r) * ld;
    uint8_t* out = dst + (b * rows + r) * pitch;
    for (unsigned c0 = threadIdx.x; c0 < (unsigned)cols; c0 += 4 * kRowThreads) {
      T v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned c = c0 + k * kRowThreads;
        v[k] = row[c < (unsigned)cols ? c : c0];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned c = c0 + k * kRowThreads;
        T q;
        if constexpr (sizeof(T) == 4) q = __fmul_rn(fd.div(__fsub_rn(v[k], m.mn)), 255.0f);
        else q = __dmul_rn(fd.div(__dsub_rn(v[k], m.mn)), 255.0);
        if (c < (unsigned)cols) out[c] = (uint8_t)(int)q;          // truncation toward zero; inputs are in [0, 255]
      }
    }
  }
}

// (u - min) / (max - min) with numpy's uint8 arithmetic and float64 true division: 256 possible quotients per image
__global__ void img_rescale_u8_kernel(const uint8_t* src, int64_t rows, int64_t cols, int64_t pitch, const unsigned* mm8,
                                      double* dst, int64_t ldo, int rpc) {
  __shared__ double s_lut[256];
  const int64_t b = blockIdx.y;
  const uint8_t mn = (uint8_t)mm8[2 * b], mx = (uint8_t)mm8[2 * b + 1];
  const double den = (double)(uint8_t)(mx - mn);
  for (int v = threadIdx.x; v < 256; v += blockDim.x) s_lut[v] = __ddiv_rn((double)(uint8_t)((uint8_t)v - mn), den);
  __syncthreads();
  const int64_t rend = min((int64_t)(blockIdx.x + 1) * rpc, rows);
  for (int64_t r = (int64_t)blockIdx.x * rpc; r < rend; ++r) {
    const uint8_t* row = src + (b * rows + r) * pitch;
    double* out = dst + (b * rows + r) * ldo;
    for (unsigned c0 = threadIdx.x; c0 < (unsigned)cols; c0 += 4 * kRowThreads) {
      uint8_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned c = c0 + k * kRowThreads;
        v[k] = row[c < (unsigned)cols ? c : c0];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned c = c0 + k * kRowThreads;
        if (c < (unsigned)cols) out[c] = s_lut[v[k]];
      }
    }
  }
}

// pitched plane -> dense [B][rows][cols] (only when the caller asks for the uint8 image)
__global__ void img_unpitch_kernel(const uint8_t* src, int64_t cols, int64_t pitch, uint8_t* dst) {
  const int64_t row = blockIdx.x;       // over B * rows
  for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) dst[row * cols + c] = src[row * pitch + c];
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

// fold the valid bytes of a packed word into running min / max
__device__ __forceinline__ void minmax_bytes(uint32_t w, int nvalid, unsigned& mn, unsigned& mx) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const unsigned v = (w >> (8 * s)) & 255u;
    if (s < nvalid) {
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
  }
}

// CTA-wide fold of per-thread {min, max} into the image slot: warp shuffles, shared atomics, one global atomic pair
__device__ __forceinline__ void fold_mm8(unsigned mn, unsigned mx, unsigned* slot) {
  __shared__ unsigned s_mm8[2];
  if (threadIdx.x == 0) {
    s_mm8[0] = 255u;
    s_mm8[1] = 0u;
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned a = __shfl_xor_sync(0xffffffffu, mn, o), c = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_mm8[0], mn);
    atomicMax(&s_mm8[1], mx);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMin(&slot[0], s_mm8[0]);
    atomicMax(&slot[1], s_mm8[1]);
  }
}

// ---- GaussianBlur, both passes in one kernel -------------------------------------------------------------------------
// Tile: kBlurRows x TC output pixels.  Shared memory holds the uint8 tile with its reflected border (kh - 1 extra rows,
// kw - 1 extra columns; byte i of a row is image column c0 - kw/2 + i, so the window of output column c starts at byte
// c and 4-pixel groups start on a word), then the Q8.8 horizontal result of all its rows.
//   horizontal: inter[r][c] = sum_d src[r][c + d - kw/2] * kx[d]        (<= 255 * 256)
//   vertical:   dst[r][c]   = (sum_e inter[r + e - kh/2][c] * ky[e] + 2^15) >> 16
// The taps are Q8.8 <= 256; when all horizontal ones fit a byte (always, except kw = 1) four of them are packed per
// word and the pass runs on IDP.4A with funnel-shifted windows: 4 pixels x 4 taps per 4 + 3 instructions.
constexpr int kBlurRows = 16;

struct BlurGeom {
  int tc;                // tile columns (multiple of 4)
  int in_words;          // words per input-tile row
  size_t smem;
};

template <int GT>   // packed tap words known at compile time (0: run-time count)
__global__ void __launch_bounds__(kImgThreads) blur_fused_kernel(const uint8_t* src, int rows, int cols, int64_t pitch,
                                                                 const uint16_t* taps, int kw, int kh, int tc, int in_words,
                                                                 int packed, uint8_t* dst, unsigned* mm8) {
  SPECGPU_DYN_SMEM(smem);
  const int rin = kBlurRows + kh - 1;                                        // input rows of the tile
  uint32_t* s_in = reinterpret_cast<uint32_t*>(smem);                        // [rin][in_words]
  uint32_t* s_h = s_in + rin * in_words;                                     // [rin][tc / 2]   (two Q8.8 values per word)
  uint32_t* s_tap = s_h + rin * (tc / 2);                                    // packed kx words, then kx, ky as uint16 pairs
  const int G = (kw + 3) / 4;
  uint16_t* s_k16 = reinterpret_cast<uint16_t*>(s_tap + G);                  // [kw + kh]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.z;
  const int r0 = blockIdx.y * kBlurRows, c0 = blockIdx.x * tc;
  const int halfw = kw / 2, halfh = kh / 2;
  const uint8_t* img = src + b * (int64_t)rows * pitch;

  for (int i = tid; i < kw + kh; i += kImgThreads) s_k16[i] = taps[i];
  for (int g = tid; g < G; g += kImgThreads) {
    uint32_t w = 0;
    for (int s = 0; s < 4; ++s)
      if (4 * g + s < kw) w |= (uint32_t)(taps[4 * g + s] & 255u) << (8 * s);
    s_tap[g] = w;
  }
  // ---- load the tile: aligned word pairs + funnel shift in the interior, per-byte reflection at the image border ----
  const int x0 = c0 - halfw;                                                 // image column of byte 0 of a tile row
  const int sh = (x0 & 3) * 8;
  const int lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < rin; j += kImgThreads / 32) {                       // a warp per tile row: no index division
    const uint8_t* row = img + (int64_t)reflect101(r0 - halfh + j, rows) * pitch;
    for (int k = lane; k < in_words; k += 32) {
      const int x = x0 + 4 * k;
      uint32_t w;
      if (x >= 0 && x + 3 < cols) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(row) + (x >> 2);
        const uint32_t lo = rw[0];
        w = sh ? __funnelshift_r(lo, rw[1], sh) : lo;
      } else {
        w = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) w |= (uint32_t)row[reflect101(x + s, cols)] << (8 * s);
      }
      s_in[j * in_words + k] = w;
    }
  }
  __syncthreads();
  // ---- horizontal pass: 4 outputs per item (tc / 4 is a power of two) ----
  const int tq = tc / 4, lq = 31 - __clz(tq);
  uint2* s_h2 = reinterpret_cast<uint2*>(s_h);                               // four Q8.8 values per 64-bit slot
  for (int i = tid; i < rin * tq; i += kImgThreads) {
    const int j = i >> lq, g = i & (tq - 1);
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (GT > 0) {
      const uint32_t* rw = s_in + j * in_words + g;
      uint32_t w0 = rw[0], w1 = rw[1];
#pragma unroll
      for (int t = 0; t < GT; ++t) {
        const uint32_t tap = s_tap[t];
        a0 = __dp4a(w0, tap, a0);
        a1 = __dp4a(__funnelshift_r(w0, w1, 8), tap, a1);
        a2 = __dp4a(__funnelshift_r(w0, w1, 16), tap, a2);
        a3 = __dp4a(__funnelshift_r(w0, w1, 24), tap, a3);
        w0 = w1;
        w1 = rw[t + 2];
      }
    } else if (packed) {
      const uint32_t* rw = s_in + j * in_words + g;
      uint32_t w0 = rw[0], w1 = rw[1];
      for (int t = 0; t < G; ++t) {
        const uint32_t tap = s_tap[t];
        a0 = __dp4a(w0, tap, a0);
        a1 = __dp4a(__funnelshift_r(w0, w1, 8), tap, a1);
        a2 = __dp4a(__funnelshift_r(w0, w1, 16), tap, a2);
        a3 = __dp4a(__funnelshift_r(w0, w1, 24), tap, a3);
        w0 = w1;
        w1 = rw[t + 2];
      }
    } else {
      const uint8_t* rb = reinterpret_cast<const uint8_t*>(s_in + j * in_words) + 4 * g;
      for (int d = 0; d < kw; ++d) {
        const unsigned k = s_k16[d];
        a0 += rb[d] * k;
        a1 += rb[d + 1] * k;
        a2 += rb[d + 2] * k;
        a3 += rb[d + 3] * k;
      }
    }
    a0 = a0 > 65535u ? 65535u : a0;        // ufixedpoint16 saturates (cannot trigger: the taps sum to 256)
    a1 = a1 > 65535u ? 65535u : a1;
    a2 = a2 > 65535u ? 65535u : a2;
    a3 = a3 > 65535u ? 65535u : a3;
    s_h2[i] = make_uint2(a0 | (a1 << 16), a2 | (a3 << 16));
  }
  __syncthreads();
  // ---- vertical pass, store, min / max ----
  unsigned mn = 255u, mx = 0u;
  const uint16_t* ky = s_k16 + kw;
  for (int i = tid; i < kBlurRows * tq; i += kImgThreads) {
    const int j = i >> lq, g = i & (tq - 1);
    const int r = r0 + j, c = c0 + 4 * g;
    if (r >= rows || c >= cols) continue;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int e = 0; e < kh; ++e) {
      const unsigned k = ky[e];
      const uint2 pq = s_h2[i + e * tq];
      a0 += (pq.x & 65535u) * k;
      a1 += (pq.x >> 16) * k;
      a2 += (pq.y & 65535u) * k;
      a3 += (pq.y >> 16) * k;
    }
    a0 = (a0 + 32768u) >> 16;
    a1 = (a1 + 32768u) >> 16;
    a2 = (a2 + 32768u) >> 16;
    a3 = (a3 + 32768u) >> 16;
    a0 = a0 > 255u ? 255u : a0;
    a1 = a1 > 255u ? 255u : a1;
    a2 = a2 > 255u ? 255u : a2;
    a3 = a3 > 255u ? 255u : a3;
    const uint32_t w = a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
    *reinterpret_cast<uint32_t*>(dst + (b * rows + r) * pitch + c) = w;      // pad columns may receive garbage
    minmax_bytes(w, cols - c, mn, mx);
  }
  fold_mm8(mn, mx, mm8 + 2 * b);
}

static BlurGeom blur_geom(int kw, int kh) {
  BlurGeom g;
  const int rin = kBlurRows + kh - 1;
  for (int tc = 512; tc >= 16; tc >>= 1) {
    g.tc = tc;
    g.in_words = (tc / 4 + (kw + 3) / 4 + 2 + 1) & ~1;      // even: the Q8.8 plane behind it is read as 64-bit slots
    g.smem = (size_t)rin * g.in_words * 4 + (size_t)rin * (tc / 2) * 4 + (size_t)((kw + 3) / 4) * 4 + (size_t)(kw + kh) * 2 + 16;
    if (g.smem <= (tc > 64 ? 64u : 200u) * 1024u) break;
  }
  return g;
}

// ---- MORPH_CLOSE rect 4x4 then MORPH_OPEN rect 3(time) x 1, one kernel --------------------------------------------------
// close = dilate, erode with the 4x4 rectangle, anchor (2, 2): offsets -2..+1 in both directions;
// open  = erode, dilate with 1 row x 3 columns, anchor (0, 1): offsets -1..+1 along the row.
// Rectangles are separable.  The tile lives in shared memory as 16-bit lanes, two pixels per word, so every min / max
// is one native VIMNMX(3).U16x2 and a one-pixel shift is a 16-bit funnel shift.  A thread owns one word column and half
// of the rows: for each 4x4 stage it walks down its rows, forms the horizontal 4-window of a row from shared memory
// and keeps the last four of them in registers for the vertical window - one shared-memory round trip per 2-D stage.
// OpenCV ignores pixels outside the image: every stage writes, at positions outside the image, the identity of the
// stage that reads it next (0 before a max, 255 before a min).  Tile kMorphRows x kMorphCols outputs with halo 4 rows
// above, 2 below and 8 columns on either side (256 pixels = 128 words per row); results in the halo are garbage
// that never reaches the interior.
constexpr int kMorphRows = 26, kMorphCols = 240;
constexpr int kMorphPR = kMorphRows + 6, kMorphW = (kMorphCols + 16) / 2;
static_assert(kMorphW == 128 && kImgThreads == 2 * kMorphW, "one thread per word column and row half");

template <bool MAX>
__device__ __forceinline__ uint32_t pk2(uint32_t a, uint32_t b) {
  return MAX ? __vmaxs2(a, b) : __vmins2(a, b);           // lanes hold 0..255: signed and unsigned agree
}
template <bool MAX>
__device__ __forceinline__ uint32_t pk3(uint32_t a, uint32_t b, uint32_t c) {
  return MAX ? __vimax3_s16x2(a, b, c) : __vimin3_s16x2(a, b, c);
}

// horizontal window -L..+1 (L = 2: four wide, L = 1: three wide) of one word; pm / pc / pp point at the previous, own
// and next word of the row (clamped at the plane edge by the caller, where the result is halo garbage anyway)
template <bool MAX, int L>
__device__ __forceinline__ uint32_t morph_hwin(const uint32_t* pm, const uint32_t* pc, const uint32_t* pp, int row) {
  const uint32_t wm = pm[row * kMorphW], w = pc[row * kMorphW], wp = pp[row * kMorphW];
  const uint32_t left = __funnelshift_r(wm, w, 16), right = __funnelshift_r(w, wp, 16);   // offsets -1, +1
  const uint32_t v = pk3<MAX>(left, w, right);
  return L == 2 ? pk2<MAX>(v, wm) : v;                                                   // offset -2 = the previous word
}

// one 4x4 stage: rows [y0, y0 + N) of plane `out` = op over rows y-2..y+1, columns -2..+1 of plane `in`.  N is a
// compile-time count so the row loop unrolls into immediate-offset shared-memory accesses.
template <bool MAX, int N>
__device__ __forceinline__ void morph_stage4(const uint32_t* in, uint32_t* out, int k, int km, int kp, int y0, int r0, int rows,
                                             uint32_t cm, uint32_t next_ident) {
  const uint32_t *pm = in + (y0 - 2) * kMorphW + km, *pc = in + (y0 - 2) * kMorphW + k, *pp = in + (y0 - 2) * kMorphW + kp;
  uint32_t* po = out + y0 * kMorphW + k;
  uint32_t h0, h1 = morph_hwin<MAX, 2>(pm, pc, pp, 0), h2 = morph_hwin<MAX, 2>(pm, pc, pp, 1),
               h3 = morph_hwin<MAX, 2>(pm, pc, pp, 2);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    h0 = h1;
    h1 = h2;
    h2 = h3;
    h3 = morph_hwin<MAX, 2>(pm, pc, pp, i + 3);
    const uint32_t v = pk2<MAX>(pk3<MAX>(h0, h1, h2), h3);
    const uint32_t m = (unsigned)(r0 - 4 + y0 + i) < (unsigned)rows ? cm : 0u;
    po[i * kMorphW] = (v & m) | (next_ident & ~m);
  }
}

__global__ void __launch_bounds__(kImgThreads) morph_fused_kernel(const uint8_t* src, int rows, int cols, int64_t pitch,
                                                                  uint8_t* dst, unsigned* mm8) {
  __shared__ uint32_t s_a[kMorphPR * kMorphW], s_b[kMorphPR * kMorphW];
  const int64_t b = blockIdx.z;
  const int r0 = blockIdx.y * kMorphRows, c0 = blockIdx.x * kMorphCols;
  const uint8_t* img = src + b * (int64_t)rows * pitch;
  const int k = threadIdx.x & (kMorphW - 1), half = threadIdx.x >> 7;
  const int km = k > 0 ? k - 1 : 0, kp = k < kMorphW - 1 ? k + 1 : k;
  const int gx = c0 - 8 + 2 * k;                                             // image column of the word's first pixel
  const uint32_t cm = ((unsigned)gx < (unsigned)cols ? 0x0000ffffu : 0u) | ((unsigned)(gx + 1) < (unsigned)cols ? 0xffff0000u : 0u);
  const uint32_t ones = 0x00ff00ffu;                                         // 255 in both lanes
  // load: outside the image -> 0 (the first stage is a max)
  {
    const uint8_t* p = img + (int64_t)(r0 - 4 + half) * pitch + gx;
#pragma unroll
    for (int i = 0; i < kMorphPR / 2; ++i) {
      const int gy = r0 - 4 + half + 2 * i;
      uint32_t w = 0;
      if ((unsigned)gy < (unsigned)rows && cm) {
        const unsigned v = *reinterpret_cast<const uint16_t*>(p + (int64_t)(2 * i) * pitch);
        w = ((v & 255u) | ((v >> 8) << 16)) & cm;
      }
      s_a[(half + 2 * i) * kMorphW + k] = w;
    }
  }
  __syncthreads();
  // Each half takes N consecutive rows of a stage; when the row count is odd the halves overlap by one row, which both
  // compute identically.
  // dilate 4x4 -> rows 2 .. PR-2 of plane b (next: min)
  {
    constexpr int ya = 2, yb = kMorphPR - 1, N = (yb - ya + 1) / 2;
    morph_stage4<true, N>(s_a, s_b, k, km, kp, half ? yb - N : ya, r0, rows, cm, ones);
  }
  __syncthreads();
  // erode 4x4 -> rows 4 .. PR-3 of plane a (next: min)
  constexpr int ya = 4, yb = kMorphPR - 2, N = (yb - ya + 1) / 2;
  const int y0 = half ? yb - N : ya;
  morph_stage4<false, N>(s_b, s_a, k, km, kp, y0, r0, rows, cm, ones);
  __syncthreads();
  // erode 1x3 -> plane b (next: max)
  {
    const uint32_t *pm = s_a + y0 * kMorphW + km, *pc = s_a + y0 * kMorphW + k, *pp = s_a + y0 * kMorphW + kp;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const uint32_t v = morph_hwin<false, 1>(pm, pc, pp, i);
      const uint32_t m = (unsigned)(r0 - 4 + y0 + i) < (unsigned)rows ? cm : 0u;
      s_b[(y0 + i) * kMorphW + k] = v & m;
    }
  }
  __syncthreads();
  // dilate 1x3 straight to global memory: words 4 .. W-5 are the tile's own columns
  unsigned mn = 255u, mx = 0u;
  if (k >= 4 && k < kMorphW - 4 && cm) {
    const uint32_t *pm = s_b + y0 * kMorphW + km, *pc = s_b + y0 * kMorphW + k, *pp = s_b + y0 * kMorphW + kp;
    uint8_t* po = dst + (b * rows + (r0 - 4 + y0)) * pitch + gx;
    uint32_t lo_mn = 0x00ff00ffu, lo_mx = 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (r0 - 4 + y0 + i < rows) {
        const uint32_t v = morph_hwin<true, 1>(pm, pc, pp, i);
        *reinterpret_cast<uint16_t*>(po + (int64_t)i * pitch) = (uint16_t)((v & 255u) | ((v >> 16) << 8));   // pad may get garbage
        lo_mn = __vmins2(lo_mn, (v & cm) | (ones & ~cm));                    // lanes outside the image cannot win
        lo_mx = __vmaxs2(lo_mx, v & cm);
      }
    }
    const unsigned a0 = lo_mn & 0xffffu, a1 = lo_mn >> 16, c0x = lo_mx & 0xffffu, c1x = lo_mx >> 16;
    mn = a0 < a1 ? a0 : a1;
    mx = c0x > c1x ? c0x : c1x;
  }
  fold_mm8(mn, mx, mm8 + 2 * b);
}

// ---- meansub: |x - mean over the row| then the image rescale, two passes over the source -------------------------------
// The source is either a float64 image or - inside the fused chain - a uint8 plane standing for its own rescale
// lut[v] = (v - min) / (max - min): the values, the summation order and therefore every bit are those of the float64
// route, without the float64 image ever being written.
struct MeanSrcF64 {
  typedef double raw_t;                    // what a thread keeps between the two sweeps
  const double* row;
  __device__ __forceinline__ double get(unsigned c) const { return row[c]; }
  __device__ __forceinline__ raw_t raw(unsigned c) const { return row[c]; }
  __device__ __forceinline__ double val(raw_t v) const { return v; }
};
struct MeanSrcU8 {
  typedef uint8_t raw_t;                   // the byte: a quarter of a register instead of two
  const uint8_t* row;
  const double* lut;
  __device__ __forceinline__ double get(unsigned c) const { return lut[row[c]]; }
  __device__ __forceinline__ raw_t raw(unsigned c) const { return row[c]; }
  __device__ __forceinline__ double val(raw_t v) const { return lut[v]; }
};
// rescale table of a uint8 image from its {min, max} slot (same expression as img_rescale_u8_kernel), once per image
__global__ void u8_lut_kernel(const unsigned* mm8, double* lut) {
  const int64_t b = blockIdx.x;
  const uint8_t mn = (uint8_t)mm8[2 * b], mx = (uint8_t)mm8[2 * b + 1];
  const double den = (double)(uint8_t)(mx - mn);
  for (int v = threadIdx.x; v < 256; v += blockDim.x) lut[b * 256 + v] = __ddiv_rn((double)(uint8_t)((uint8_t)v - mn), den);
}
__device__ __forceinline__ void load_u8_lut(double* s_lut, const double* lut, int64_t b) {
  for (int v = threadIdx.x; v < 256; v += blockDim.x) s_lut[v] = lut[b * 256 + v];
  __syncthreads();
}
template <bool U8>
struct MeanSrcSel {
  using type = MeanSrcF64;
};
template <>
struct MeanSrcSel<true> {
  using type = MeanSrcU8;
};
template <bool U8>
__device__ __forceinline__ typename MeanSrcSel<U8>::type mean_src(const void* src, int64_t row, int64_t ld, const double* lut) {
  if constexpr (U8) return MeanSrcU8{static_cast<const uint8_t*>(src) + row * ld, lut};
  else return MeanSrcF64{static_cast<const double*>(src) + row * ld};
}

// CTA-wide min / max: xor-shuffle tree inside each warp, then the warp results
__device__ __forceinline__ void cta_minmax(double& mn, double& mx, double* s8a, double* s8b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, o), c = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    s8a[threadIdx.x >> 5] = mn;
    s8b[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  mn = s8a[0];
  mx = s8b[0];
#pragma unroll
  for (int w = 1; w < kRowThreads / 32; ++w) {
    mn = s8a[w] < mn ? s8a[w] : mn;
    mx = s8b[w] > mx ? s8b[w] : mx;
  }
  __syncthreads();
}

// pass 1: one CTA per row: float64 mean (fixed-order reduction => deterministic), then min / max of |x - mean| (from
// registers for rows up to 4096 columns, else a second read that comes from cache) -> rowstat[row] = {mean, min, max}
constexpr int kMeanRegs = 16;      // rows up to 16 * 256 = 4096 columns stay in registers between the two sweeps

// np.mean(src, axis=1) sums a contiguous float64 row with numpy's pairwise summation (umath loops, `pairwise_sum`):
//   n < 8: sequential from 0;  n <= 128: eight accumulators r[j] += a[i + j] over blocks of eight,
//   ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), then the n % 8 tail sequentially;  n > 128: split at
//   n2 = n / 2 - (n / 2) % 8 and add the two halves.
// The row mean feeds a truncating uint8 quantisation two stages later, so the sum is reproduced operation for
// operation: the host lists the leaves (start, length <= 128) and the post-order combine steps of the recursion for
// this row length; an 8-lane group owns a leaf (lane j = accumulator j, xor-shuffles 1, 2, 4 are exactly the bracketed
// tree), then one thread runs the combine steps.
constexpr int kMeanMaxLevels = 40;
struct MeanPlan {
  const int2* leaves;   // (start, length) in row order
  const int2* steps;    // L[x] = L[x] + L[y]; sorted by the height of the node in the recursion tree
  int nleaf, nlevels;
  int stage_words;      // uint8 source: 32-bit words of a row to stage in shared memory (0: read the bytes from global)
  int level_end[kMeanMaxLevels];   // steps [level_end[h - 1], level_end[h]) are independent of each other
};
constexpr int kMeanMaxLeaves = 2048;     // leaf sums live in shared memory: rows up to ~131 000 columns

template <bool INREG, bool U8>
__global__ void __launch_bounds__(kRowThreads, U8 ? 6 : 4) meansub_stats_kernel(const void* src, int64_t rows, int64_t cols, int64_t ld, const double* lut,
                                     MeanPlan plan, double* rowstat, int rpc) {
  __shared__ double s_lut[U8 ? 256 : 1];
  __shared__ uint8_t s_present[U8 ? 256 : 1];                // which byte values occur in the row
  __shared__ double s8a[kRowThreads / 32], s8b[kRowThreads / 32];
  SPECGPU_DYN_SMEM(smem);
  // dynamic shared memory: leaf sums | the plan's leaves and steps | (uint8 source) the staged row
  const int nlp = (plan.nleaf + 1) & ~1;
  double* s_leaf = reinterpret_cast<double*>(smem);          // [nlp]
  int2* s_leaves = reinterpret_cast<int2*>(s_leaf + nlp);    // [nlp]
  int2* s_steps = s_leaves + nlp;                            // [nlp]
  uint32_t* s_row = reinterpret_cast<uint32_t*>(s_steps + nlp);
  const int64_t b = blockIdx.y;
  const int lane8 = threadIdx.x & 7, group = threadIdx.x >> 3;
  const int64_t rend = min((int64_t)(blockIdx.x + 1) * rpc, rows);
  for (int64_t r = (int64_t)blockIdx.x * rpc; r < rend; ++r) {      // rpc rows per CTA (1 on the GPU)
    const int64_t row = b * rows + r;
    const auto in = mean_src<U8>(src, row, ld, s_lut);
    // float64 source: a strided copy of the row for the second sweep (it also pulls the row into L1 for the leaf sums).
    // uint8 source: the deviations take at most 256 values, so the second sweep runs over the byte values that
    // occur - the leaf sums flag them on the way.
    // All the global loads of the prologue (rescale table, plan, row) are issued together, then one barrier.
    double v[(INREG && !U8) ? kMeanRegs : 1];
    const uint8_t* rowb = nullptr;                                  // uint8 source: the row, staged in shared memory
    if (r == (int64_t)blockIdx.x * rpc) {
      for (int i = threadIdx.x; i < plan.nleaf; i += blockDim.x) {
        s_leaves[i] = plan.leaves[i];
        if (i + 1 < plan.nleaf) s_steps[i] = plan.steps[i];
      }
      if constexpr (U8)
        for (int u = threadIdx.x; u < 256; u += blockDim.x) s_lut[u] = lut[b * 256 + u];
    }
    if constexpr (U8) {
      for (int u = threadIdx.x; u < 256; u += blockDim.x) s_present[u] = 0;
      if (plan.stage_words > 0) {                                   // coalesced words in, strided bytes out
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(in.row);
        for (int i = threadIdx.x; i < plan.stage_words; i += blockDim.x) s_row[i] = gw[i];
        rowb = reinterpret_cast<const uint8_t*>(s_row);
      } else {
        rowb = in.row;
      }
    } else if (INREG) {
#pragma unroll
      for (int q = 0; q < kMeanRegs; ++q) {
        const unsigned c = threadIdx.x + q * kRowThreads;
        v[q] = in.get(c < (unsigned)cols ? c : 0u);
      }
    }
    __syncthreads();
    auto fetch = [&](unsigned c) -> double {
      if constexpr (U8) {
        const uint8_t u = rowb[c];
        s_present[u] = 1;                                           // every writer stores the same value
        return in.lut[u];
      } else {
        return in.get(c);
      }
    };
    // ---- numpy's pairwise row sum ----
    for (int l0 = 0; l0 < plan.nleaf; l0 += kRowThreads / 8) {      // uniform trip count: the shuffles stay converged
      const int leaf = l0 + group;
      const int2 lf = leaf < plan.nleaf ? s_leaves[leaf] : make_int2(0, 0);
      const int start = lf.x, len = lf.y;
      double acc = 0.0;
      if (len >= 8) {
        acc = fetch(start + lane8);
        const int body = len - (len & 7);
#pragma unroll 4
        for (int i = 8; i < body; i += 8) acc = __dadd_rn(acc, fetch(start + i + lane8));
      }
      acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
      acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
      acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
      if (lane8 == 0 && leaf < plan.nleaf) {
        if (len < 8) acc = 0.0;                                     // short rows: sequential from zero
        for (int i = len >= 8 ? len - (len & 7) : 0; i < len; ++i) acc = __dadd_rn(acc, fetch(start + i));
        s_leaf[leaf] = acc;
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                         // warp 0 climbs the tree level by level
      int begin = 0;
      for (int h = 0; h < plan.nlevels; ++h) {
        const int end = plan.level_end[h];
        for (int t = begin + (int)threadIdx.x; t < end; t += 32) {
          const int2 st = s_steps[t];
          s_leaf[st.x] = __dadd_rn(s_leaf[st.x], s_leaf[st.y]);
        }
        begin = end;
        __syncwarp();
      }
      if (threadIdx.x == 0) s8a[0] = __ddiv_rn(s_leaf[0], (double)cols);
    }
    __syncthreads();
    const double mean = s8a[0];
    __syncthreads();
    double mn = INFINITY, mx = -INFINITY;
    if constexpr (U8) {
      for (int u = threadIdx.x; u < 256; u += blockDim.x)
        if (s_present[u]) {
          const double d = fabs(s_lut[u] - mean);
          mn = d < mn ? d : mn;
          mx = d > mx ? d : mx;
        }
    } else if (INREG) {
#pragma unroll
      for (int q = 0; q < kMeanRegs; ++q)
        if (threadIdx.x + q * kRowThreads < (unsigned)cols) {
          const double d = fabs(v[q] - mean);
          mn = d < mn ? d : mn;
          mx = d > mx ? d : mx;
        }
    } else {
      for (unsigned c = threadIdx.x; c < (unsigned)cols; c += blockDim.x) {
        const double d = fabs(in.get(c) - mean);
        mn = d < mn ? d : mn;
        mx = d > mx ? d : mx;
      }
    }
    cta_minmax(mn, mx, s8a, s8b);
    if (threadIdx.x == 0) {
      rowstat[row * 3] = mean;
      rowstat[row * 3 + 1] = mn;
      rowstat[row * 3 + 2] = mx;
    }
  }
}

// fold the row statistics of image b -> imgstat[b] = {min, max - min}; optionally reset the {min, max} slot the next
// uint8 producer folds into
__global__ void meansub_fold_kernel(const double* rowstat, int64_t rows, double* imgstat, unsigned* mm8_next) {
  __shared__ double s8a[kRowThreads / 32], s8b[kRowThreads / 32];
  const int64_t b = blockIdx.x;
  double mn = INFINITY, mx = -INFINITY;
  for (int64_t q = threadIdx.x; q < rows; q += blockDim.x) {
    const double a = rowstat[(b * rows + q) * 3 + 1], c = rowstat[(b * rows + q) * 3 + 2];
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  cta_minmax(mn, mx, s8a, s8b);
  if (threadIdx.x == 0) {
    imgstat[2 * b] = mn;
    imgstat[2 * b + 1] = mx - mn;
    if (mm8_next) {
      mm8_next[2 * b] = 255u;
      mm8_next[2 * b + 1] = 0u;
    }
  }
}

// pass 2: m = (|x - mean| - min) / (max - min).
// QUANT = false: write m (float64).  QUANT = true: write morph's uint8 quantisation of m straight away - m is a
// rescale, so its own min / max are exactly 0 and 1 and (rescale(m) * 255).astype(uint8) is trunc(m * 255).
template <bool U8, bool QUANT>
__global__ void meansub_apply_kernel(const void* src, int64_t rows, int64_t cols, int64_t ld, const double* lut,
                                     const double* rowstat, const double* imgstat, void* dst, int64_t ldo, int rpc) {
  __shared__ double s_lut[U8 ? 256 : 1];
  __shared__ double s_rowd[(U8 && !QUANT) ? 256 : 1];       // a uint8 source has 256 possible results per row:
  __shared__ uint8_t s_rowq[(U8 && QUANT) ? 256 : 1];       // evaluate the expression once per value, then look up
  const int64_t b = blockIdx.y;
  if (U8) load_u8_lut(s_lut, lut, b);
  const double mn = imgstat[2 * b], den = imgstat[2 * b + 1];
  const FastDivD fd(den);
  const int64_t rend = min((int64_t)(blockIdx.x + 1) * rpc, rows);
  for (int64_t r = (int64_t)blockIdx.x * rpc; r < rend; ++r) {
    const double mean = rowstat[(b * rows + r) * 3];
    if constexpr (U8) {
      for (int u = threadIdx.x; u < 256; u += blockDim.x) {
        const double m = fd.div(fabs(s_lut[u] - mean) - mn);
        if constexpr (QUANT) s_rowq[u] = (uint8_t)(int)__dmul_rn(m, 255.0);
        else s_rowd[u] = m;
      }
      __syncthreads();
      const uint8_t* row = static_cast<const uint8_t*>(src) + (b * rows + r) * ld;
      if constexpr (QUANT) {
        // uint8 plane -> uint8 plane, both pitched to 16 bytes: four pixels per word (pad columns receive table values
        // of pad bytes - nobody reads them)
        if (((ld | ldo) & 3) == 0) {
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
          uint32_t* ow = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(dst) + (b * rows + r) * ldo);
          const unsigned nw = ((unsigned)cols + 3) >> 2;
          for (unsigned i = threadIdx.x; i < nw; i += kRowThreads) {
            const uint32_t w = rw[i];
            ow[i] = (uint32_t)s_rowq[w & 255u] | ((uint32_t)s_rowq[(w >> 8) & 255u] << 8) |
                    ((uint32_t)s_rowq[(w >> 16) & 255u] << 16) | ((uint32_t)s_rowq[w >> 24] << 24);
          }
          __syncthreads();
          continue;
        }
      }
      for (unsigned c0 = threadIdx.x; c0 < (unsigned)cols; c0 += 4 * kRowThreads) {
        uint8_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned c = c0 + k * kRowThreads;
          v[k] = row[c < (unsigned)cols ? c : c0];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned c = c0 + k * kRowThreads;
          if (c < (unsigned)cols) {
            if constexpr (QUANT) static_cast<uint8_t*>(dst)[(b * rows + r) * ldo + c] = s_rowq[v[k]];
            else static_cast<double*>(dst)[(b * rows + r) * ldo + c] = s_rowd[v[k]];
          }
        }
      }
      __syncthreads();                       // the next row rebuilds the table
    } else {
      const double* row = static_cast<const double*>(src) + (b * rows + r) * ld;
      for (unsigned c0 = threadIdx.x; c0 < (unsigned)cols; c0 += 4 * kRowThreads) {
        double v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned c = c0 + k * kRowThreads;
          v[k] = row[c < (unsigned)cols ? c : c0];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned c = c0 + k * kRowThreads;
          if (c < (unsigned)cols) {
            const double m = fd.div(fabs(v[k] - mean) - mn);
            if constexpr (QUANT) static_cast<uint8_t*>(dst)[(b * rows + r) * ldo + c] = (uint8_t)(int)__dmul_rn(m, 255.0);
            else static_cast<double*>(dst)[(b * rows + r) * ldo + c] = m;
          }
        }
      }
    }
  }
}

// ---- launchers ----------------------------------------------------------------------------------------------------
size_t imgchain_workspace_bytes(int64_t B, int64_t rows, int64_t cols) {
  const size_t plane = (size_t)B * rows * img_pitch(cols);
  // two pitched uint8 planes, float/double min-max partials, uint8 min-max slots, row statistics, kernel taps
  return 2 * (plane + 256) + (size_t)B * kImgParts * 2 * 8 + 256 + 2 * ((size_t)B * 2 * 4 + 256) + (size_t)B * rows * 3 * 8 + 256 +
         (size_t)B * 256 * 8 + 256 + (size_t)B * 2 * 8 + 256 + 2 * ((size_t)kMeanMaxLeaves * 8 + 256) + 4096;
}

namespace {
// rows per CTA of the row kernels.  On the GPU one row per CTA measured best (40 x 256 rows: filter chain 0.81 ms
// against 0.84 ms with 9 rows per CTA in a single wave - the short CTAs overlap their prologues better than a loop
// amortises them); the CPU emulation runs CTAs one after another and pays per CTA, so it folds the rows into 16 CTAs.
#if defined(SPECGPU_EMULATE)
constexpr int64_t kRowSlots = 16;
#else
constexpr int64_t kRowSlots = (int64_t)1 << 40;
#endif
inline int rows_per_cta(int64_t B, int64_t rows) { return (int)std::max<int64_t>(1, ceil_div(B * rows, kRowSlots)); }
inline dim3 row_grid(int64_t B, int64_t rows) { return dim3((unsigned)ceil_div(rows, rows_per_cta(B, rows)), (unsigned)B); }
struct ImgWs {
  uint8_t *u8a, *u8b;
  void* part;
  unsigned *mm8, *mm8b;
  double *rowstat, *lut, *imgstat;
  int2 *leaves, *steps;
  uint16_t* taps;
};
ImgWs img_carve(void* ws, int64_t B, int64_t rows, int64_t cols) {
  const size_t plane = (size_t)B * rows * img_pitch(cols);
  char* p = static_cast<char*>(ws);
  auto take = [&](size_t bytes) {
    char* q = p;
    p += (bytes + 255) & ~(size_t)255;
    return q;
  };
  ImgWs w;
  w.u8a = reinterpret_cast<uint8_t*>(take(plane));
  w.u8b = reinterpret_cast<uint8_t*>(take(plane));
  w.part = take((size_t)B * kImgParts * 2 * 8);
  w.mm8 = reinterpret_cast<unsigned*>(take((size_t)B * 2 * 4));
  w.mm8b = reinterpret_cast<unsigned*>(take((size_t)B * 2 * 4));
  w.rowstat = reinterpret_cast<double*>(take((size_t)B * rows * 3 * 8));
  w.lut = reinterpret_cast<double*>(take((size_t)B * 256 * 8));
  w.imgstat = reinterpret_cast<double*>(take((size_t)B * 2 * 8));
  w.leaves = reinterpret_cast<int2*>(take((size_t)kMeanMaxLeaves * 8));
  w.steps = reinterpret_cast<int2*>(take((size_t)kMeanMaxLeaves * 8));
  w.taps = reinterpret_cast<uint16_t*>(take(4096));
  return w;
}
template <class T>
void run_minmax(const T* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, T* part, cudaStream_t st) {
  SPECGPU_LAUNCH((img_minmax_kernel<T>), dim3((unsigned)std::min<int64_t>(kImgParts, rows), (unsigned)B), kRowThreads, 0, st, src, rows, cols, ld, part);
}
// min / max of the source, then its uint8 quantisation into plane `dst`
void run_quantise(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, const ImgWs& w, uint8_t* dst,
                  cudaStream_t st) {
  const dim3 rowgrid = row_grid(B, rows);
  const int rpc = rows_per_cta(B, rows);
  const int64_t pitch = img_pitch(cols);
  if (in_f64) {
    run_minmax<double>((const double*)src, B, rows, cols, ld, (double*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<double>), rowgrid, kRowThreads, 0, st, (const double*)src, rows, cols, ld,
                   (const double*)w.part, dst, pitch, w.mm8, rpc);
  } else {
    run_minmax<float>((const float*)src, B, rows, cols, ld, (float*)w.part, st);
    SPECGPU_LAUNCH((img_quantise_kernel<float>), rowgrid, kRowThreads, 0, st, (const float*)src, rows, cols, ld,
                   (const float*)w.part, dst, pitch, w.mm8, rpc);
  }
}
// rescale of the uint8 result (+ the dense copy when the caller wants the uint8 image)
void run_rescale_u8(const uint8_t* plane, int64_t B, int64_t rows, int64_t cols, const ImgWs& w, double* dst, int64_t ldo,
                    uint8_t* u8_out, cudaStream_t st) {
  const int64_t pitch = img_pitch(cols);
  SPECGPU_LAUNCH(img_rescale_u8_kernel, row_grid(B, rows), kRowThreads, 0, st, plane, rows, cols, pitch,
                 (const unsigned*)w.mm8, dst, ldo, rows_per_cta(B, rows));
  if (u8_out) SPECGPU_LAUNCH(img_unpitch_kernel, (unsigned)(B * rows), kRowThreads, 0, st, plane, cols, pitch, u8_out);
}
// both meansub passes from source `src` (float64 image or uint8 plane + its {min, max} slot)
// leaves and combine steps of numpy's pairwise_sum recursion for a row of n elements; a step carries the height of its
// node so that the steps of one height (independent of each other) can run in parallel.  Returns (first leaf, height).
struct MeanStep {
  int x, y, h;
};
std::pair<int, int> mean_plan_build(int start, int n, std::vector<int2>& leaves, std::vector<MeanStep>& steps) {
  if (n <= 128) {
    leaves.push_back(make_int2(start, n));
    return {(int)leaves.size() - 1, 0};
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  const auto l = mean_plan_build(start, n2, leaves, steps);
  const auto r = mean_plan_build(start + n2, n - n2, leaves, steps);
  const int h = std::max(l.second, r.second) + 1;
  steps.push_back(MeanStep{l.first, r.first, h});
  return {l.first, h};
}

template <bool U8, bool QUANT>
int run_meansub(const void* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, const unsigned* mm8, const ImgWs& w, void* dst,
                int64_t ldo, unsigned* mm8_next, cudaStream_t st) {
  const dim3 rowgrid = row_grid(B, rows);
  const int rpc = rows_per_cta(B, rows);
  std::vector<int2> leaves, steps;
  std::vector<MeanStep> tree;
  const int height = mean_plan_build(0, (int)cols, leaves, tree).second;
  if ((int)leaves.size() > kMeanMaxLeaves || height > kMeanMaxLevels) return -1;
  MeanPlan plan{};
  plan.leaves = w.leaves;
  plan.steps = w.steps;
  plan.nleaf = (int)leaves.size();
  plan.nlevels = height;
  for (int h = 1; h <= height; ++h) {                 // stable by height: children always sit in lower levels
    for (const MeanStep& t : tree)
      if (t.h == h) steps.push_back(make_int2(t.x, t.y));
    plan.level_end[h - 1] = (int)steps.size();
  }
  cudaError_t e = cudaMemcpyAsync(w.leaves, leaves.data(), leaves.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && !steps.empty())
    e = cudaMemcpyAsync(w.steps, steps.data(), steps.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  size_t smem = ((leaves.size() + 1) & ~(size_t)1) * (sizeof(double) + 2 * sizeof(int2));      // leaf sums + plan copy
  if (U8 && (ld & 3) == 0 && cols <= 16384) {         // pitched plane rows: whole words, up to 16 KB per row
    plan.stage_words = (int)((cols + 3) / 4);
    smem += (size_t)plan.stage_words * 4;
  }
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(meansub_stats_kernel<true, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(meansub_stats_kernel<false, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (U8) SPECGPU_LAUNCH(u8_lut_kernel, (unsigned)B, kRowThreads, 0, st, mm8, w.lut);
  if (cols <= kMeanRegs * kRowThreads)
    SPECGPU_LAUNCH((meansub_stats_kernel<true, U8>), rowgrid, kRowThreads, smem, st, src, rows, cols, ld, (const double*)w.lut,
                   plan, w.rowstat, rpc);
  else
    SPECGPU_LAUNCH((meansub_stats_kernel<false, U8>), rowgrid, kRowThreads, smem, st, src, rows, cols, ld, (const double*)w.lut,
                   plan, w.rowstat, rpc);
  SPECGPU_LAUNCH(meansub_fold_kernel, (unsigned)B, kRowThreads, 0, st, (const double*)w.rowstat, rows, w.imgstat, mm8_next);
  SPECGPU_LAUNCH((meansub_apply_kernel<U8, QUANT>), rowgrid, kRowThreads, 0, st, src, rows, cols, ld, (const double*)w.lut,
                 (const double*)w.rowstat, (const double*)w.imgstat, dst, ldo, rpc);
  return (int)cudaGetLastError();
}
int run_blur(const ImgWs& w, int64_t B, int64_t rows, int64_t cols, const uint16_t* taps_host, int kw, int kh, const uint8_t* in,
             uint8_t* out, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(w.taps, taps_host, (size_t)(kw + kh) * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  int packed = 1;
  for (int d = 0; d < kw; ++d) packed &= taps_host[d] <= 255;
  const BlurGeom g = blur_geom(kw, kh);
  const dim3 grid((unsigned)ceil_div(cols, g.tc), (unsigned)ceil_div(rows, kBlurRows), (unsigned)B);
  const int G = (kw + 3) / 4;
#define SPECGPU_BLUR(GT)                                                                                                    \
  do {                                                                                                                      \
    e = cudaFuncSetAttribute(blur_fused_kernel<GT>, cudaFuncAttributeMaxDynamicSharedMemorySize,                           \
                             (int)std::max<size_t>(g.smem, 48 * 1024));                                                     \
    if (e != cudaSuccess) return (int)e;                                                                                    \
    SPECGPU_LAUNCH(blur_fused_kernel<GT>, grid, kImgThreads, g.smem, st, in, (int)rows, (int)cols, img_pitch(cols),         \
                   (const uint16_t*)w.taps, kw, kh, g.tc, g.in_words, packed, out, w.mm8);                                  \
  } while (0)
  if (packed && G == 8) SPECGPU_BLUR(8);          // ksize 29 / 31 (the reference's (31, 3))
  else if (packed && G == 2) SPECGPU_BLUR(2);     // ksize 5 / 7
  else SPECGPU_BLUR(0);
#undef SPECGPU_BLUR
  return 0;
}
}  // namespace

int launch_meansub(const double* src, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst, int64_t ldo,
                   cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  return run_meansub<false, false>(src, B, rows, cols, ld, nullptr, w, dst, ldo, nullptr, st);
}

// gaussblr -> meansub -> morph -> meansub of pipeline_data.py:104-110 on an already thresholded float32 image, with
// every intermediate kept as a uint8 plane + its {min, max}: 12 launches, the only float64 traffic is the result.
int launch_filter_tail(const float* q, int64_t B, int64_t rows, int64_t cols, int64_t ld, const uint16_t* taps_host, int kw,
                       int kh, void* ws, double* dst, int64_t ldo, cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  const int64_t pitch = img_pitch(cols);
  run_quantise(q, 0, B, rows, cols, ld, w, w.u8a, st);                                    // gaussblr: quantise ...
  int rc = run_blur(w, B, rows, cols, taps_host, kw, kh, w.u8a, w.u8b, st);              // ... blur -> plane b, slot mm8
  if (rc) return rc;
  // meansub of the blurred image + morph's quantisation -> plane a; resets slot mm8b
  if ((rc = run_meansub<true, true>(w.u8b, B, rows, cols, pitch, w.mm8, w, w.u8a, pitch, w.mm8b, st))) return rc;
  const dim3 grid((unsigned)ceil_div(cols, kMorphCols), (unsigned)ceil_div(rows, kMorphRows), (unsigned)B);
  SPECGPU_LAUNCH(morph_fused_kernel, grid, kImgThreads, 0, st, (const uint8_t*)w.u8a, (int)rows, (int)cols, pitch, w.u8b,
                 w.mm8b);                                                                 // -> plane b, slot mm8b
  return run_meansub<true, false>(w.u8b, B, rows, cols, pitch, w.mm8b, w, dst, ldo, nullptr, st);  // final meansub -> float64
}

// taps_host: kw + kh Q8.8 taps (kx then ky), built by the caller.
int launch_gaussblr(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, const uint16_t* taps_host,
                    int kw, int kh, void* ws, double* dst, int64_t ldo, uint8_t* u8_out, cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  run_quantise(src, in_f64, B, rows, cols, ld, w, w.u8a, st);
  const int rc = run_blur(w, B, rows, cols, taps_host, kw, kh, w.u8a, w.u8b, st);
  if (rc) return rc;
  run_rescale_u8(w.u8b, B, rows, cols, w, dst, ldo, u8_out, st);
  return (int)cudaGetLastError();
}

int launch_morph(const void* src, int in_f64, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* ws, double* dst,
                 int64_t ldo, uint8_t* u8_out, cudaStream_t st) {
  if (B * rows * cols == 0) return 0;
  ImgWs w = img_carve(ws, B, rows, cols);
  run_quantise(src, in_f64, B, rows, cols, ld, w, w.u8a, st);
  const dim3 grid((unsigned)ceil_div(cols, kMorphCols), (unsigned)ceil_div(rows, kMorphRows), (unsigned)B);
  SPECGPU_LAUNCH(morph_fused_kernel, grid, kImgThreads, 0, st, (const uint8_t*)w.u8a, (int)rows, (int)cols, img_pitch(cols),
                 w.u8b, w.mm8);
  run_rescale_u8(w.u8b, B, rows, cols, w, dst, ldo, u8_out, st);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
