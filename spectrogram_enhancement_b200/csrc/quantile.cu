// K4: per-column quantile threshold (quantfilt, spec_denoising/pipeline_data.py:46-49).
//
// np.quantile(src, thr, axis=0) with numpy's default 'linear' method needs the order statistics
// lo = floor((rows-1)*thr) and lo+1 of every column.  A CTA stages a [rows x 32 columns] tile in shared
// memory (row-contiguous loads), each warp takes four of the columns and either pops the few extreme elements it
// needs (order statistics near one end, e.g. the reference's 0.9 quantile) or fully sorts the column with a
// register/shuffle bitonic network (element e = r*32 + lane, E = rows_pad/32 registers per lane), interpolates with
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

// Selection instead of a full sort when the wanted order statistics sit near one end (the reference's thr = 0.9 on 256
// rows needs the 26th and 27th largest): every lane sorts its E values in registers, then the warp pops the extreme
// `npop` times (warp max/min of the lane heads, the lowest owning lane shifts its list).  The four columns of a warp run together: four independent
// pop chains hide the shuffle latency of each other.  `last` / `before` are the last two values popped per column.
template <int E, bool TOP>
__device__ __forceinline__ void pop_select4(unsigned (&v)[4][E], int npop, int lane, unsigned (&last)[4], unsigned (&before)[4]) {
  // keys are order-preserving uint32 images of the floats, so the warp extreme is one redux.sync instead of a
  // five-step shuffle ladder
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < E; ++i)
#pragma unroll
      for (int j = 0; j + 1 < E - i; ++j) {
        const unsigned a = v[c][j], b = v[c][j + 1];
        v[c][j] = TOP ? max(a, b) : min(a, b);
        v[c][j + 1] = TOP ? min(a, b) : max(a, b);
      }
#pragma unroll
  for (int c = 0; c < 4; ++c) last[c] = before[c] = 0u;
  for (int r = 0; r < npop; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const unsigned m = TOP ? __reduce_max_sync(0xffffffffu, v[c][0]) : __reduce_min_sync(0xffffffffu, v[c][0]);
      const unsigned owners = __ballot_sync(0xffffffffu, v[c][0] == m);
      if (lane == __ffs((int)owners) - 1) {      // one instance leaves, duplicates stay with their lanes
#pragma unroll
        for (int i = 0; i + 1 < E; ++i) v[c][i] = v[c][i + 1];
        v[c][E - 1] = TOP ? 0u : 0xffffffffu;
      }
      before[c] = last[c];
      last[c] = m;
    }
  }
}

// sel: 0 = full bitonic sort; 1 = pop the (rows - lo) largest; 2 = pop the (lo + 2) smallest (E <= 8 only)
template <int E>
__global__ void __launch_bounds__(kQThreads) quantfilt_kernel(const float* src, int rows, int64_t cols, int64_t ld, int lo,
                                                              float g, int sel, float* dst, float* thr_out, uint8_t* mask) {
  SPECGPU_DYN_SMEM(smem);
  float* tile = reinterpret_cast<float*>(smem);               // [rows][kQPitch]
  float* s_thr = tile + (size_t)rows * kQPitch;                // [kQCols]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.y;
  const int64_t c0 = (int64_t)blockIdx.x * kQCols;
  const int ncol = (int)((cols - c0 < kQCols) ? (cols - c0) : kQCols);
  const float* sb = src + b * rows * ld;

  {
    const float* p = sb + (int64_t)warp * ld + c0 + lane;
    const int64_t step = (int64_t)(kQThreads / 32) * ld;
    const bool ok = lane < ncol;
    for (int r = warp; r < rows; r += kQThreads / 32, p += step) tile[r * kQPitch + lane] = ok ? __ldg(p) : 0.f;
  }
  __syncthreads();

  // numpy _lerp in float32: a + (b-a)*g, replaced by b - (b-a)*(1-g) where g >= 0.5
  auto lerp_np = [&](float a, float bb) {
    const float d = __fsub_rn(bb, a);
    if (lo >= rows - 1) return a;
    if (g >= 0.5f) return __fsub_rn(bb, __fmul_rn(d, __fsub_rn(1.0f, g)));
    return __fadd_rn(a, __fmul_rn(d, g));
  };
  if (E <= 8 && sel != 0) {
    if constexpr (E <= 8) {
      // the warp's four columns (warp, warp + 8, warp + 16, warp + 24) together; columns past ncol hold zeros
      static_assert(kQCols == 4 * (kQThreads / 32), "four columns per warp");
      unsigned v[4][E];
      const unsigned padv = (sel == 1) ? 0u : 0xffffffffu;       // padding never reaches the popped end
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int i = r * 32 + lane;
          v[c][r] = (i < rows) ? float_to_ordered(tile[i * kQPitch + warp + 8 * c]) : padv;
        }
      unsigned last[4], before[4];
      if (sel == 1) pop_select4<E, true>(v, rows - lo, lane, last, before);      // last = sorted[lo], before = sorted[lo + 1]
      else pop_select4<E, false>(v, lo + 2, lane, last, before);                  // last = sorted[lo + 1], before = sorted[lo]
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a = ordered_to_float((sel == 1) ? last[c] : before[c]);
        float bb = ordered_to_float((sel == 1) ? before[c] : last[c]);
        if (lo + 1 >= rows) bb = a;
        if (lane == 0 && warp + 8 * c < ncol) s_thr[warp + 8 * c] = lerp_np(a, bb);
      }
    }
  } else {
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
      if (lane == 0) s_thr[c] = lerp_np(a, bb);
    }
  }
  __syncthreads();

  if (thr_out != nullptr && tid < ncol) thr_out[b * cols + c0 + tid] = s_thr[tid];
  const float q = (lane < ncol) ? s_thr[lane] : 0.f;
  if (lane < ncol) {
    int64_t o = (b * rows + warp) * ld + c0 + lane;
    const int64_t step = (int64_t)(kQThreads / 32) * ld;
    for (int r = warp; r < rows; r += kQThreads / 32, o += step) {
      const float x = tile[r * kQPitch + lane];
      const bool below = x < q;
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
  // selection beats the full sort when few elements have to be popped (about 22 instructions per pop against ~1600
  // for the 256-element bitonic network)
  int sel = 0;
  if (E <= 8) {
    const int ktop = rows - lo, kbot = lo + 2;
    if (ktop <= kbot && ktop <= 40) sel = 1;
    else if (kbot < ktop && kbot <= 40 && lo + 1 < rows) sel = 2;
  }
  SPECGPU_LAUNCH(kern, dim3((unsigned)ceil_div(cols, kQCols), (unsigned)B), kQThreads, smem, stream, src, rows, cols, ld,
                 lo, g, sel, dst, thr_out, mask);
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
