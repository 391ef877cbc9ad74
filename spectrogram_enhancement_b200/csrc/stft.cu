// K1/K2a: batched overlapped-segment detrend + window + real FFT with fused epilogues.
//
// Replaces scipy.signal.spectrogram / stft / the spectra half of csd as called by the reference
// (spec_denoising/pipeline_data.py:32-35, interferometer/crosspowerspec.py:39).  One CTA handles a
// tile of TT consecutive segments of one signal; groups of G = M/R0 threads each transform one
// segment (M = nperseg/2 complex points, see fft.cuh), untangle the packed real transform into the
// one-sided spectrum, apply the epilogue of the requested mode and stage the result in a shared-
// memory tile so that the [freq][time] output is written in time-contiguous runs.
#include <type_traits>

#include "fft.cuh"
#include "kernels.h"

namespace specgpu {

constexpr int kStftThreads = 256;

template <int LOG2N>
struct StftCfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2M = LOG2N - 1;
  static constexpr int M = N / 2;
  static constexpr int F = M + 1;
  static constexpr int R0 = fft_radix_at(LOG2M, 0);
  static constexpr int G = M / R0;                      // threads per segment
  static constexpr int NG = kStftThreads / G;           // segments in flight per CTA
  static constexpr int LINE = M + (M >> 4) + 1;          // padded float2 per FFT line
  // Extra columns of the transposed output tile.  A warp holds 32/G segment groups (consecutive tile columns) whose
  // threads store 16 consecutive frequencies each: with a pitch of TT + 32/G (float tiles) the 32 lanes of such a
  // store fall into 32 different banks (pitch TT + 1 left the groups one bank apart: 2-way conflicts).
  static constexpr int PAD4 = (G >= 2 && G <= 16) ? 32 / G : 1;
  // tile width (segments per CTA): >= NG, grown towards 32 while the float tile stays <= 72 KB
  __host__ __device__ static constexpr int tile_w(int bytes_per_elem) {
    int tt = NG;
    while (tt < 32 && (long)F * (2 * tt + 1) * bytes_per_elem <= 72 * 1024) tt *= 2;
    return tt;
  }
};

struct StftSmem {
  int window_off, twm_off, twn_off, line_off, red_off, tile_off, total;
};

template <int LOG2N>
__host__ __device__ inline StftSmem stft_smem_layout(int mode) {
  using C = StftCfg<LOG2N>;
  StftSmem s;
  int off = 0;
  s.window_off = off; off += C::N * 4;
  s.twm_off = off;    off += C::M * 8;
  s.twn_off = off;    off += (C::M / 2 + 1) * 8;
  off = (off + 15) & ~15;
  s.line_off = off;   off += C::NG * C::LINE * 8;
  s.red_off = off;    off += (kStftThreads / 32) * 2 * 4 + 64;
  off = (off + 15) & ~15;
  s.tile_off = off;
  if (mode == STFT_MODE_PSD || mode == STFT_MODE_LOGPSD) off += C::F * (C::tile_w(4) + C::PAD4) * 4;
  else if (mode == STFT_MODE_COMPLEX) off += C::F * (C::tile_w(8) + 1) * 8;
  s.total = off;
  return s;
}

// Sum of (a, b) over the G threads of a segment group.
template <int G>
__device__ __forceinline__ void group_sum2(float& a, float& b, float* red, int tid) {
  if constexpr (G == 1) {
    return;
  } else if constexpr (G <= 32) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  } else {
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = tid >> 5;
    if ((tid & 31) == 0) {
      red[2 * w] = a;
      red[2 * w + 1] = b;
    }
    __syncthreads();
    constexpr int WPG = G / 32;
    const int w0 = (w / WPG) * WPG;
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < WPG; ++i) {
      sa += red[2 * (w0 + i)];
      sb += red[2 * (w0 + i) + 1];
    }
    a = sa;
    b = sb;
    __syncthreads();
  }
}

// Resident CTAs per SM the register allocator should aim for: what the shared-memory footprint of the mode allows
// (the spectra mode has no output tile).  Without it the 3-pass sizes compile to ~195 registers = one CTA per SM.
__host__ __device__ constexpr int stft_min_blocks(int log2n, int mode) {
  if (mode == STFT_MODE_COMPLEX) return 1;
  if (log2n <= 9) return 3;
  if (mode == STFT_MODE_SPECTRA) return log2n <= 12 ? 3 : 1;
  return log2n == 10 ? 2 : 1;
}

template <int LOG2N, int MODE>
__global__ void __launch_bounds__(kStftThreads, stft_min_blocks(LOG2N, MODE)) stft_kernel(StftArgs a) {
  using C = StftCfg<LOG2N>;
  constexpr int N = C::N, M = C::M, F = C::F, R0 = C::R0, G = C::G, NG = C::NG;
  constexpr int TT = (MODE == STFT_MODE_COMPLEX) ? C::tile_w(8) : C::tile_w(4);
  constexpr int PITCH = (MODE == STFT_MODE_COMPLEX) ? TT + 1 : TT + C::PAD4;
  SPECGPU_DYN_SMEM(smem);
  const StftSmem L = stft_smem_layout<LOG2N>(MODE);
  float* s_win = reinterpret_cast<float*>(smem + L.window_off);
  float2* s_twm = reinterpret_cast<float2*>(smem + L.twm_off);
  float2* s_twn = reinterpret_cast<float2*>(smem + L.twn_off);
  float2* s_line = reinterpret_cast<float2*>(smem + L.line_off);
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);
  float* s_tile = reinterpret_cast<float*>(smem + L.tile_off);
  float2* s_tile2 = reinterpret_cast<float2*>(smem + L.tile_off);

  const int tid = threadIdx.x;
  const int grp = tid / G;  // which in-flight segment
  const int tg = tid % G;   // thread within the segment group

  for (int i = tid; i < N; i += kStftThreads) s_win[i] = a.window[i];
  for (int i = tid; i < M; i += kStftThreads) s_twm[i] = a.twM[i];
  for (int i = tid; i <= M / 2; i += kStftThreads) s_twn[i] = a.twN[i];
  __syncthreads();

  float2* line = s_line + grp * C::LINE;
  // persistent CTAs: the tables above are loaded once, then the CTA walks tiles (signal b, TT segments)
  const unsigned tps = (unsigned)a.tiles_per_signal;      // ntiles < 2^31 is checked by the launcher
  for (unsigned tile = blockIdx.x; tile < (unsigned)a.ntiles; tile += gridDim.x) {
  const unsigned b32 = tile / tps;
  const int64_t b = b32;
  const int64_t seg0 = (int64_t)(tile - b32 * tps) * TT;
  const float* xb = a.x + b * a.ldx;
  float vmin = INFINITY, vmax = -INFINITY;

  for (int round = 0; round < TT / NG; ++round) {
    const int tl = round * NG + grp;  // column inside the tile
    const int64_t seg = seg0 + tl;
    const bool live = seg < a.nseg;
    const int64_t s0 = a.first_start + seg * (int64_t)a.hop;

    // ---- load the group's segment: element r of this thread is complex sample m = tg + r*G ----
    float2 v[R0];
    if (live && a.vec_ok && s0 >= 0 && s0 + N <= a.n) {
      const float2* p = reinterpret_cast<const float2*>(xb + s0);
#pragma unroll
      for (int r = 0; r < R0; ++r) v[r] = __ldg(p + tg + r * G);
    } else {
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const int64_t i0 = s0 + 2 * (tg + r * G);
        v[r].x = (live && i0 >= 0 && i0 < a.n) ? __ldg(xb + i0) : 0.f;
        v[r].y = (live && i0 + 1 >= 0 && i0 + 1 < a.n) ? __ldg(xb + i0 + 1) : 0.f;
      }
    }
    // ---- detrend (scipy.signal.detrend per segment) ----
    if (a.detrend != SPECGPU_DETREND_NONE) {
      float sx = 0.f, sc = 0.f;
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const float c0 = (float)(2 * (tg + r * G)) - 0.5f * (float)(N - 1);
        sx += v[r].x + v[r].y;
        sc += c0 * v[r].x + (c0 + 1.0f) * v[r].y;
      }
      group_sum2<G>(sx, sc, s_red, tid);
      const float mean = sx * (1.0f / (float)N);
      // sum_n c_n^2 = N (N^2 - 1) / 12
      const float slope = (a.detrend == SPECGPU_DETREND_LINEAR)
                              ? sc * (12.0f / ((float)N * ((float)N * (float)N - 1.0f)))
                              : 0.f;
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const float c0 = (float)(2 * (tg + r * G)) - 0.5f * (float)(N - 1);
        v[r].x -= mean + slope * c0;
        v[r].y -= mean + slope * (c0 + 1.0f);
      }
    }
    // ---- window ----
#pragma unroll
    for (int r = 0; r < R0; ++r) {
      const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * (tg + r * G));
      v[r].x *= w.x;
      v[r].y *= w.y;
    }
    // ---- M-point complex FFT of the packed segment ----
    fft_group<C::LOG2M>(v, line, s_twm, tg);

    // ---- untangle: 2 X[k] = E + W_N^k O, 2 X[M-k] = conj(E - W_N^k O) with E = Z[k] + conj(Z[M-k]),
    //      O = -i (Z[k] - conj(Z[M-k])); the factor 2 is folded into the output scales.  (Splitting the special bins
    //      k = 0 and k = M/2 out of the loop was measured 4 % slower: they belong to one thread per group, so the
    //      warp issues two more, nearly empty iterations.) ----
    const float cscale = 0.5f * a.scale;             // complex / spectra outputs
    const float pscale1 = 0.25f * a.scale;           // |2X|^2 -> PSD, bins 0 and Nyquist
    const float pscale2 = 0.5f * a.scale;            // one-sided doubling for every other bin
    auto bin_pair = [&](int k, auto generic_c) {
      constexpr bool GENERIC = decltype(generic_c)::value;     // 0 < k < M/2: two distinct, doubled bins
      const float2 zk = line[fft_pad(k)];
      const float2 zm = line[fft_pad((M - k) & (M - 1))];
      const float2 e = make_float2(zk.x + zm.x, zk.y - zm.y);
      const float2 o = make_float2(zk.y + zm.y, zm.x - zk.x);
      const float2 wo = cmul(s_twn[k], o);
      const float2 xk = cadd(e, wo);
      float2 xm = csub(e, wo);
      xm.y = -xm.y;
      const int km = M - k;
      const bool two = GENERIC || km != k;
      if (MODE == STFT_MODE_SPECTRA) {
        if (live) {
          float2* o2 = reinterpret_cast<float2*>(a.out) + (b * a.nseg + seg) * a.ld_out;
          o2[k] = make_float2(0.5f * xk.x, 0.5f * xk.y);
          if (two) o2[km] = make_float2(0.5f * xm.x, 0.5f * xm.y);
        }
      } else if (MODE == STFT_MODE_COMPLEX) {
        s_tile2[k * PITCH + tl] = make_float2(xk.x * cscale, xk.y * cscale);
        if (two) s_tile2[km * PITCH + tl] = make_float2(xm.x * cscale, xm.y * cscale);
      } else {
        // conj(X) X scale, doubled on 1..M-1 (one-sided, even nfft): only k == 0 (bins 0 and M) is not doubled
        const float ps = (GENERIC || k != 0) ? pscale2 : pscale1;
        float pk = xk.x * xk.x + xk.y * xk.y;
        float pm = xm.x * xm.x + xm.y * xm.y;
        if (MODE == STFT_MODE_LOGPSD) {
          // log2 (one MUFU): the min-max normalisation that follows is invariant to the base of the logarithm, only the
          // exported (min, max) are converted to natural logs.  lg2.approx: absolute error ~1e-6 on values in [-37, 14].
          pk = __log2f(fmaf(pk, ps, a.eps));
          pm = __log2f(fmaf(pm, ps, a.eps));
          if (live) {
            vmin = fminf(vmin, fminf(pk, pm));       // k == km (k = M/2) gives pk == pm: harmless
            vmax = fmaxf(vmax, fmaxf(pk, pm));
          }
        } else {
          pk *= ps;
          pm *= ps;
        }
        s_tile[k * PITCH + tl] = pk;
        if (two) s_tile[km * PITCH + tl] = pm;
      }
    };
    for (int k = tg; k <= M / 2; k += G) bin_pair(k, std::false_type{});
    fft_group_sync<G>();  // line is reused by the next round
  }

  if (MODE == STFT_MODE_SPECTRA) continue;
  __syncthreads();   // the tile is complete

  // ---- write the tile: rows = frequency, runs of up to TT consecutive segments ----
  const int64_t ncol = (a.nseg - seg0 < TT) ? (a.nseg - seg0) : TT;
  const int rows_out = (MODE == STFT_MODE_LOGPSD) ? (F - 1) : F;  // Nyquist row dropped after min/max
  constexpr int LANES_T = TT < 32 ? TT : 32;   // lanes along time
  constexpr int ROWS_W = 32 / LANES_T;         // rows per warp step
  const int lane = tid & 31, warp = tid >> 5;
  const int lt = lane % LANES_T, lr = lane / LANES_T;
  for (int k = warp * ROWS_W + lr; k < rows_out; k += (kStftThreads / 32) * ROWS_W) {
    for (int t = lt; t < ncol; t += LANES_T) {
      const int64_t o = (b * rows_out + k) * a.ld_out + seg0 + t;
      if (MODE == STFT_MODE_COMPLEX) reinterpret_cast<float2*>(a.out)[o] = s_tile2[k * PITCH + t];
      else reinterpret_cast<float*>(a.out)[o] = s_tile[k * PITCH + t];
    }
  }

  if (MODE == STFT_MODE_LOGPSD) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) {
      s_red[2 * warp] = vmin;
      s_red[2 * warp + 1] = vmax;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kStftThreads / 32; ++w) {
        vmin = fminf(vmin, s_red[2 * w]);
        vmax = fmaxf(vmax, s_red[2 * w + 1]);
      }
      atomicMin(a.minmax + 2 * b, float_to_ordered(vmin));
      atomicMax(a.minmax + 2 * b + 1, float_to_ordered(vmax));
    }
  }
  __syncthreads();   // tile and reduction scratch are reused by the next tile
  }  // tile loop
}

// minmax[b] = {0xffffffff, 0}
__global__ void minmax_init_kernel(unsigned* mm, int64_t B) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    mm[2 * i] = 0xffffffffu;
    mm[2 * i + 1] = 0u;
  }
}

// S = (L - min) / (max - min) in place; optionally export (min, max) as floats.  One CTA per (signal, row).
__global__ void lognorm_kernel(float* S, int64_t rows, int64_t cols, int64_t ld, const unsigned* mm, float* mm_out) {
  const int64_t b = blockIdx.y;
  const int64_t r = blockIdx.x;
  const float mn = ordered_to_float(mm[2 * b]);
  const float mx = ordered_to_float(mm[2 * b + 1]);
  const float den = mx - mn;
  const float inv = 1.0f / den;
  if (mm_out != nullptr && r == 0 && threadIdx.x == 0) {   // the log image is kept in base 2 (see stft_kernel)
    mm_out[2 * b] = mn * 0.69314718055994531f;
    mm_out[2 * b + 1] = mx * 0.69314718055994531f;
  }
  float* row = S + (b * rows + r) * ld;
  const unsigned n = (unsigned)cols;
  for (unsigned c = threadIdx.x; c < n; c += blockDim.x) row[c] = div_by(row[c] - mn, den, inv);
}

static int stft_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
  }
  return n;
}

template <int LOG2N, int MODE>
static int launch_stft_t(const StftArgs& a, int64_t B, cudaStream_t stream) {
  using C = StftCfg<LOG2N>;
  constexpr int TT = (MODE == STFT_MODE_COMPLEX) ? C::tile_w(8) : C::tile_w(4);
  const StftSmem L = stft_smem_layout<LOG2N>(MODE);
  auto kern = stft_kernel<LOG2N, MODE>;
  if (L.total > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t tiles = ceil_div(a.nseg, TT);
  if (tiles == 0 || B == 0) return 0;
  StftArgs args = a;
  args.tiles_per_signal = tiles;
  args.ntiles = tiles * B;
  if (args.ntiles >= ((int64_t)1 << 31)) return (int)cudaErrorInvalidValue;
  // persistent grid: as many CTAs as can be resident (the kernel is smem/register limited to 1-3 per SM)
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStftThreads, L.total) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t grid = std::min<int64_t>(args.ntiles, (int64_t)per_sm * stft_num_sms());
  SPECGPU_LAUNCH(kern, (unsigned)grid, kStftThreads, L.total, stream, args);
  return (int)cudaGetLastError();
}

template <int MODE>
static int launch_stft_m(int log2n, const StftArgs& a, int64_t B, cudaStream_t stream) {
  switch (log2n) {
    case 3: return launch_stft_t<3, MODE>(a, B, stream);
    case 4: return launch_stft_t<4, MODE>(a, B, stream);
    case 5: return launch_stft_t<5, MODE>(a, B, stream);
    case 6: return launch_stft_t<6, MODE>(a, B, stream);
    case 7: return launch_stft_t<7, MODE>(a, B, stream);
    case 8: return launch_stft_t<8, MODE>(a, B, stream);
    case 9: return launch_stft_t<9, MODE>(a, B, stream);
    case 10: return launch_stft_t<10, MODE>(a, B, stream);
    case 11: return launch_stft_t<11, MODE>(a, B, stream);
    case 12: return launch_stft_t<12, MODE>(a, B, stream);
    case 13: return launch_stft_t<13, MODE>(a, B, stream);
    default: return -1;
  }
}

int launch_stft(int log2n, int mode, const StftArgs& a, int64_t B, cudaStream_t stream) {
  switch (mode) {
    case STFT_MODE_PSD: return launch_stft_m<STFT_MODE_PSD>(log2n, a, B, stream);
    case STFT_MODE_LOGPSD: return launch_stft_m<STFT_MODE_LOGPSD>(log2n, a, B, stream);
    case STFT_MODE_COMPLEX: return launch_stft_m<STFT_MODE_COMPLEX>(log2n, a, B, stream);
    case STFT_MODE_SPECTRA: return launch_stft_m<STFT_MODE_SPECTRA>(log2n, a, B, stream);
    default: return -1;
  }
}

int launch_minmax_init(unsigned* mm, int64_t B, cudaStream_t stream) {
  SPECGPU_LAUNCH(minmax_init_kernel, (unsigned)ceil_div(B, 128), 128, 0, stream, mm, B);
  return (int)cudaGetLastError();
}

int launch_lognorm(float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const unsigned* mm, float* mm_out,
                   cudaStream_t stream) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  SPECGPU_LAUNCH(lognorm_kernel, dim3((unsigned)rows, (unsigned)B), 256, 0, stream, S, rows, cols, ld, mm, mm_out);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
