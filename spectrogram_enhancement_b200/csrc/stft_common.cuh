// Pieces of the STFT kernels shared between stft.cu (the general kernel) and stft_gram.cu (the nperseg-512 log-PSD kernel
// that also accumulates the Gram matrix): tile geometry, shared-memory layout, group reductions, fast log2.
#pragma once
#include "fft.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace specgpu {

constexpr int kStftThreads = 256;
// internal variant of STFT_MODE_LOGPSD: eps >= FLT_MIN, so log2's argument is never subnormal and the
// flush-to-zero MUFU form (no range-scaling instructions) is exact enough
constexpr int STFT_MODE_LOGPSD_FAST = 4;

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
  // tile width (segments per CTA tile): >= NG, grown towards 16 while the tile stays <= 36 KB.  (nperseg 512: 16
  // segments = 64-byte rows of the float tile, one round of the 16 segment groups per tile.)
  __host__ __device__ static constexpr int tile_w(int bytes_per_elem) {
    int tt = NG;
    while (tt < 16 && (long)F * (2 * tt) * bytes_per_elem <= 36 * 1024) tt *= 2;
    return tt;
  }
  static constexpr int LINEH = M + (M >> 4);             // floats per half-width line (see fft_group<.., HALF_LINE>)
};

__host__ __device__ constexpr bool stft_mode_is_log(int mode) { return mode == STFT_MODE_LOGPSD || mode == STFT_MODE_LOGPSD_FAST; }
__host__ __device__ constexpr int stft_elem_bytes(int mode) { return mode == STFT_MODE_COMPLEX ? 8 : 4; }

// The output tile [F rows][TT segments] lives in shared memory as dense rows of ROWB = TT * elem bytes with the
// 16-byte chunks of a row XOR-swizzled by the row index -- exactly the CU_TENSOR_MAP_SWIZZLE_{32,64,128}B patterns
// (address bits 4..6 ^= bits 7..9, masked to the span), so that a TMA tensor store can read it in place while the
// column-wise writes of the segment groups spread over the banks.  The tile base is 1024-byte aligned.
__host__ __device__ constexpr int stft_swizzle_mask(int rowb) { return rowb >= 128 ? 0x70 : (rowb == 64 ? 0x30 : (rowb == 32 ? 0x10 : 0)); }
// (Tried: SWIZZLE_128B for the 64-byte-row tile, which would take the column writes of rows k and k + 8 of a 16-thread
// group off the same bank -- the tensor store then faults: with a 64-byte box row the engine does not read the tile as
// dense 64-byte rows.  The 2-way conflict on the 17 tile stores per thread stays.)

// Log-PSD sizes that can run the barrier-free tile loop (one round of the segment groups per tile, groups inside a
// warp, 16..128-byte tile rows for the tensor store) keep TWO output tiles in shared memory, so that a warp can write
// tile t + 1 while the tensor store of tile t is still reading its buffer and slower warps are still filling it (with one
// buffer 16 % of the kernel's executed instructions were mbarrier polls on "tile free").  The second tile is paid for by
// the half-width exchange line where the transform has two passes.
template <int LOG2N>
__host__ __device__ constexpr bool stft_flow_capable(int mode) {
  using C = StftCfg<LOG2N>;
  const int rowb = C::tile_w(4) * 4;
  return (mode == STFT_MODE_LOGPSD || mode == STFT_MODE_LOGPSD_FAST) && C::tile_w(4) == C::NG && C::G <= 32 && rowb >= 16 && rowb <= 128;
}
// output tiles in shared memory (experiment knob SPECGPU_STFT_TILES: 1 = single tile)
template <int LOG2N>
__host__ __device__ constexpr int stft_num_tiles(int mode) {
#ifdef SPECGPU_STFT_TILES
  return stft_flow_capable<LOG2N>(mode) ? SPECGPU_STFT_TILES : 1;
#else
  return stft_flow_capable<LOG2N>(mode) ? 2 : 1;
#endif
}
template <int LOG2N>
__host__ __device__ constexpr bool stft_half_line(int mode) {
  return stft_flow_capable<LOG2N>(mode) && fft_num_passes(LOG2N - 1) == 2;
}

struct StftSmem {
  int window_off, twm_off, twn_off, line_off, red_off, bar_off, in_off, tile_off, tile_stride, total;
};

// span_floats: samples of the staged input span of one tile, (TT-1)*hop + N (0: segments are loaded straight from
// global memory).
template <int LOG2N>
__host__ __device__ inline StftSmem stft_smem_layout(int mode, int span_floats) {
  using C = StftCfg<LOG2N>;
  StftSmem s;
  int off = 0;
  s.window_off = off; off += C::N * 4;
  s.twm_off = off;    off += (fft_twiddle_count(C::LOG2M) > 0 ? fft_twiddle_count(C::LOG2M) : 1) * 8;
  s.twn_off = off;    off += (C::M / 2 + 1) * 8;
  off = (off + 15) & ~15;
  s.line_off = off;   off += stft_half_line<LOG2N>(mode) ? C::NG * C::LINEH * 4 : C::NG * C::LINE * 8;
  s.red_off = off;    off += (kStftThreads / 32) * 2 * 4 + 64;
  off = (off + 15) & ~15;
  s.bar_off = off;    off += 48 + 64;   // three mbarriers (span full, tile 0 / 1 free) + three arrival counters; tile ring [4][4]
  off = (off + 127) & ~127;
  s.in_off = off;     off += (span_floats * 4 + 127) & ~127;
  s.tile_off = off;
  s.tile_stride = 0;
  if (mode != STFT_MODE_SPECTRA) {
    // the swizzle pattern XORs address bits 4..6 with bits 7..9: a 64-byte-row tile (bits 4, 5 <- 7, 8) needs a 512-byte
    // aligned base, wider rows 1024
    const int rowb = C::tile_w(stft_elem_bytes(mode)) * stft_elem_bytes(mode);
    const int al = rowb <= 64 ? 512 : 1024;
    off = (off + al - 1) & ~(al - 1);
    s.tile_off = off;
    const int tile_bytes = C::F * rowb;
    off += tile_bytes;
    if (stft_num_tiles<LOG2N>(mode) > 1) {
      s.tile_stride = (tile_bytes + al - 1) & ~(al - 1);
      off = s.tile_off + s.tile_stride + tile_bytes;
    }
  }
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

// log2 for arguments known to be normal (>= FLT_MIN): one MUFU, no subnormal range scaling.
__device__ __forceinline__ float log2_normal(float x) {
#if defined(SPECGPU_EMULATE)
  return __log2f(x);
#else
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// Resident CTAs per SM the register allocator should aim for: what the shared-memory footprint of the mode allows
// (the spectra mode has no output tile).  Without it the 3-pass sizes compile to ~195 registers = one CTA per SM.
__host__ __device__ constexpr int stft_min_blocks(int log2n, int mode) {
  if (mode == STFT_MODE_COMPLEX) return log2n <= 10 ? 2 : 1;
#ifdef SPECGPU_STFT_MINBLOCKS9     // experiment knob (tools/build_variant.sh): resident CTAs the nperseg <= 512 kernels are compiled for
  if (log2n <= 9) return SPECGPU_STFT_MINBLOCKS9;
#endif
  if (log2n <= 9) return 3;
  if (mode == STFT_MODE_SPECTRA) return log2n <= 12 ? 3 : 1;
  return log2n == 10 ? 2 : 1;
}

#ifndef SPECGPU_GRID_CONSTANT
#if defined(SPECGPU_EMULATE)
#define SPECGPU_GRID_CONSTANT
#else
#define SPECGPU_GRID_CONSTANT __grid_constant__
#endif
#endif

}  // namespace specgpu
