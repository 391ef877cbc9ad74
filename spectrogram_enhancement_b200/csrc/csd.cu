// K2b: all-pairs segment-averaged cross-power spectrum,
//   P[i][j][f] = mean_t conj(X_i[t][f]) X_j[t][f] * scale   (x2 on bins 1..nfreq-2),
// i.e. scipy.signal.csd(x_i, x_j, average='mean') for every channel pair -- the arithmetic the
// north-star assigns to `ae_co2` (interferometer/crosspowerspec.py:39).
//
// Input: the unscaled one-sided spectra X[C][nseg][ldf] produced by the STFT kernel (STFT_MODE_SPECTRA).
// The pair matrix is cut into 8x4 channel tiles (4x4 for fewer than 8 rows).  A warp owns one tile for 32 consecutive
// frequencies (lane = frequency): per segment it loads the 8 + 4 spectra values of its tile straight from global memory
// (256-byte coalesced rows; the warps of a CTA work on the same (segment, frequency block), so shared channels
// hit in L1) and accumulates the 32 products in registers over its segments -- the segment averaging never
// leaves the register file.  A CTA holds 16 warps = 16 tiles of one (frequency block, segment chunk); when there
// are fewer tiles than warps (C = 4: one tile) the spare warps take interleaved segments of the same tile and
// the CTA folds them through shared memory.  With all rows requested (i0 = 0, ni = C) only tiles on or above
// the diagonal are computed and the reduce kernel fills the rest by Hermitian symmetry.  Segment-chunk partials
// are summed by a second, deterministic kernel that also applies the scale (and accumulates across blocks of
// segments for the chunked multi-GPU exchange).
#include "kernels.h"
#include "ptx.cuh"
#include "tensor_map.h"

namespace specgpu {

#ifndef SPECGPU_GRID_CONSTANT
#if defined(SPECGPU_EMULATE)
#define SPECGPU_GRID_CONSTANT
#else
#define SPECGPU_GRID_CONSTANT __grid_constant__
#endif
#endif

constexpr int kCsdWarps = 16, kCsdThreads = kCsdWarps * 32, kCsdTileJ = 4;   // tiles are TI x 4 pairs, TI in {4, 8}

struct CsdArgs {
  const float2* X;
  int C;
  int64_t nseg, ldf;
  int nfreq;
  int i0, ni;
  int sym;              // only tiles that touch the upper triangle (requires i0 == 0, ni == C)
  int ti;               // tile rows (4 or 8)
  int ntiles, nbj;
  int64_t seg_per_chunk;
  float2* partial;      // [chunk][ni][C][nfreq]
};

// Symmetric mode keeps the tiles (bi, bj) whose last column reaches the first row of the tile: bj >= bi * TI / 4.
// (For every pair at least one of (i, j), (j, i) then lies in a kept tile.)
__device__ __forceinline__ void csd_tile_coords(const CsdArgs& a, int t, int* bi, int* bj) {
  if (!a.sym) {
    *bi = t / a.nbj;
    *bj = t - *bi * a.nbj;
  } else {
    const int skip = a.ti / kCsdTileJ;     // first kept column block of row block i is i * skip
    int i = 0, rem = t;
    while (rem >= a.nbj - i * skip) {
      rem -= a.nbj - i * skip;
      ++i;
    }
    *bi = i;
    *bj = i * skip + rem;
  }
}

template <int TI>
__global__ void __launch_bounds__(kCsdThreads) csd_pairs_kernel(CsdArgs a) {
  constexpr int TJ = kCsdTileJ;
  SPECGPU_DYN_SMEM(smem);
  float* s_fold = reinterpret_cast<float*>(smem);      // [warp][2 TI TJ accumulators][32 lanes], only when reps > 1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x * 32 + lane;
  const bool f_ok = f < a.nfreq;
  const int chunk = blockIdx.y;
  const int tile0 = blockIdx.z * kCsdWarps;
  const int tiles_here = (a.ntiles - tile0 < kCsdWarps) ? (a.ntiles - tile0) : kCsdWarps;
  const int reps = kCsdWarps / tiles_here;              // warps per tile (interleaved segments)
  const int my_tile = warp % tiles_here, my_rep = warp / tiles_here;
  const bool active = my_rep < reps;
  int bi = 0, bj = 0;
  csd_tile_coords(a, tile0 + my_tile, &bi, &bj);
  const int64_t t0 = (int64_t)chunk * a.seg_per_chunk;
  const int64_t t1 = (t0 + a.seg_per_chunk < a.nseg) ? t0 + a.seg_per_chunk : a.nseg;

  // channel indices of the tile (clamped; out-of-range ones contribute zeros)
  int ci[TI], cj[TJ];
  bool vi[TI], vj[TJ];
#pragma unroll
  for (int k = 0; k < TI; ++k) {
    const int ii = bi * TI + k;
    vi[k] = ii < a.ni;
    ci[k] = a.i0 + (vi[k] ? ii : 0);
  }
#pragma unroll
  for (int k = 0; k < TJ; ++k) {
    const int jj = bj * TJ + k;
    vj[k] = jj < a.C;
    cj[k] = vj[k] ? jj : 0;
  }
  float2 acc[TI][TJ];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j] = make_float2(0.f, 0.f);

  if (active && f_ok) {
    const int64_t cstride = a.nseg * a.ldf;
    // UN segments per trip: all their loads are in flight before the first product (this kernel has no staging ring;
    // with one segment per trip a warp waited a full global round trip for every 8-12 loads)
    constexpr int UN = (TI == 4) ? 4 : 1;
    int64_t t = t0 + my_rep;
    for (; t + (UN - 1) * reps < t1; t += UN * reps) {
      float2 xi[UN][TI], xj[UN][TJ];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const float2* xt = a.X + (t + u * reps) * a.ldf + f;
#pragma unroll
        for (int k = 0; k < TI; ++k) xi[u][k] = __ldg(xt + ci[k] * cstride);
#pragma unroll
        for (int k = 0; k < TJ; ++k) xj[u][k] = __ldg(xt + cj[k] * cstride);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
          for (int j = 0; j < TJ; ++j) acc[i][j] = cmac_conj(acc[i][j], xi[u][i], xj[u][j]);     // conj(xi) * xj
    }
    for (; t < t1; t += reps) {
      const float2* xt = a.X + t * a.ldf + f;
      float2 xi[TI], xj[TJ];
#pragma unroll
      for (int k = 0; k < TI; ++k) xi[k] = __ldg(xt + ci[k] * cstride);
#pragma unroll
      for (int k = 0; k < TJ; ++k) xj[k] = __ldg(xt + cj[k] * cstride);
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          // conj(xi) * xj
          acc[i][j] = cmac_conj(acc[i][j], xi[i], xj[j]);
        }
    }
  }
  if (reps > 1) {
    // fold the interleaved-segment warps of every tile into its first warp, in warp order (deterministic)
    if (active && my_rep > 0) {
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          s_fold[(warp * (2 * TI * TJ) + (i * TJ + j) * 2) * 32 + lane] = acc[i][j].x;
          s_fold[(warp * (2 * TI * TJ) + (i * TJ + j) * 2 + 1) * 32 + lane] = acc[i][j].y;
        }
    }
    __syncthreads();
    if (my_rep == 0) {
      for (int r = 1; r < reps; ++r) {
        const int w = r * tiles_here + my_tile;
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
          for (int j = 0; j < TJ; ++j) {
            acc[i][j].x += s_fold[(w * (2 * TI * TJ) + (i * TJ + j) * 2) * 32 + lane];
            acc[i][j].y += s_fold[(w * (2 * TI * TJ) + (i * TJ + j) * 2 + 1) * 32 + lane];
          }
      }
    }
  }
  if (active && my_rep == 0 && f_ok) {
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TJ; ++j)
        if (vi[i] && vj[j])
          a.partial[(((int64_t)chunk * a.ni + (bi * TI + i)) * a.C + (bj * TJ + j)) * a.nfreq + f] = acc[i][j];
  }
}

// ---- staged variant for wide stacks (every warp of the CTA owns its own tile) ---------------------------------------
// The direct-load kernel above makes every warp fetch its 12 operands from global memory and wait for them; with one
// CTA of 16 warps per SM (8x4 tiles need 128 registers) that leaves the FMA pipe a third busy.  Here the CTA streams
// the [C x 32 frequencies] spectra slab of two segments per stage through a cp.async ring in shared memory; all 16
// warps read their operands from the same slab (conflict-free LDS.64, lane = frequency), so global latency is hidden by
// the ring depth instead of by warps, and every spectra value is fetched once per CTA rather than once per warp.

__device__ __forceinline__ void csd_stage_copy(float2* dst, const float2* src, bool valid) {
#if defined(SPECGPU_EMULATE)
  *dst = valid ? *src : make_float2(0.f, 0.f);
#else
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int nbytes = valid ? 8 : 0;                      // src-size 0 zero-fills the destination
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
#endif
}
// 16-byte copy of two bins; only the first `nbytes` (0, 8 or 16) are read, the rest of the destination is zero-filled
__device__ __forceinline__ void csd_stage_copy16(float2* dst, const float2* src, int nbytes) {
#if defined(SPECGPU_EMULATE)
  dst[0] = nbytes >= 8 ? src[0] : make_float2(0.f, 0.f);
  dst[1] = nbytes >= 16 ? src[1] : make_float2(0.f, 0.f);
#else
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
#endif
}
__device__ __forceinline__ void csd_stage_commit() {
#if !defined(SPECGPU_EMULATE)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void csd_stage_wait() {
#if !defined(SPECGPU_EMULATE)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

template <int NST, int SP>   // ring stages, segment pairs per stage
__global__ void __launch_bounds__(kCsdThreads, 1) csd_pairs_staged_kernel(CsdArgs a, int tiles_per_group) {
  constexpr int TI = 8, TJ = kCsdTileJ, kCsdStageSegs = 2 * SP;
  SPECGPU_DYN_SMEM(smem);
  float2* ring = reinterpret_cast<float2*>(smem);        // [NST][2 SP][C][32] + 8 rows of slack (see below)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x * 32 + lane;
  const bool f_ok = f < a.nfreq;
  const int fl = f_ok ? f : a.nfreq - 1;
  const int chunk = blockIdx.y;
  const int tile0 = blockIdx.z * tiles_per_group;
  const int tiles_here = (a.ntiles - tile0 < tiles_per_group) ? (a.ntiles - tile0) : tiles_per_group;
  const bool active = warp < tiles_here;
  int bi = 0, bj = 0;
  csd_tile_coords(a, tile0 + (active ? warp : 0), &bi, &bj);
  const int t0 = (int)((int64_t)chunk * a.seg_per_chunk);
  const int t1 = (int)((t0 + a.seg_per_chunk < a.nseg) ? t0 + a.seg_per_chunk : a.nseg);
  const int nstage = (t1 - t0 + kCsdStageSegs - 1) / kCsdStageSegs;
  const int C = a.C;
  const int seg_elems = C * 32, stage_elems = kCsdStageSegs * seg_elems;

  // The tile's rows are consecutive channels, so its operands sit at compile-time offsets from two bases.  Rows past
  // the stack (ragged last tiles) read up to 7 rows beyond a segment slab - into the next slab or the slack after the
  // ring - and only ever feed accumulators that are not stored.
  const int base_i = (a.i0 + bi * TI) * 32 + lane;
  const int base_j = bj * TJ * 32 + lane;
  float2 acc[TI][TJ];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j] = make_float2(0.f, 0.f);

  // warp w copies channels w, w + 16, ... : both segments of a stage per channel (element offsets fit 32 bits, checked
  // by the launcher)
  const uint32_t cstride = (uint32_t)(a.nseg * a.ldf);
  const uint32_t ldf = (uint32_t)a.ldf;
  // 16-byte copies when the rows allow it: lanes 0-15 carry segment 0 of a channel, lanes 16-31 segment 1
  const bool vec16 = (a.ldf % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0);
  const int half = lane >> 4, l2 = (lane & 15) * 2;
  const int fv = a.nfreq - (blockIdx.x * 32 + l2);         // valid frequencies from this lane's pair onwards
  const int vbytes = fv >= 2 ? 16 : (fv == 1 ? 8 : 0);
  auto load_stage = [&](int st) {
    if (st < nstage) {
      const int tb = t0 + st * kCsdStageSegs;
#pragma unroll
      for (int p = 0; p < SP; ++p) {
        const int ta = tb + 2 * p;                         // segments ta, ta + 1 of this pair
        if (vec16) {
          float2* dst = ring + (st % NST) * stage_elems + (2 * p + half) * seg_elems + l2;
          const bool ok = ta + half < t1;
          const float2* src = a.X + (int64_t)(ok ? ta + half : t0) * a.ldf + (vbytes ? blockIdx.x * 32 + l2 : 0);
          const int nb = ok ? vbytes : 0;
          for (int ch = warp; ch < C; ch += kCsdWarps) csd_stage_copy16(dst + ch * 32, src + ch * cstride, nb);
        } else {
          float2* dst = ring + (st % NST) * stage_elems + 2 * p * seg_elems + lane;
          const bool first = ta < t1, second = ta + 1 < t1;
          const float2* src = a.X + (int64_t)(first ? ta : t0) * a.ldf + fl;
          for (int ch = warp; ch < C; ch += kCsdWarps) {
            const float2* q = src + ch * cstride;
            csd_stage_copy(dst + ch * 32, q, first);
            csd_stage_copy(dst + seg_elems + ch * 32, second ? q + ldf : q, second);
          }
        }
      }
    }
    csd_stage_commit();
  };

#pragma unroll 1
  for (int st = 0; st < NST - 1; ++st) load_stage(st);
#pragma unroll 1
  for (int st = 0; st < nstage; ++st) {
    csd_stage_wait<NST - 2>();
    __syncthreads();                       // stage st has landed for everyone; everyone is done with stage st - 1
    load_stage(st + NST - 1);              // refills the slot stage st - 1 used
    if (active) {
      const float2* slab = ring + (st % NST) * stage_elems;
#pragma unroll
      for (int sg = 0; sg < kCsdStageSegs; ++sg) {
        const float2* si = slab + sg * seg_elems + base_i;
        const float2* sj = slab + sg * seg_elems + base_j;
        float2 xi[TI], xj[TJ];
#pragma unroll
        for (int k = 0; k < TI; ++k) xi[k] = si[k * 32];
#pragma unroll
        for (int k = 0; k < TJ; ++k) xj[k] = sj[k * 32];
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
          for (int j = 0; j < TJ; ++j) {
            acc[i][j] = cmac_conj(acc[i][j], xi[i], xj[j]);
          }
      }
    }
  }
  csd_stage_wait<0>();
  if (active && f_ok) {
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TJ; ++j)
        if (bi * TI + i < a.ni && bj * TJ + j < C)
          a.partial[(((int64_t)chunk * a.ni + (bi * TI + i)) * C + (bj * TJ + j)) * a.nfreq + f] = acc[i][j];
  }
}

// ---- TMA-fed variant of the staged kernel ---------------------------------------------------------------------------
// The cp.async ring above spends a third of its instructions on the copies (per copy a 64-bit address, a predicate, a
// size) -- integer multiply-adds that share the FMA pipe with the products (ncu: FFMA2 48 % of the executed instructions,
// integer / control 30 %, `math_pipe_throttle` the first stall reason).  Here ONE lane of a producer warp issues one
// cp.async.bulk.tensor.3d per stage: the box [C channels][SEGS segments][32 bins] of X[C][nseg][ldf] lands in shared
// memory as [channel][segment][32] float2 behind an mbarrier (bins past nfreq and segments past the record arrive as
// zeros), the 15 consumer warps read their 8 + 4 operands per segment at compile-time offsets from two bases, and hand
// the slot back through a second mbarrier (one arrival per consumer warp): no CTA barrier, no per-thread addressing.
// Segments of the stage that belong to the next chunk are skipped by a warp-uniform test.
#if !defined(SPECGPU_EMULATE)
constexpr int kCsdTmaWarps = 15;                       // consumer warps (= tiles per CTA); warp 15 is the producer
template <int NST, int SEGS>
__global__ void __launch_bounds__(kCsdThreads, 1) csd_pairs_tma_kernel(CsdArgs a, int tiles_per_group, const SPECGPU_GRID_CONSTANT TensorMap tmap) {
  constexpr int TI = 8, TJ = kCsdTileJ;
  SPECGPU_DYN_SMEM(smem);
  const int C = a.C;
  const int stage_elems = C * SEGS * 32;                                   // float2 per stage
  float2* ring = reinterpret_cast<float2*>(smem);                          // [NST][C][SEGS][32] + slack (ragged tiles)
  // barriers behind the ring and its slack (7 channels past the last slab)
  const uint32_t bars = smem_u32(smem) + (uint32_t)(((size_t)NST * stage_elems + 7 * SEGS * 32) * sizeof(float2));
  auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty = [&](int s) { return bars + 8u * (uint32_t)(NST + s); };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.x * 32 + lane;
  const bool f_ok = f < a.nfreq;
  const int chunk = blockIdx.y;
  const int tile0 = blockIdx.z * tiles_per_group;
  const int tiles_here = (a.ntiles - tile0 < tiles_per_group) ? (a.ntiles - tile0) : tiles_per_group;
  const int t0 = (int)((int64_t)chunk * a.seg_per_chunk);
  const int t1 = (int)((t0 + a.seg_per_chunk < a.nseg) ? t0 + a.seg_per_chunk : a.nseg);
  const int nstage = (t1 - t0 + SEGS - 1) / SEGS;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), (uint32_t)tiles_here);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  __syncthreads();
  if (warp == kCsdTmaWarps) {
    // ---- producer ----
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();      // every slab is read once per tile group
      const uint32_t bytes = (uint32_t)(stage_elems * sizeof(float2));
      for (int st = 0; st < nstage; ++st) {
        const int s = st % NST;
        if (st >= NST) mbar_wait(empty(s), ((st / NST) - 1) & 1);          // the consumers are done with the slot's last use
        mbar_arrive_expect_tx(full(s), bytes);
        tma_load_3d(smem_u32(ring + (size_t)s * stage_elems), &tmap, blockIdx.x * 64, t0 + st * SEGS, 0, full(s), pol);
      }
    }
    return;
  }
  if (warp >= tiles_here) return;
  int bi = 0, bj = 0;
  csd_tile_coords(a, tile0 + warp, &bi, &bj);
  // operands of the tile at compile-time offsets from two bases (rows are consecutive channels)
  const int base_i = (a.i0 + bi * TI) * SEGS * 32 + lane;
  const int base_j = bj * TJ * SEGS * 32 + lane;
  float2 acc[TI][TJ];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int st = 0; st < nstage; ++st) {
    const int s = st % NST;
    mbar_wait(full(s), (st / NST) & 1);
    const float2* slab = ring + (size_t)s * stage_elems;
    const int left = t1 - (t0 + st * SEGS);            // segments of this stage that belong to the chunk
#pragma unroll
    for (int sg = 0; sg < SEGS; ++sg) {
      if (sg < left) {
        const float2* si = slab + sg * 32 + base_i;
        const float2* sj = slab + sg * 32 + base_j;
        float2 xi[TI], xj[TJ];
#pragma unroll
        for (int k = 0; k < TI; ++k) xi[k] = si[k * SEGS * 32];
#pragma unroll
        for (int k = 0; k < TJ; ++k) xj[k] = sj[k * SEGS * 32];
#pragma unroll
        for (int i = 0; i < TI; ++i)
#pragma unroll
          for (int j = 0; j < TJ; ++j) acc[i][j] = cmac_conj(acc[i][j], xi[i], xj[j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty(s));
  }
  if (f_ok) {
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TJ; ++j)
        if (bi * TI + i < a.ni && bj * TJ + j < C)
          a.partial[(((int64_t)chunk * a.ni + (bi * TI + i)) * C + (bj * TJ + j)) * a.nfreq + f] = acc[i][j];
  }
}
#endif

// f0 / nfreq_total: the nfreq columns are bins f0 .. f0 + nfreq - 1 of a one-sided spectrum of nfreq_total bins (a
// frequency block of a sharded run); only global bins 0 and nfreq_total - 1 are not doubled.
__global__ void csd_reduce_kernel(const float2* partial, int nchunk, int ni, int C, int nfreq, int sym, int ti, float scale,
                                  int accumulate, int f0, int nfreq_total, float2* P) {
  const int64_t count = (int64_t)ni * C * nfreq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % nfreq);
    const int64_t ij = idx / nfreq;
    const int j = (int)(ij % C), i = (int)(ij / C);
    // tiles entirely below the diagonal were not computed in symmetric mode: P[i][j] = conj(P[j][i])
    const bool mirror = sym && (j / kCsdTileJ < (i / ti) * (ti / kCsdTileJ));
    const int64_t src = mirror ? (((int64_t)j * C + i) * nfreq + f) : idx;
    float2 s = make_float2(0.f, 0.f);
    for (int c = 0; c < nchunk; ++c) {
      const float2 v = partial[(int64_t)c * count + src];
      s.x += v.x;
      s.y += v.y;
    }
    if (mirror) s.y = -s.y;
    const float sc = (f0 + f == 0 || f0 + f == nfreq_total - 1) ? scale : 2.0f * scale;
    float2 r = make_float2(s.x * sc, s.y * sc);
    if (accumulate) {            // a later block of segments of the same average (channel-sharded, chunked exchange)
      const float2 p = P[idx];
      r.x += p.x;
      r.y += p.y;
    }
    P[idx] = r;
  }
}

// Time-resolved cross-power amplitude (the `ampsp[n_time, n_freq]` the interferometer script plots,
// interferometer/crosspowerspec.py:39-50): frame k averages segments [k*seg_stride, k*seg_stride + navg) of the
// record's spectra, amp[k][f] = | mean_t conj(X_i) X_j | * scale (one-sided doubled).  One thread per (frame, bin).
__global__ void csd_frames_kernel(const float2* X, int64_t nseg, int64_t ldf, int nfreq, int ci, int cj, int64_t seg_stride,
                                  int navg, int64_t nframes, float scale, float* amp) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = blockIdx.y;
  if (f >= nfreq || k >= nframes) return;
  const float2* xi = X + ((int64_t)ci * nseg + k * seg_stride) * ldf + f;
  const float2* xj = X + ((int64_t)cj * nseg + k * seg_stride) * ldf + f;
  float re = 0.f, im = 0.f;
  for (int t = 0; t < navg; ++t) {
    const float2 a = __ldg(xi + (int64_t)t * ldf), b = __ldg(xj + (int64_t)t * ldf);
    re = fmaf(a.x, b.x, fmaf(a.y, b.y, re));
    im = fmaf(a.x, b.y, fmaf(-a.y, b.x, im));
  }
  const float sc = ((f == 0 || f == nfreq - 1) ? scale : 2.0f * scale) / (float)navg;
  amp[k * nfreq + f] = sqrtf(re * re + im * im) * sc;
}

int launch_csd_frames(const float* X, int64_t nseg, int64_t ldf, int nfreq, int ci, int cj, int64_t seg_stride, int navg,
                      int64_t nframes, float scale, float* amp, cudaStream_t stream) {
  if (nframes == 0 || nfreq == 0) return 0;
  SPECGPU_LAUNCH(csd_frames_kernel, dim3((unsigned)ceil_div(nfreq, 128), (unsigned)nframes), 128, 0, stream,
                 reinterpret_cast<const float2*>(X), nseg, ldf, nfreq, ci, cj, seg_stride, navg, nframes, scale, amp);
  return (int)cudaGetLastError();
}

struct CsdGeom {
  int sym, ti, nbi, nbj, ntiles, ngroups, nchunk;
  int staged, tiles_per_group;
  int64_t seg_per_chunk;
};

// Chunk count: the smallest one that fills >= 90 % of its last wave of CTA slots, else the best-filling one.
static int csd_pick_chunks(int64_t ctas_per_chunk, int64_t slots, int64_t maxc) {
  int best = 1;
  double best_eff = 0.0;
  for (int64_t k = 1; k <= maxc; ++k) {
    const int64_t total = ctas_per_chunk * k;
    const double eff = (double)total / (double)(ceil_div(total, slots) * slots);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = (int)k;
    }
    if (eff >= 0.9) return (int)k;
  }
  return best;
}

static CsdGeom csd_geom(int64_t C, int64_t i0, int64_t ni, int nfreq, int64_t nseg, int64_t ldf) {
  CsdGeom g;
  g.sym = (i0 == 0 && ni == C) ? 1 : 0;
  g.ti = ni >= 8 ? 8 : 4;       // 8 x 4 tiles halve the bytes loaded per product; small stacks keep 4 x 4
  g.nbi = (int)ceil_div(ni, g.ti);
  g.nbj = (int)ceil_div(C, kCsdTileJ);
  if (g.sym) {
    g.ntiles = 0;
    for (int bi = 0; bi < g.nbi; ++bi) g.ntiles += std::max(0, g.nbj - bi * (g.ti / kCsdTileJ));
  } else {
    g.ntiles = g.nbi * g.nbj;
  }
  g.ngroups = (int)ceil_div(g.ntiles, kCsdWarps);
  g.tiles_per_group = (int)ceil_div(g.ntiles, g.ngroups);        // balanced groups (30 tiles -> 15 + 15)
  // (nearly) every warp owns a tile: stream through smem (the kernel indexes the spectra with 32-bit element offsets)
  g.staged = (g.ti == 8 && g.tiles_per_group >= 12 && C * nseg * ldf < ((int64_t)1 << 31)) ? 1 : 0;
  const int64_t ctas = ceil_div(nfreq, 32) * g.ngroups;
  const int64_t maxc = std::max<int64_t>(1, std::min<int64_t>(nseg / 32, 64));   // keep >= 32 segments per chunk
  if (g.staged) {
    g.nchunk = csd_pick_chunks(ctas, 148, maxc);                 // one 512-thread CTA per SM
  } else {
    // one 512-thread CTA per SM here too (106-128 registers): fill whole waves (4 chords x nperseg 4096: 65 frequency
    // blocks x 5 chunks = 325 CTAs ran as 2.2 waves)
    g.nchunk = csd_pick_chunks(ctas, 148, maxc);
  }
  g.seg_per_chunk = ceil_div(nseg, g.nchunk);
  g.nchunk = (int)ceil_div(nseg, g.seg_per_chunk);
  return g;
}

static int launch_csd_reduce(const CsdArgs& a, const CsdGeom& g, int64_t ni, int64_t C, int nfreq, float scale, int accumulate,
                             int f0, int nfreq_total, float* P, cudaStream_t stream) {
  const int64_t count = ni * C * nfreq;
  SPECGPU_LAUNCH(csd_reduce_kernel, (unsigned)std::min<int64_t>(ceil_div(count, 256), 148 * 8), 256, 0, stream,
                 (const float2*)a.partial, g.nchunk, (int)ni, (int)C, nfreq, g.sym, g.ti, scale, accumulate, f0, nfreq_total,
                 reinterpret_cast<float2*>(P));
  return (int)cudaGetLastError();
}

size_t csd_pairs_workspace_bytes(int64_t C, int64_t ni, int nfreq, int64_t nseg) {
  // the chunk count depends on i0 only through the symmetric shortcut and on the pitch only through the staged switch
  int nchunk = 1;
  for (int64_t i0 = 0; i0 < 2; ++i0)
    for (int64_t ldf : {(int64_t)nfreq, (int64_t)1 << 40}) nchunk = std::max(nchunk, csd_geom(C, i0, ni, nfreq, nseg, ldf).nchunk);
  return (size_t)nchunk * ni * C * nfreq * sizeof(float2) + 256;
}

int launch_csd_pairs(const float* X, int64_t C, int64_t nseg, int64_t nseg_total, int64_t ldf, int nfreq, int64_t i0,
                     int64_t ni, float scale, int accumulate, float* partial_ws, float* P, cudaStream_t stream, int f0,
                     int nfreq_total) {
  if (C == 0 || ni == 0 || nfreq == 0) return 0;
  if (nfreq_total <= 0) nfreq_total = nfreq;
  if (C > 64) return -1;
  const CsdGeom g = csd_geom(C, i0, ni, nfreq, nseg, ldf);
  CsdArgs a{};
  a.X = reinterpret_cast<const float2*>(X);
  a.C = (int)C;
  a.nseg = nseg;
  a.ldf = ldf;
  a.nfreq = nfreq;
  a.i0 = (int)i0;
  a.ni = (int)ni;
  a.sym = g.sym;
  a.ti = g.ti;
  a.ntiles = g.ntiles;
  a.nbj = g.nbj;
  a.seg_per_chunk = g.seg_per_chunk;
  a.partial = reinterpret_cast<float2*>(partial_ws);
#if !defined(SPECGPU_EMULATE)
  static const bool tma_env = !(std::getenv("SPECGPU_CSD_TMA") && std::getenv("SPECGPU_CSD_TMA")[0] == '0');
  if (g.staged && tma_env && g.tiles_per_group <= kCsdTmaWarps) {
    // TMA-fed kernel: box [C][SEGS][32 bins] per stage, as many stages as fit ~200 KB (at least 3)
    constexpr int SEGS = 4;
    const size_t stage_bytes = (size_t)C * SEGS * 32 * sizeof(float2);
    const size_t tail = (size_t)7 * SEGS * 32 * sizeof(float2) + 64;
    const int nst = (4 * stage_bytes + tail <= 208 * 1024) ? 4 : ((3 * stage_bytes + tail <= 208 * 1024) ? 3 : 0);
    TensorMap tmap{};
    if (nst && make_tensor_map_f32_3d(&tmap, X, (uint64_t)nfreq * 2, (uint64_t)nseg, (uint64_t)C, (uint64_t)ldf * 2,
                                      (uint64_t)ldf * 2 * (uint64_t)nseg, 64, SEGS, 0, (uint32_t)C)) {
      const dim3 sgrid((unsigned)ceil_div(nfreq, 32), (unsigned)g.nchunk, (unsigned)g.ngroups);
      const size_t sm = (size_t)nst * stage_bytes + tail;
      auto kern = nst == 4 ? csd_pairs_tma_kernel<4, SEGS> : csd_pairs_tma_kernel<3, SEGS>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) return (int)e;
      SPECGPU_LAUNCH(kern, sgrid, kCsdThreads, sm, stream, a, g.tiles_per_group, tmap);
      int err = (int)cudaGetLastError();
      if (err) return err;
      return launch_csd_reduce(a, g, ni, C, nfreq, scale / (float)nseg_total, accumulate, f0, nfreq_total, P, stream);
    }
  }
#endif
  if (g.staged) {
    const size_t pair_bytes = (size_t)2 * C * 32 * sizeof(float2);      // two segments of the [C x 32] slab
    const dim3 sgrid((unsigned)ceil_div(nfreq, 32), (unsigned)g.nchunk, (unsigned)g.ngroups);
    int err;
    const size_t slack = 8 * 32 * sizeof(float2);          // ragged tiles read up to 7 rows past the last slab
#define SPECGPU_CSD_STAGED(NST, SP)                                                                                       \
  do {                                                                                                                    \
    const size_t sm = (size_t)(NST) * (SP) * pair_bytes + slack;                                                          \
    cudaError_t e = cudaFuncSetAttribute(csd_pairs_staged_kernel<NST, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    if (e != cudaSuccess) return (int)e;                                                                                  \
    SPECGPU_LAUNCH((csd_pairs_staged_kernel<NST, SP>), sgrid, kCsdThreads, sm, stream, a, g.tiles_per_group);        \
  } while (0)
    // measured at 40 channels (B200): 4 stages x 4 segments 0.261 ms, 3 x 4 0.264, 6 x 2 0.277, 8 x 2 0.275
    if (4 * 2 * pair_bytes <= 200 * 1024) SPECGPU_CSD_STAGED(4, 2);
    else if (6 * pair_bytes <= 200 * 1024) SPECGPU_CSD_STAGED(6, 1);
    else SPECGPU_CSD_STAGED(4, 1);
#undef SPECGPU_CSD_STAGED
    if ((err = (int)cudaGetLastError())) return err;
    return launch_csd_reduce(a, g, ni, C, nfreq, scale / (float)nseg_total, accumulate, f0, nfreq_total, P, stream);
  }
  const size_t smem = (size_t)kCsdWarps * (2 * g.ti * kCsdTileJ) * 32 * sizeof(float);
  const dim3 grid((unsigned)ceil_div(nfreq, 32), (unsigned)g.nchunk, (unsigned)g.ngroups);
  if (g.ti == 8) {
    cudaError_t e = cudaFuncSetAttribute(csd_pairs_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(csd_pairs_kernel<8>, grid, kCsdThreads, smem, stream, a);
  } else {
    cudaError_t e = cudaFuncSetAttribute(csd_pairs_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(csd_pairs_kernel<4>, grid, kCsdThreads, smem, stream, a);
  }
  int err = (int)cudaGetLastError();
  if (err) return err;
  return launch_csd_reduce(a, g, ni, C, nfreq, scale / (float)nseg_total, accumulate, f0, nfreq_total, P, stream);
}

}  // namespace specgpu
