// K2b: all-pairs segment-averaged cross-power spectrum,
//   P[i][j][f] = mean_t conj(X_i[t][f]) X_j[t][f] * scale   (x2 on bins 1..nfreq-2),
// i.e. scipy.signal.csd(x_i, x_j, average='mean') for every channel pair -- the arithmetic the
// north-star assigns to `ae_co2` (interferometer/crosspowerspec.py:39).
//
// Input: the unscaled one-sided spectra X[C][nseg][ldf] produced by the STFT kernel (STFT_MODE_SPECTRA).
// The pair matrix is cut into 4x4 channel tiles.  A warp owns one tile for 32 consecutive frequencies
// (lane = frequency): per segment it loads the 4 + 4 spectra values of its tile straight from global memory
// (256-byte coalesced rows; the warps of a CTA work on the same (segment, frequency block), so shared channels
// hit in L1) and accumulates the 16 products in registers over its segments -- the segment averaging never
// leaves the register file.  A CTA holds 16 warps = 16 tiles of one (frequency block, segment chunk); when there
// are fewer tiles than warps (C = 4: one tile) the spare warps take interleaved segments of the same tile and
// the CTA folds them through shared memory.  With all rows requested (i0 = 0, ni = C) only tiles on or above
// the diagonal are computed and the reduce kernel fills the rest by Hermitian symmetry.  Segment-chunk partials
// are summed by a second, deterministic kernel that also applies the scale (and accumulates across blocks of
// segments for the chunked multi-GPU exchange).
#include "kernels.h"

namespace specgpu {

constexpr int kCsdWarps = 16, kCsdThreads = kCsdWarps * 32, kCsdTile = 4;

struct CsdArgs {
  const float2* X;
  int C;
  int64_t nseg, ldf;
  int nfreq;
  int i0, ni;
  int sym;              // only tiles with bi <= bj (requires i0 == 0, ni == C)
  int ntiles, nbj;
  int64_t seg_per_chunk;
  float2* partial;      // [chunk][ni][C][nfreq]
};

__device__ __forceinline__ void csd_tile_coords(const CsdArgs& a, int t, int* bi, int* bj) {
  if (!a.sym) {
    *bi = t / a.nbj;
    *bj = t - *bi * a.nbj;
  } else {
    int i = 0, rem = t;
    while (rem >= a.nbj - i) {
      rem -= a.nbj - i;
      ++i;
    }
    *bi = i;
    *bj = i + rem;
  }
}

__global__ void __launch_bounds__(kCsdThreads) csd_pairs_kernel(CsdArgs a) {
  SPECGPU_DYN_SMEM(smem);
  float* s_fold = reinterpret_cast<float*>(smem);      // [warp][32 accumulators][32 lanes], only when reps > 1
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
  int ci[kCsdTile], cj[kCsdTile];
  bool vi[kCsdTile], vj[kCsdTile];
#pragma unroll
  for (int k = 0; k < kCsdTile; ++k) {
    const int ii = bi * kCsdTile + k, jj = bj * kCsdTile + k;
    vi[k] = ii < a.ni;
    vj[k] = jj < a.C;
    ci[k] = a.i0 + (vi[k] ? ii : 0);
    cj[k] = vj[k] ? jj : 0;
  }
  float2 acc[kCsdTile][kCsdTile];
#pragma unroll
  for (int i = 0; i < kCsdTile; ++i)
#pragma unroll
    for (int j = 0; j < kCsdTile; ++j) acc[i][j] = make_float2(0.f, 0.f);

  if (active && f_ok) {
    const int64_t cstride = a.nseg * a.ldf;
    for (int64_t t = t0 + my_rep; t < t1; t += reps) {
      const float2* xt = a.X + t * a.ldf + f;
      float2 xi[kCsdTile], xj[kCsdTile];
#pragma unroll
      for (int k = 0; k < kCsdTile; ++k) {
        xi[k] = __ldg(xt + ci[k] * cstride);
        xj[k] = __ldg(xt + cj[k] * cstride);
      }
#pragma unroll
      for (int i = 0; i < kCsdTile; ++i)
#pragma unroll
        for (int j = 0; j < kCsdTile; ++j) {
          // conj(xi) * xj
          acc[i][j].x = fmaf(xi[i].x, xj[j].x, fmaf(xi[i].y, xj[j].y, acc[i][j].x));
          acc[i][j].y = fmaf(xi[i].x, xj[j].y, fmaf(-xi[i].y, xj[j].x, acc[i][j].y));
        }
    }
  }
  if (reps > 1) {
    // fold the interleaved-segment warps of every tile into its first warp, in warp order (deterministic)
    if (active && my_rep > 0) {
#pragma unroll
      for (int i = 0; i < kCsdTile; ++i)
#pragma unroll
        for (int j = 0; j < kCsdTile; ++j) {
          s_fold[(warp * 32 + (i * kCsdTile + j) * 2) * 32 + lane] = acc[i][j].x;
          s_fold[(warp * 32 + (i * kCsdTile + j) * 2 + 1) * 32 + lane] = acc[i][j].y;
        }
    }
    __syncthreads();
    if (my_rep == 0) {
      for (int r = 1; r < reps; ++r) {
        const int w = r * tiles_here + my_tile;
#pragma unroll
        for (int i = 0; i < kCsdTile; ++i)
#pragma unroll
          for (int j = 0; j < kCsdTile; ++j) {
            acc[i][j].x += s_fold[(w * 32 + (i * kCsdTile + j) * 2) * 32 + lane];
            acc[i][j].y += s_fold[(w * 32 + (i * kCsdTile + j) * 2 + 1) * 32 + lane];
          }
      }
    }
  }
  if (active && my_rep == 0 && f_ok) {
#pragma unroll
    for (int i = 0; i < kCsdTile; ++i)
#pragma unroll
      for (int j = 0; j < kCsdTile; ++j)
        if (vi[i] && vj[j])
          a.partial[(((int64_t)chunk * a.ni + (bi * kCsdTile + i)) * a.C + (bj * kCsdTile + j)) * a.nfreq + f] = acc[i][j];
  }
}

__global__ void csd_reduce_kernel(const float2* partial, int nchunk, int ni, int C, int nfreq, int sym, float scale,
                                  int accumulate, float2* P) {
  const int64_t count = (int64_t)ni * C * nfreq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % nfreq);
    const int64_t ij = idx / nfreq;
    const int j = (int)(ij % C), i = (int)(ij / C);
    // below-diagonal tiles were not computed in symmetric mode: P[i][j] = conj(P[j][i])
    const bool mirror = sym && (i / kCsdTile > j / kCsdTile);
    const int64_t src = mirror ? (((int64_t)j * C + i) * nfreq + f) : idx;
    float2 s = make_float2(0.f, 0.f);
    for (int c = 0; c < nchunk; ++c) {
      const float2 v = partial[(int64_t)c * count + src];
      s.x += v.x;
      s.y += v.y;
    }
    if (mirror) s.y = -s.y;
    const float sc = (f == 0 || f == nfreq - 1) ? scale : 2.0f * scale;
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
  int sym, nbi, nbj, ntiles, ngroups, nchunk;
  int64_t seg_per_chunk;
};

static CsdGeom csd_geom(int64_t C, int64_t i0, int64_t ni, int nfreq, int64_t nseg) {
  CsdGeom g;
  g.sym = (i0 == 0 && ni == C) ? 1 : 0;
  g.nbi = (int)ceil_div(ni, kCsdTile);
  g.nbj = (int)ceil_div(C, kCsdTile);
  g.ntiles = g.sym ? g.nbj * (g.nbj + 1) / 2 : g.nbi * g.nbj;
  g.ngroups = (int)ceil_div(g.ntiles, kCsdWarps);
  const int64_t ctas = ceil_div(nfreq, 32) * g.ngroups;
  const int64_t want = ceil_div(148 * 2, ctas);                 // ~2 CTAs per SM in flight
  const int64_t maxc = std::max<int64_t>(1, nseg / 32);          // keep >= 32 segments per chunk
  g.nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, maxc), 64));
  g.seg_per_chunk = ceil_div(nseg, g.nchunk);
  g.nchunk = (int)ceil_div(nseg, g.seg_per_chunk);
  return g;
}

size_t csd_pairs_workspace_bytes(int64_t C, int64_t ni, int nfreq, int64_t nseg) {
  const CsdGeom g = csd_geom(C, 0, ni, nfreq, nseg);             // chunk count does not depend on i0 beyond sym
  const CsdGeom g2 = csd_geom(C, 1, ni, nfreq, nseg);
  return (size_t)std::max(g.nchunk, g2.nchunk) * ni * C * nfreq * sizeof(float2) + 256;
}

int launch_csd_pairs(const float* X, int64_t C, int64_t nseg, int64_t nseg_total, int64_t ldf, int nfreq, int64_t i0,
                     int64_t ni, float scale, int accumulate, float* partial_ws, float* P, cudaStream_t stream) {
  if (C == 0 || ni == 0 || nfreq == 0) return 0;
  if (C > 64) return -1;
  const CsdGeom g = csd_geom(C, i0, ni, nfreq, nseg);
  CsdArgs a{};
  a.X = reinterpret_cast<const float2*>(X);
  a.C = (int)C;
  a.nseg = nseg;
  a.ldf = ldf;
  a.nfreq = nfreq;
  a.i0 = (int)i0;
  a.ni = (int)ni;
  a.sym = g.sym;
  a.ntiles = g.ntiles;
  a.nbj = g.nbj;
  a.seg_per_chunk = g.seg_per_chunk;
  a.partial = reinterpret_cast<float2*>(partial_ws);
  if (g.sym) {
    // the mirrored (below-diagonal) tiles are never written: the reduce kernel does not read them either
  }
  const size_t smem = (size_t)kCsdWarps * 32 * 32 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(csd_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  SPECGPU_LAUNCH(csd_pairs_kernel, dim3((unsigned)ceil_div(nfreq, 32), (unsigned)g.nchunk, (unsigned)g.ngroups), kCsdThreads,
                 smem, stream, a);
  int err = (int)cudaGetLastError();
  if (err) return err;
  const int64_t count = ni * C * nfreq;
  SPECGPU_LAUNCH(csd_reduce_kernel, (unsigned)std::min<int64_t>(ceil_div(count, 256), 148 * 8), 256, 0, stream,
                 (const float2*)a.partial, g.nchunk, (int)ni, (int)C, nfreq, g.sym, scale / (float)nseg_total, accumulate,
                 reinterpret_cast<float2*>(P));
  return (int)cudaGetLastError();
}

}  // namespace specgpu
