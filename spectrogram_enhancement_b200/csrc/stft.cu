// K1/K2a: batched overlapped-segment detrend + window + real FFT with fused epilogues.
//
// Replaces scipy.signal.spectrogram / stft / the spectra half of csd as called by the reference
// (spec_denoising/pipeline_data.py:32-35, interferometer/crosspowerspec.py:39).  One CTA handles a
// tile of TT consecutive segments of one signal; groups of G = M/R0 threads each transform one
// segment (M = nperseg/2 complex points, see fft.cuh), untangle the packed real transform into the
// one-sided spectrum, apply the epilogue of the requested mode and stage the result in a shared-
// memory tile so that the [freq][time] output is written in time-contiguous runs.
#include <type_traits>

#include "stft_common.cuh"

namespace specgpu {

// Timing ablations (tools/build_variant.sh; results are WRONG on purpose, never defined in the product build): a bit set in
// SPECGPU_STFT_ABL replaces one class of shared-memory accesses by a thread-dependent register value, so that the
// arithmetic stays and only the LDS / STS traffic goes.  1: window, 2: inter-pass twiddles, 4: tile stores, 8: staged
// input, 16: untangle twiddles.
#ifndef SPECGPU_STFT_ABL
#define SPECGPU_STFT_ABL 0
#endif
// 32: no tensor store of the tile, 64: 16-byte bulk copies instead of the tile's span (barrier protocol unchanged),
// 128: no transform, 256: no block-level fences around the arrival counters
__device__ __forceinline__ uint32_t kAblSpan(int span) { return (SPECGPU_STFT_ABL & 64) ? 16u : (uint32_t)span * 4u; }

template <int LOG2N, int MODE>
__global__ void __launch_bounds__(kStftThreads, stft_min_blocks(LOG2N, MODE))
stft_kernel(const StftArgs a, const SPECGPU_GRID_CONSTANT TensorMap tmap) {
  using C = StftCfg<LOG2N>;
  constexpr int N = C::N, M = C::M, F = C::F, R0 = C::R0, G = C::G, NG = C::NG;
  constexpr bool LOGM = stft_mode_is_log(MODE);
  constexpr int E = stft_elem_bytes(MODE);
  constexpr int TT = C::tile_w(E);
  constexpr int ROUNDS = TT / NG;
  constexpr int ROWB = TT * E;                       // bytes per tile row
  constexpr int SWMASK = stft_swizzle_mask(ROWB);
  constexpr int JSTEP = G * ROWB;                     // tile byte offset between rows k and k + G
  static_assert(JSTEP % 1024 == 0, "row stride between a thread's bins must not touch the swizzle bits");
  constexpr int NTB = stft_num_tiles<LOG2N>(MODE);          // output tiles in shared memory (two on the barrier-free path)
  constexpr bool HALF_LINE = stft_half_line<LOG2N>(MODE);   // half-width exchange line (pays for the second tile)
  SPECGPU_DYN_SMEM(smem);
  const StftSmem L = stft_smem_layout<LOG2N>(MODE, a.stage_in ? a.span : 0);
  float* s_win = reinterpret_cast<float*>(smem + L.window_off);
  float2* s_twm = reinterpret_cast<float2*>(smem + L.twm_off);
  float2* s_twn = reinterpret_cast<float2*>(smem + L.twn_off);
  float2* s_line = reinterpret_cast<float2*>(smem + L.line_off);
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);
  float* s_in = reinterpret_cast<float*>(smem + L.in_off);
  unsigned char* s_tile = smem + L.tile_off;

  const int tid = threadIdx.x;
  const int grp = tid / G;  // which in-flight segment
  const int tg = tid % G;   // thread within the segment group
  const bool stage = a.stage_in != 0;
  const bool tma_out = (MODE != STFT_MODE_SPECTRA) && a.tma_out != 0;

  for (int i = tid; i < N; i += kStftThreads) s_win[i] = a.window[i];
  for (int i = tid; i < fft_twiddle_count(C::LOG2M); i += kStftThreads) s_twm[i] = a.twM[i];
  for (int i = tid; i <= M / 2; i += kStftThreads) s_twn[i] = a.twN[i];
#if !defined(SPECGPU_EMULATE)
  // Barrier-free tile loop (the log-PSD hot path: one round per tile, groups inside a warp, staged bulk input, the whole
  // tile through the tensor store).  No CTA barrier: the LAST warp to hold its samples issues the next bulk copy, the LAST
  // warp to finish its columns issues the tensor store (two shared arrival counters); the only waits are on data (span
  // landed) and on the previous store having read the tile, both normally long satisfied.  Warps of a CTA drift apart,
  // so their shared-memory and FP32 phases overlap instead of colliding.
  constexpr bool FLOWC = stft_flow_capable<LOG2N>(MODE);
  const bool flow = FLOWC && a.flow != 0;
  const uint32_t bar = smem_u32(smem + L.bar_off);
  const uint32_t bar_tfree = bar + 8;                  // [2]: tile buffer p has been read by its tensor store
  unsigned* cnt_free = reinterpret_cast<unsigned*>(smem + L.bar_off + 24);
  unsigned* cnt_full = cnt_free + 1;                   // [2]: warps that have finished their columns of tile buffer p
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_tfree, 1);
    mbar_init(bar_tfree + 8, 1);
    *cnt_free = 0;
    cnt_full[0] = 0;
    cnt_full[1] = 0;
    mbar_fence_init();
    if (tma_out) tma_prefetch_desc(&tmap);
  }
  // the input is read once (both overlapping segments come out of the staged span): do not let it displace the log
  // image, which the Gram and projection kernels read back while it is still in L2
  const uint64_t pol_in = l2_policy_evict_first();
  const uint64_t pol_out = a.l2_pin > 0.f ? l2_policy_pin_fraction(a.l2_pin) : l2_policy_evict_normal();
#else
  constexpr bool flow = false;
#endif
  const unsigned tps = (unsigned)a.tiles_per_signal;      // ntiles < 2^31 is checked by the launcher
  const unsigned ntiles = (unsigned)a.ntiles;
  // Dynamic tile walk (barrier-free path, a.dyn != nullptr): the first tile of a CTA is its index, every further one comes
  // from a global counter, fetched TWO tiles ahead by whichever lane issues the bulk copy of the next tile and published
  // through a small ring in shared memory ({tile, signal, tile within the signal}; the copy's mbarrier orders it).  With
  // the static walk (tile += gridDim.x) a CTA that becomes resident late -- this kernel launched while another stream's
  // Gram / eigen kernels still hold SMs, api.ShotStreams -- finishes late by the same amount while the early ones idle.
  unsigned* s_dyn = reinterpret_cast<unsigned*>(smem + L.bar_off + 48);     // [4][4]
#if !defined(SPECGPU_EMULATE)
  const bool dyn = flow && a.dyn != nullptr;
  auto dyn_publish = [&](unsigned slot, bool fetch) {
    unsigned nn = 0xffffffffu, nb = 0, nt = 0;
    if (fetch) {
      nn = gridDim.x + atomicAdd(a.dyn, 1u);
      nb = nn / tps;
      nt = nn - nb * tps;
    }
    s_dyn[4 * slot] = nn;
    s_dyn[4 * slot + 1] = nb;
    s_dyn[4 * slot + 2] = nt;
  };
  if (dyn && tid == 0) dyn_publish(1u, blockIdx.x < ntiles);
#else
  constexpr bool dyn = false;
  (void)s_dyn;
#endif
  pdl_trigger();     // persistent grid, all CTAs resident: the next kernel may move in as CTAs retire
  pdl_wait();        // the previous kernel of the stream (reader of the image this one overwrites) has completed
  __syncthreads();

  // Can the span of this tile come in as ONE aligned, in-range bulk copy?  (Otherwise -- first/last tiles of a
  // signal, odd alignments -- all threads fill the span with guarded loads, zero outside [0, n).)
  auto tile_bulk = [&](int64_t b, int64_t s0) -> bool {
#if defined(SPECGPU_EMULATE)
    (void)b; (void)s0;
    return false;
#else
    return a.bulk_ok && s0 >= 0 && s0 + a.span <= a.n && (((b * a.ldx + s0) & 3) == 0);
#endif
  };
  // Stage the span of `tile` (all threads call this; only after every thread is done reading the previous span).
  // (signal b32, tile t32 within the signal): kept incrementally along the CTA's tile walk, no division per tile
  auto tile_start = [&](unsigned t32) -> int64_t { return a.first_start + (int64_t)t32 * TT * (int64_t)a.hop; };
  auto prefetch = [&](unsigned b32, unsigned t32) {
    const int64_t b = b32;
    const int64_t s0 = tile_start(t32);
    const float* xb = a.x + b * a.ldx;
    if (tile_bulk(b, s0)) {
#if !defined(SPECGPU_EMULATE)
      if (tid == 0) {
        mbar_arrive_expect_tx(bar, kAblSpan(a.span));
        bulk_g2s(smem_u32(s_in), xb + s0, kAblSpan(a.span), bar, pol_in);
      }
#endif
    } else {
      for (int i = tid; i < a.span; i += kStftThreads) {
        const int64_t idx = s0 + i;
        s_in[i] = (idx >= 0 && idx < a.n) ? __ldg(xb + idx) : 0.f;
      }
    }
  };

  float2* line = HALF_LINE ? s_line + grp * (C::LINEH / 2) : s_line + grp * C::LINE;
  unsigned bulk_parity = 0;
  unsigned flow_it = 0;                                    // tiles this CTA has stored (flow mode)
  unsigned pend = 0;                                       // flow mode, lane 0: 1 + buffer of a tensor store this thread issued
                                                           // and has not yet seen through its shared-memory reads
  unsigned char* const s_tile0 = s_tile;
  unsigned cur_b = blockIdx.x / tps, cur_t = blockIdx.x - cur_b * tps;
  const unsigned step_b = gridDim.x / tps, step_t = gridDim.x - step_b * tps;
  unsigned nxt_b = 0, nxt_t = 0;
  unsigned cur_tile = blockIdx.x, nxt_tile = 0, it = 0;
  if (flow) {
    // bulk tiles are prefetched by one thread; the others (first / last tiles of a signal) are filled at the loop top
    if (blockIdx.x < ntiles && tile_bulk(cur_b, tile_start(cur_t)) && tid == 0) prefetch(cur_b, cur_t);
  } else if (stage && blockIdx.x < ntiles) {
    prefetch(cur_b, cur_t);
  }
  // persistent CTAs: the tables above are loaded once, then the CTA walks tiles (signal b, TT segments)
  for (; cur_tile < ntiles; cur_tile = nxt_tile, cur_b = nxt_b, cur_t = nxt_t, ++it) {
  const int64_t b = cur_b;
  const int64_t seg0 = (int64_t)cur_t * TT;
  if (!dyn) {
    nxt_tile = cur_tile + gridDim.x;
    nxt_b = cur_b + step_b;
    nxt_t = cur_t + step_t;
    if (nxt_t >= tps) {
      nxt_t -= tps;
      ++nxt_b;
    }
  }
  const float* xb = a.x + b * a.ldx;
  float vmin = INFINITY, vmax = -INFINITY;
  const bool bulk = stage && tile_bulk(b, a.first_start + seg0 * (int64_t)a.hop);
  const unsigned tbuf = (NTB > 1 && flow) ? (flow_it % NTB) : 0u;        // output tile buffer of this tile
  s_tile = s_tile0 + tbuf * (unsigned)L.tile_stride;

#if !defined(SPECGPU_EMULATE)
  if (flow) {
    if (bulk) {
      mbar_wait(bar, bulk_parity);
      bulk_parity ^= 1u;
    } else {
      __syncthreads();          // every warp holds its samples of the previous tile
      prefetch(cur_b, cur_t);   // guarded cooperative fill
      __syncthreads();
    }
    if (dyn) {                  // published before this tile's copy was issued (or before the barrier above)
      nxt_tile = s_dyn[4 * ((it + 1u) & 3u)];
      nxt_b = s_dyn[4 * ((it + 1u) & 3u) + 1];
      nxt_t = s_dyn[4 * ((it + 1u) & 3u) + 2];
    }
  } else {
    // the previous tile's tensor store must have finished reading the shared tile before anybody rewrites it; with one
    // round per tile the barrier inside the round orders this wait before the first tile write
    if (tma_out && tid == 0) bulk_wait_read<0>();
    if (bulk) {
      mbar_wait(bar, bulk_parity);
      bulk_parity ^= 1u;
    }
  }
#endif
  if (!flow && (!bulk || (tma_out && ROUNDS > 1))) __syncthreads();   // guarded fill visible / tile free

  for (int round = 0; round < ROUNDS; ++round) {
    const int tl = round * NG + grp;  // column inside the tile
    const int64_t seg = seg0 + tl;
    const bool live = seg < a.nseg;

    // ---- the group's segment: element r of this thread is complex sample m = tg + r*G ----
    float2 v[R0];
    if (stage) {
      const float* p = s_in + (int64_t)tl * a.hop + 2 * tg;
      if (SPECGPU_STFT_ABL & 8) {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = make_float2((float)(tid + r) * 1e-3f, (float)(tid - r) * 1e-3f);
      } else if ((a.hop & 1) == 0) {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = *reinterpret_cast<const float2*>(p + 2 * r * G);
      } else {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = make_float2(p[2 * r * G], p[2 * r * G + 1]);
      }
      if (flow) {
#if !defined(SPECGPU_EMULATE)
        // the last warp to hold its samples issues the bulk copy of the CTA's next tile (non-bulk tiles are filled
        // cooperatively at the top of their own iteration)
        const bool nxt_ok = nxt_tile < ntiles;
        const bool nxt_bulk = nxt_ok && tile_bulk(nxt_b, tile_start(nxt_t));
        if (dyn || nxt_bulk) {
          __syncwarp();
          if ((tid & 31) == 0) {
            if (!(SPECGPU_STFT_ABL & 256)) __threadfence_block();
            if (atomicAdd(cnt_free, 1u) == kStftThreads / 32 - 1) {
              *cnt_free = 0;
              if (dyn) dyn_publish((it + 2u) & 3u, nxt_ok);        // the tile after the next one
              if (!(SPECGPU_STFT_ABL & 256)) __threadfence_block();
              if (nxt_bulk) {
                mbar_arrive_expect_tx(bar, kAblSpan(a.span));
                bulk_g2s(smem_u32(s_in), a.x + (int64_t)nxt_b * a.ldx + tile_start(nxt_t), kAblSpan(a.span), bar, pol_in);
              }
            }
          }
        }
#endif
      } else if (round == ROUNDS - 1) {
        __syncthreads();     // every thread holds its samples: the span may be overwritten
        if (nxt_tile < ntiles) prefetch(nxt_b, nxt_t);
      }
    } else {
      const int64_t s0 = a.first_start + seg * (int64_t)a.hop;
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
    }
    // ---- detrend (scipy.signal.detrend per segment) + window, in packed (x[2m], x[2m+1]) pairs ----
    // sum x and sum c x with c_n = n - (N-1)/2: lane-wise suffix sums T_r = sum_{q >= r} v[q], Q = sum_{r >= 1} T_r =
    // sum_r r v[r], so that sum c x = cb T + 2G Q + T.y (cb = c of this thread's first sample) -- two packed adds per pair.
    if (a.detrend != SPECGPU_DETREND_NONE) {
      const float cb = (float)(2 * tg) - 0.5f * (float)(N - 1);
      float2 T = v[R0 - 1], Q = v[R0 - 1];
#pragma unroll
      for (int r = R0 - 2; r >= 1; --r) {
        T = add2(T, v[r]);
        Q = add2(Q, T);
      }
      T = add2(T, v[0]);
      float sx = T.x + T.y;
      float sc = fmaf(cb, sx, fmaf((float)(2 * G), Q.x + Q.y, T.y));
      group_sum2<G>(sx, sc, s_red, tid);
      const float mean = sx * (1.0f / (float)N);
      // sum_n c_n^2 = N (N^2 - 1) / 12
      const float slope = (a.detrend == SPECGPU_DETREND_LINEAR)
                              ? sc * (12.0f / ((float)N * ((float)N * (float)N - 1.0f)))
                              : 0.f;
      const float t0 = fmaf(slope, cb, mean);
      float2 t = make_float2(t0, t0 + slope);                 // the fitted line at this thread's pair r
      const float2 step = bcast2(slope * (float)(2 * G));
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const float2 w = (SPECGPU_STFT_ABL & 1) ? make_float2(1.f + (float)tg * 1e-3f, 1.f - (float)tg * 1e-3f)
                                                : *reinterpret_cast<const float2*>(s_win + 2 * (tg + r * G));
        v[r] = mul2(sub2(v[r], t), w);
        t = add2(t, step);
      }
    } else {
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * (tg + r * G));
        v[r] = mul2(v[r], w);
      }
    }
    // ---- M-point complex FFT of the packed segment ----
    // Groups that live inside one warp keep the outputs of the last pass in registers and fetch the mirrored bins
    // from their partner thread with shuffles (no second trip through the line); wider groups go through the line.
    constexpr bool REGS = (G <= 32);
    if (!(SPECGPU_STFT_ABL & 128)) fft_group<C::LOG2M, REGS, HALF_LINE, (SPECGPU_STFT_ABL & 2) != 0>(v, line, s_twm, tg);   // 128: no transform

    // ---- untangle: 2 X[k] = E + W_N^k O, 2 X[M-k] = conj(E - W_N^k O) with E = Z[k] + conj(Z[M-k]),
    //      O = -i (Z[k] - conj(Z[M-k])); the factor 2 is folded into the output scales.  Thread tg owns the bin pairs
    //      k = tg + j G, j < R0/2 (thread 0 also k = M/2); the loop is unrolled with the shared-memory offsets of the
    //      twiddles and the tile rows as compile-time strides from per-round bases. ----
    const float cscale = 0.5f * a.scale;             // complex / spectra outputs
    const float pscale1 = 0.25f * a.scale;           // |2X|^2 -> PSD, bins 0 and Nyquist
    const float pscale2 = 0.5f * a.scale;            // one-sided doubling for every other bin
    float2* o2 = nullptr;
    if (MODE == STFT_MODE_SPECTRA) o2 = reinterpret_cast<float2*>(a.out) + (b * a.nseg + seg) * a.ld_out;
    const float fblock_inv = (MODE == STFT_MODE_SPECTRA && a.fblock_w > 0) ? 1.0f / (float)a.fblock_w : 0.f;
    float rmin = INFINITY, rmax = -INFINITY;         // this segment's extremes (merged below if the segment is live)
    auto bin_pair = [&](int k, float2 zk, float2 zm, int offk, int offm, auto generic_c) {
      constexpr bool GENERIC = decltype(generic_c)::value;     // 0 < k < M/2: two distinct, doubled bins
      // packed: e = zk + conj(zm), d = zk - conj(zm) (O = -i d), 2 X[k] = e + W O, 2 conj(X[M-k]) = e - W O with
      // W O = W.x (d.y, -d.x) + W.y (d.x, d.y): six two-lane instructions per bin pair (patterns on the data operand)
      const float2 zc = make_float2(zm.x, -zm.y);
      const float2 e = add2(zk, zc);
      const float2 d = sub2(zk, zc);
      const float2 w = (SPECGPU_STFT_ABL & 16) ? make_float2(1.f - (float)tg * 1e-3f, (float)tg * 1e-3f) : s_twn[k];
      const float2 xk = fma2(d, bcast2(w.y), fma2(make_float2(d.y, -d.x), bcast2(w.x), e));
      float2 xm = fma2(make_float2(-d.x, -d.y), bcast2(w.y), fma2(make_float2(-d.y, d.x), bcast2(w.x), e));
      if (MODE == STFT_MODE_SPECTRA || MODE == STFT_MODE_COMPLEX) xm.y = -xm.y;
      const int km = M - k;
      const bool two = GENERIC || km != k;
      if (MODE == STFT_MODE_SPECTRA) {
        if (live) {
          const int ck = k, cm = km;
          if (a.fblock_w > 0) {     // blocked columns; (k + 0.5) / w in float is exact for these small integers
            const int hk = (int)(((float)k + 0.5f) * fblock_inv), hm = (int)(((float)km + 0.5f) * fblock_inv);
            o2[hk * a.fblock_stride + (k - hk * a.fblock_w)] = make_float2(0.5f * xk.x, 0.5f * xk.y);
            if (two) o2[hm * a.fblock_stride + (km - hm * a.fblock_w)] = make_float2(0.5f * xm.x, 0.5f * xm.y);
          } else {
            o2[ck] = make_float2(0.5f * xk.x, 0.5f * xk.y);
            if (two) o2[cm] = make_float2(0.5f * xm.x, 0.5f * xm.y);
          }
        }
      } else if (MODE == STFT_MODE_COMPLEX) {
        *reinterpret_cast<float2*>(s_tile + offk) = make_float2(xk.x * cscale, xk.y * cscale);
        if (two) *reinterpret_cast<float2*>(s_tile + offm) = make_float2(xm.x * cscale, xm.y * cscale);
      } else {
        // conj(X) X scale, doubled on 1..M-1 (one-sided, even nfft): only k == 0 (bins 0 and M) is not doubled
        const float ps = (GENERIC || k != 0) ? pscale2 : pscale1;
        float pk = xk.x * xk.x + xk.y * xk.y;
        float pm = xm.x * xm.x + xm.y * xm.y;
        if (LOGM) {
          // log2 (one MUFU): the min-max normalisation that follows is invariant to the base of the logarithm, only the
          // exported (min, max) are converted to natural logs.  lg2.approx: absolute error ~1e-6 on values in [-37, 14].
          if (MODE == STFT_MODE_LOGPSD_FAST) {
            pk = log2_normal(fmaf(pk, ps, a.eps));
            pm = log2_normal(fmaf(pm, ps, a.eps));
          } else {
            pk = __log2f(fmaf(pk, ps, a.eps));
            pm = __log2f(fmaf(pm, ps, a.eps));
          }
          rmin = fmin3(rmin, pk, pm);              // k == km (k = M/2) gives pk == pm: harmless
          rmax = fmax3(rmax, pk, pm);
        } else {
          pk *= ps;
          pm *= ps;
        }
        if (!(SPECGPU_STFT_ABL & 4) || pk == 123.456f) {
          *reinterpret_cast<float*>(s_tile + offk) = pk;
          if (two) *reinterpret_cast<float*>(s_tile + offm) = pm;
        }
      }
    };
#if !defined(SPECGPU_EMULATE)
    if (flow) {
      // a store this thread issued one tile ago has long finished reading its buffer: release that buffer now (not
      // right after issuing it, which stalled the last -- slowest -- warp of every tile for the engine's read latency)
      if (pend) {
        bulk_wait_read<0>();
        mbar_arrive(bar_tfree + 8u * (pend - 1u));
        pend = 0;
      }
      // buffer tbuf was last used NTB tiles ago: its store (completion flow_it / NTB - 1 of that barrier) has read it
      if (flow_it >= NTB) mbar_wait(bar_tfree + 8u * tbuf, ((flow_it / NTB) - 1u) & 1u);
    }
#endif
    {
      auto tile_off = [&](int row) {
        const int o = row * ROWB + tl * E;
        return o ^ ((o >> 3) & SWMASK);
      };
      const int offk0 = tile_off(tg);            // rows tg + j G: + j * JSTEP (JSTEP is a multiple of 1024: the
      const int offm0 = tile_off(M - tg);        // swizzle bits do not change); rows M - tg - j G: - j * JSTEP
      constexpr int J = R0 / 2;
      if constexpr (REGS) {
        // Thread tg holds Z[tg + G m] in v[fft_out_reg(m)].  Z[M - (tg + G j)] = Z[(G - tg) + G (R0-1-j)] sits in thread
        // G - tg, register R0-1-j; thread 0 is its own partner with Z[M - G j] = Z[G (R0 - j)] in register (R0 - j) % R0.
        const int partner = ((tid & 31) & ~(G - 1)) | ((G - tg) & (G - 1));
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float2 zk = v[fft_out_reg(C::LOG2M, j)];
          float2 zm = v[fft_out_reg(C::LOG2M, (R0 - j) % R0)];
          if constexpr (G > 1) {
            const float2 snd = v[fft_out_reg(C::LOG2M, R0 - 1 - j)];
            const float rx = __shfl_sync(0xffffffffu, snd.x, partner);
            const float ry = __shfl_sync(0xffffffffu, snd.y, partner);
            if (tg != 0) zm = make_float2(rx, ry);
          }
          if (j == 0) bin_pair(tg, zk, zm, offk0, offm0, std::false_type{});
          else bin_pair(tg + j * G, zk, zm, offk0 + j * JSTEP, offm0 - j * JSTEP, std::true_type{});
        }
        if (tg == 0) {
          const float2 zh = v[fft_out_reg(C::LOG2M, R0 / 2)];
          bin_pair(M / 2, zh, zh, tile_off(M / 2), tile_off(M / 2), std::false_type{});
        }
        if constexpr (G > 1) __syncwarp();   // the line (pass exchanges) is reused by the next round
      } else {
        // j = 0: k = tg; for thread 0 that is the DC / Nyquist pair, whose partner index wraps to 0
        bin_pair(tg, line[fft_pad(tg)], line[fft_pad((M - tg) & (M - 1))], offk0, offm0, std::false_type{});
#pragma unroll
        for (int j = 1; j < J; ++j) {
          const int k = tg + j * G;
          // fft_pad is linear along a thread's bins (G is a multiple of 16 here)
          const int ik = fft_pad(tg) + j * (G + G / 16);
          const int im = fft_pad(M - tg) - j * (G + G / 16);
          bin_pair(k, line[ik], line[im], offk0 + j * JSTEP, offm0 - j * JSTEP, std::true_type{});
        }
        if (tg == 0) bin_pair(M / 2, line[fft_pad(M / 2)], line[fft_pad(M / 2)], tile_off(M / 2), tile_off(M / 2), std::false_type{});
        fft_group_sync<G>();  // line is reused by the next round
      }
    }
    if (LOGM && live) {
      vmin = fminf(vmin, rmin);
      vmax = fmaxf(vmax, rmax);
    }
    if (LOGM && a.ld_out < 0 && !live) {
      // tiled scratch image: the columns of dead segments (past the end of the record, last tile only) are part of the
      // tile the Gram kernel loads -- they must be exact zeros.  Overwrite what this thread stored for them.
      const int ob = tl * E;
      auto zero_row = [&](int row) {
        const int o = row * ROWB + ob;
        *reinterpret_cast<float*>(s_tile + (o ^ ((o >> 3) & SWMASK))) = 0.f;
      };
      for (int j = 0; j < R0 / 2; ++j) {
        zero_row(tg + j * G);
        zero_row(M - tg - j * G);
      }
      if (tg == 0) zero_row(M / 2);
    }
  }

  if (MODE == STFT_MODE_SPECTRA) continue;
#if !defined(SPECGPU_EMULATE)
  if (flow) {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kStftThreads / 32;
    float* red = s_red + (flow_it & 1u) * (2 * NW);      // two sets: a fast warp may be one tile ahead of the reducer
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    fence_proxy_async();       // this thread's tile writes -> visible to the TMA engine (async proxy)
    __syncwarp();
    if (lane == 0) {
      red[2 * warp] = vmin;
      red[2 * warp + 1] = vmax;
      if (!(SPECGPU_STFT_ABL & 256)) __threadfence_block();
      if (atomicAdd(cnt_full + tbuf, 1u) == NW - 1) {    // last warp of the tile: store it
        cnt_full[tbuf] = 0;
        if (!(SPECGPU_STFT_ABL & 256)) __threadfence_block();
        const int rows_out = F - 1;
        for (int rb = 0; rb < ((SPECGPU_STFT_ABL & 32) ? 0 : a.tma_nbox); ++rb) {
          if (a.ld_out < 0)
            tma_store_3d(&tmap, smem_u32(s_tile + (size_t)rb * a.tma_rows * ROWB), (int)(seg0 & 31), (int)(seg0 >> 5) * rows_out + rb * a.tma_rows,
                         (int)b, pol_out);
          else
            tma_store_3d(&tmap, smem_u32(s_tile + (size_t)rb * a.tma_rows * ROWB), (int)(seg0 * (E / 4)), rb * a.tma_rows, (int)b, pol_out);
        }
        bulk_commit();
        float mn = red[0], mx = red[1];
        for (int w = 1; w < NW; ++w) {
          mn = fminf(mn, red[2 * w]);
          mx = fmaxf(mx, red[2 * w + 1]);
        }
        atomicMax(a.minmax + 2 * b, minmax_word_min(a.minmax_gen, mn));
        atomicMax(a.minmax + 2 * b + 1, minmax_word_max(a.minmax_gen, mx));
        pend = tbuf + 1u;
      }
    }
    ++flow_it;
    continue;
  }
#endif
  constexpr int LANES_T = TT < 32 ? TT : 32;   // lanes along time
  constexpr int ROWS_W = 32 / LANES_T;         // rows per warp step
  const int lane = tid & 31, warp = tid >> 5;
  if (LOGM) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) {
      s_red[2 * warp] = vmin;
      s_red[2 * warp + 1] = vmax;
    }
  }
#if !defined(SPECGPU_EMULATE)
  if (tma_out) fence_proxy_async();   // this thread's tile writes -> visible to the TMA engine (async proxy)
#endif
  __syncthreads();   // the tile is complete

  // ---- write the tile: rows = frequency, runs of up to TT consecutive segments ----
  // (the tiled scratch image takes all TT columns: dead segments were zeroed above and the Gram kernel reads whole tiles)
  const int64_t ncol = (a.ld_out >= 0 && a.nseg - seg0 < TT) ? (a.nseg - seg0) : TT;
  const int rows_out = LOGM ? (F - 1) : F;  // Nyquist row dropped after min/max
  int row_first = 0;                        // rows below this leave through the tensor store
#if !defined(SPECGPU_EMULATE)
  if (tma_out) {
    if (tid == 0) {
      for (int rb = 0; rb < a.tma_nbox; ++rb) {
        if (a.ld_out < 0)   // tiled scratch image: tile seg0 / 32, columns seg0 % 32 ... of its [rows x 32] block
          tma_store_3d(&tmap, smem_u32(s_tile + (size_t)rb * a.tma_rows * ROWB), (int)(seg0 & 31), (int)(seg0 >> 5) * rows_out + rb * a.tma_rows,
                       (int)b, pol_out);
        else
          tma_store_3d(&tmap, smem_u32(s_tile + (size_t)rb * a.tma_rows * ROWB), (int)(seg0 * (E / 4)), rb * a.tma_rows, (int)b, pol_out);
      }
      bulk_commit();
    }
    row_first = a.tma_nbox * a.tma_rows;
  }
#endif
  const int lt = lane % LANES_T, lr = lane / LANES_T;
  for (int k = row_first + warp * ROWS_W + lr; k < rows_out; k += (kStftThreads / 32) * ROWS_W) {
    for (int t = lt; t < ncol; t += LANES_T) {
      const int64_t o = img_off(b, k, seg0 + t, rows_out, a.ld_out);
      const int so = k * ROWB + t * E;
      const unsigned char* sp = s_tile + (so ^ ((so >> 3) & SWMASK));
      if (MODE == STFT_MODE_COMPLEX) reinterpret_cast<float2*>(a.out)[o] = *reinterpret_cast<const float2*>(sp);
      else reinterpret_cast<float*>(a.out)[o] = *reinterpret_cast<const float*>(sp);
    }
  }

  if (LOGM && tid == 0) {
    for (int w = 1; w < kStftThreads / 32; ++w) {
      vmin = fminf(vmin, s_red[2 * w]);
      vmax = fmaxf(vmax, s_red[2 * w + 1]);
    }
    atomicMax(a.minmax + 2 * b, minmax_word_min(a.minmax_gen, vmin));
    atomicMax(a.minmax + 2 * b + 1, minmax_word_max(a.minmax_gen, vmax));
  }
  // tile and reduction scratch are reused by the next tile.  With a staged span and a tensor store the barrier inside the
  // next tile's round (and thread 0's wait on the store) already orders both.
  if (!(stage && tma_out && ROUNDS == 1 && G <= 32)) __syncthreads();
  }  // tile loop
#if !defined(SPECGPU_EMULATE)
  if (tma_out && tid == 0 && !flow) bulk_wait<0>();   // the last store must be complete before the CTA's shared memory goes away
  if (flow && pend) bulk_wait_read<0>();              // flow mode: whichever thread issued a store sees it through its reads
  if (dyn && tid == 0) {
    // every fetch of this CTA precedes this point (the lane that fetches for tile i + 2 does so before tile i + 1's copy,
    // which this thread has consumed): the last CTA to leave rewinds the counter for the next launch
    __threadfence();
    if (atomicAdd(a.dyn + 1, 1u) == gridDim.x - 1) {
      a.dyn[0] = 0;
      a.dyn[1] = 0;
      __threadfence();
    }
  }
#endif
}

// S = (L - min) / (max - min) in place; optionally export (min, max) as floats.  One CTA per (signal, row).
__global__ void lognorm_kernel(float* S, int64_t rows, int64_t cols, int64_t ld, const MinMaxWord* mm, float* mm_out) {
  const int64_t b = blockIdx.y;
  const int64_t r = blockIdx.x;
  const float mn = minmax_get_min(mm, b);
  const float mx = minmax_get_max(mm, b);
  const float den = mx - mn;
  const float inv = 1.0f / den;
  if (mm_out != nullptr && r == 0 && threadIdx.x == 0) {   // the log image is kept in base 2 (see stft_kernel)
    mm_out[2 * b] = mn * 0.69314718055994531f;
    mm_out[2 * b + 1] = mx * 0.69314718055994531f;
  }
  float* row = S + (b * rows + r) * ld;
  const unsigned n = (unsigned)cols;
  if (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0)) {
    // 16-byte aligned rows (the pitched images of Runtime.empty_image): whole float4, then the ragged tail
    float4* row4 = reinterpret_cast<float4*>(row);
    const unsigned n4 = n >> 2;
    for (unsigned c = threadIdx.x; c < n4; c += blockDim.x) {
      float4 v = row4[c];
      v.x = div_by(v.x - mn, den, inv);
      v.y = div_by(v.y - mn, den, inv);
      v.z = div_by(v.z - mn, den, inv);
      v.w = div_by(v.w - mn, den, inv);
      row4[c] = v;
    }
    for (unsigned c = 4 * n4 + threadIdx.x; c < n; c += blockDim.x) row[c] = div_by(row[c] - mn, den, inv);
    return;
  }
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

// Largest dynamic shared memory a CTA may ask for (227 KB on sm_100).
static int stft_max_smem() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0) ? v : 227 * 1024;
  }
  return n;
}

template <int LOG2N, int MODE>
static int launch_stft_t(const StftArgs& a, int64_t B, cudaStream_t stream) {
  using C = StftCfg<LOG2N>;
  constexpr int E = stft_elem_bytes(MODE);
  constexpr int TT = C::tile_w(E);
  constexpr int ROWB = TT * E;
  auto kern = stft_kernel<LOG2N, MODE>;
  int64_t tiles = ceil_div(a.nseg, TT);
  if (tiles == 0 || B == 0) return 0;
  // tiled scratch image: cover every 32-column tile completely (an all-dead CTA tile writes the zeros the Gram kernel
  // expects in the columns past the end of the record)
  if (a.ld_out < 0 && TT < kTileCols) tiles = ceil_div(tiles, kTileCols / TT) * (kTileCols / TT);
  StftArgs args = a;
  args.tiles_per_signal = tiles;
  args.ntiles = tiles * B;
  if (args.ntiles >= ((int64_t)1 << 31)) return (int)cudaErrorInvalidValue;
  if (a.ld_out < 0 && (!stft_mode_is_log(MODE) || kTileCols % TT != 0 || -a.ld_out < ceil_div(a.nseg, kTileCols)))
    return (int)cudaErrorInvalidValue;     // the tiled layout is for the log image only, tiles of 32 = whole CTA tiles
  // ---- input staging: the span of a tile ((TT-1) hops + one segment) goes through shared memory when it fits ----
  const int64_t span = (int64_t)(TT - 1) * a.hop + C::N;
  args.span = (int)span;
  args.stage_in = stft_smem_layout<LOG2N>(MODE, (int)span).total <= stft_max_smem() ? 1 : 0;
  args.bulk_ok = (args.stage_in && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.hop % 4 == 0 && a.first_start % 4 == 0 &&
                  (span * 4) % 16 == 0) ? 1 : 0;
  const StftSmem L = stft_smem_layout<LOG2N>(MODE, args.stage_in ? (int)span : 0);
  // ---- output: TMA tensor store of the swizzled shared tile when the layout allows it ----
  TensorMap tmap{};
  args.tma_out = 0;
  args.flow = 0;
  args.tma_rows = args.tma_nbox = 0;
#if !defined(SPECGPU_EMULATE)
  if (MODE != STFT_MODE_SPECTRA && (ROWB == 16 || ROWB == 32 || ROWB == 64 || ROWB == 128)) {
    const int rows_out = stft_mode_is_log(MODE) ? C::F - 1 : C::F;
    const int box_rows = rows_out < 256 ? rows_out : 256;
    const uint64_t w = E / 4;    // floats per element
    bool ok;
    if (a.ld_out < 0) {   // tiled scratch image [B][ntile][rows_out][32]: rows of 128 bytes, a tile's rows contiguous
      const uint64_t nt = (uint64_t)(-a.ld_out);
      ok = make_tensor_map_f32_3d(&tmap, a.out, kTileCols, nt * rows_out, (uint64_t)B, kTileCols, nt * rows_out * kTileCols,
                                  (uint32_t)TT, (uint32_t)box_rows, ROWB >= 32 ? ROWB : 0);
    } else {
      ok = make_tensor_map_f32_3d(&tmap, a.out, (uint64_t)a.nseg * w, (uint64_t)rows_out, (uint64_t)B, (uint64_t)a.ld_out * w,
                                  (uint64_t)rows_out * a.ld_out * w, (uint32_t)(TT * w), (uint32_t)box_rows, ROWB >= 32 ? ROWB : 0);
    }
    if (ok) {
      args.tma_out = 1;
      args.tma_rows = box_rows;
      args.tma_nbox = rows_out / box_rows;
    }
    static const bool flow_env = !(std::getenv("SPECGPU_STFT_FLOW") && std::getenv("SPECGPU_STFT_FLOW")[0] == '0');
    args.flow = (flow_env && ok && args.stage_in && args.bulk_ok && stft_mode_is_log(MODE) && args.tma_nbox * args.tma_rows == rows_out &&
                 TT == C::NG && C::G <= 32) ? 1 : 0;
  }
#endif
  int smem_total = L.total;
  if (const char* pad = std::getenv("SPECGPU_STFT_SMEM_PAD")) smem_total += std::atoi(pad);   // occupancy experiments
  if (smem_total > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
    if (e != cudaSuccess) return (int)e;
  }
  // persistent grid: as many CTAs as can be resident (the kernel is smem/register limited to 1-3 per SM)
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStftThreads, smem_total) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t grid = std::min<int64_t>(args.ntiles, (int64_t)per_sm * stft_num_sms());
  SPECGPU_LAUNCH_PDL(kern, (unsigned)grid, kStftThreads, smem_total, stream, 1, args, tmap);
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
    case STFT_MODE_LOGPSD:
      // eps >= FLT_MIN: log2's argument is never subnormal
      if (a.eps >= 1.17549435e-38f) return launch_stft_m<STFT_MODE_LOGPSD_FAST>(log2n, a, B, stream);
      return launch_stft_m<STFT_MODE_LOGPSD>(log2n, a, B, stream);
    case STFT_MODE_COMPLEX: return launch_stft_m<STFT_MODE_COMPLEX>(log2n, a, B, stream);
    case STFT_MODE_SPECTRA: return launch_stft_m<STFT_MODE_SPECTRA>(log2n, a, B, stream);
    default: return -1;
  }
}

int launch_lognorm(float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const MinMaxWord* mm, float* mm_out,
                   cudaStream_t stream) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  SPECGPU_LAUNCH(lognorm_kernel, dim3((unsigned)rows, (unsigned)B), 256, 0, stream, S, rows, cols, ld, mm, mm_out);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
