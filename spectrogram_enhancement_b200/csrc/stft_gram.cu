// K1 + K2a + K3a in one kernel for the reference defaults (nperseg 512, log-PSD; pipeline_data.py:32-35 feeding
// denoising_by_svd.ipynb:209): the STFT tile that sits in shared memory for its TMA store is ALSO the K-major operand of
// the Gram matrix G = L L^T, so the log image is never re-read for it.
//
// One 768-thread CTA per SM (tensor memory holds one set of accumulators per CTA) run as THREE independent 256-thread
// groups: each group owns its FFT lines, input span, output tile and barriers and walks every third tile of the CTA's
// contiguous tile range exactly like a CTA of stft_kernel, synchronising on its own named barrier.  A 25th warp holds
// three control lanes, one per group: when a group's tile is complete (mbarrier) its control lane (a) sends it through
// the TMA tensor store into the tiled scratch image and (b) issues
// tcgen05.mma.kind::tf32 on it: D1[128 x 256] += tile[0:128] tile^T, D2[128 x 128] += tile[128:256] tile[128:256]^T
// (symmetry: 3/4 of the products) and two 16-column MMAs against a slab of ones for the row sums (the tensor pipe is
// nearly idle here, so unlike in the stand-alone Gram kernel they are free).  All three groups accumulate into the same
// zero-initialised TMEM columns, so the order of their tiles does not matter.  The operands are the RAW log2 values
// (truncated to TF32 by the tensor core); gram_eig_kernel applies the min-max normalisation algebraically (gram_tc.cu).
// A CTA's range touches at most two signals: at the boundary all groups meet, the accumulators are written out as the
// (CTA, segment) partial [128][388] that gram_eig_kernel sums, and are cleared for the next signal.
//
// Not in the emulation build (named barriers, tcgen05): the host only selects this kernel on the device.
#include <cstdlib>
#include <type_traits>

#include "stft_common.cuh"

namespace specgpu {

#if !defined(SPECGPU_EMULATE)

constexpr int kFgSubs = 3, kFgSubThreads = 256, kFgCompute = kFgSubs * kFgSubThreads;
constexpr int kFgThreads = kFgCompute + 32;      // + one control warp: lane s drives the tensor store and the MMAs of group s
constexpr int kFgPW = 384, kFgPP = 388;      // accumulator columns / pitch of a partial row (see gram_tc.cu)

struct FusedArgs {
  StftArgs st;
  float* partial;      // [grid][2][128][388]
  int64_t per;         // tiles per CTA
  int debug;           // timing ablations (SPECGPU_FG_DEBUG): 1 = no MMAs, 2 = no tensor store (results invalid)
};

__device__ __forceinline__ void fg_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kFgSubThreads) : "memory"); }
__device__ __forceinline__ void fg_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fg_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes, 8-row groups 512 bytes apart.
__device__ __forceinline__ uint64_t fg_desc_k_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t fg_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void fg_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void fg_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fg_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 columns of zeros into tensor memory
__device__ __forceinline__ void fg_tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
}

struct FgSmem {
  int window_off, tw_off, twn_off, ones_off, sub_off, sub_stride, line_off, red_off, bar_off, in_off, tile_off, total;
};

__host__ __device__ inline FgSmem fg_smem_layout(int span_floats) {
  using C = StftCfg<9>;
  FgSmem s;
  int off = 0;
  s.window_off = off; off += C::N * 4;
  s.tw_off = off;     off += fft_twiddle_count(C::LOG2M) * 8;
  s.twn_off = off;    off += (C::M / 2 + 1) * 8;
  off = (off + 1023) & ~1023;
  s.ones_off = off;   off += 16 * 64;            // [16 rows x 64 B] of 1.0f (B operand of the row-sum MMAs)
  s.sub_off = off;                               // per-group region (1024-byte aligned)
  int so = 0;
  s.tile_off = so;    so += ((C::F * 64) + 1023) & ~1023;       // [257 rows][16 floats], SWIZZLE_64B
  s.line_off = so;    so += C::NG * C::LINE * 8;
  so = (so + 127) & ~127;
  s.in_off = so;      so += (span_floats * 4 + 127) & ~127;
  s.red_off = so;     so += (kFgSubThreads / 32) * 2 * 4;
  so = (so + 15) & ~15;
  s.bar_off = so;     so += 32;                  // mbarriers: input landed, tile complete, tile free
  s.sub_stride = (so + 1023) & ~1023;
  s.total = s.sub_off + kFgSubs * s.sub_stride;
  return s;
}

template <bool FASTLOG>
__global__ void __launch_bounds__(kFgThreads, 1) stft_gram_kernel(const FusedArgs fa, const __grid_constant__ TensorMap tmap) {
  using C = StftCfg<9>;
  constexpr int N = C::N, M = C::M, F = C::F, R0 = C::R0, G = C::G, NG = C::NG;
  constexpr int TT = 16, E = 4, ROWB = TT * E, SWMASK = stft_swizzle_mask(ROWB), JSTEP = G * ROWB;
  static_assert(C::tile_w(4) == TT && NG == TT && JSTEP % 1024 == 0, "geometry of the nperseg-512 tile");
  const StftArgs& a = fa.st;
  SPECGPU_DYN_SMEM(smem);
  const FgSmem L = fg_smem_layout(a.span);
  float* s_win = reinterpret_cast<float*>(smem + L.window_off);
  float2* s_tw = reinterpret_cast<float2*>(smem + L.tw_off);
  float2* s_twn = reinterpret_cast<float2*>(smem + L.twn_off);
  float* s_ones = reinterpret_cast<float*>(smem + L.ones_off);
  __shared__ uint32_t s_tmem;

  const int tid_all = threadIdx.x;
  const int warp_all = tid_all >> 5;
  const bool control = tid_all >= kFgCompute;                 // the 25th warp
  // compute thread: group = tid_all / 256; control lane s < 3: drives group s
  const int sub = control ? (tid_all - kFgCompute) % kFgSubs : tid_all / kFgSubThreads;
  const int tid = tid_all % kFgSubThreads;       // thread within the group (compute threads)
  const int grp = tid / G, tg = tid % G;
  const int lane = tid_all & 31, warp = tid >> 5;    // warp within the group
  unsigned char* sbase = smem + L.sub_off + sub * L.sub_stride;
  unsigned char* s_tile = sbase + L.tile_off;
  float2* s_line = reinterpret_cast<float2*>(sbase + L.line_off);
  float* s_in = reinterpret_cast<float*>(sbase + L.in_off);
  float* s_red = reinterpret_cast<float*>(sbase + L.red_off);
  const uint32_t bar_in = smem_u32(sbase + L.bar_off);      // the bulk copy of the group's input span has landed
  const uint32_t bar_full = bar_in + 8;                     // the group's tile is complete (one arrival per compute warp)
  const uint32_t bar_free = bar_in + 16;                    // ... has been read by the tensor store and its MMAs have retired

  for (int i = tid_all; i < N; i += kFgThreads) s_win[i] = a.window[i];
  for (int i = tid_all; i < fft_twiddle_count(C::LOG2M); i += kFgThreads) s_tw[i] = a.twM[i];
  for (int i = tid_all; i <= M / 2; i += kFgThreads) s_twn[i] = a.twN[i];
  for (int i = tid_all; i < 16 * 16; i += kFgThreads) s_ones[i] = 1.0f;
  if (!control && tid == 0) {
    mbar_init(bar_in, 1);
    mbar_init(bar_full, kFgSubThreads / 32);
    mbar_init(bar_free, 2);          // tcgen05.commit + the control lane (after the store has read the tile)
    mbar_fence_init();
    if (sub == 0) tma_prefetch_desc(&tmap);
  }
  if (warp_all == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();          // the ones slab is read by the tensor core
  fg_tc_fence_before();
  __syncthreads();
  fg_tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint64_t pol_in = l2_policy_evict_first();
  const uint64_t pol_out = a.l2_pin > 0.f ? l2_policy_pin_fraction(a.l2_pin) : l2_policy_evict_normal();
  const uint32_t idesc1 = fg_idesc_tf32(128, 256), idesc2 = fg_idesc_tf32(128, 128), idesc3 = fg_idesc_tf32(128, 16);

  const int64_t tps = a.tiles_per_signal, total = a.ntiles;
  const int64_t t_first = (int64_t)blockIdx.x * fa.per;
  const int64_t t_last = (t_first + fa.per < total) ? t_first + fa.per : total;     // [t_first, t_last)
  if (t_first >= t_last) {      // uniform: nothing to do, but the allocation must be returned
    __syncthreads();
    if (warp_all == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    return;
  }
  float* part0 = fa.partial + (size_t)blockIdx.x * 2 * 128 * kFgPP;
  float2* line = s_line + grp * C::LINE;
  unsigned in_parity = 0;
  unsigned ntile_done = 0;        // tiles of this group handled so far: tile k completes phase k of bar_full and bar_free

  auto tile_bulk = [&](int64_t b, int64_t s0) -> bool {
    return a.bulk_ok && s0 >= 0 && s0 + a.span <= a.n && (((b * a.ldx + s0) & 3) == 0);
  };
  auto prefetch = [&](int64_t tile) {   // all compute threads of the group; the previous span has been consumed
    const int64_t b = tile / tps;
    const int64_t s0 = a.first_start + (tile - b * tps) * TT * (int64_t)a.hop;
    const float* xb = a.x + b * a.ldx;
    if (tile_bulk(b, s0)) {
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_in, (uint32_t)a.span * 4u);
        bulk_g2s(smem_u32(s_in), xb + s0, (uint32_t)a.span * 4u, bar_in, pol_in);
      }
    } else {
      for (int i = tid; i < a.span; i += kFgSubThreads) {
        const int64_t idx = s0 + i;
        s_in[i] = (idx >= 0 && idx < a.n) ? __ldg(xb + idx) : 0.f;
      }
    }
  };

  int sg = 0;
  for (int64_t lo = t_first; lo < t_last; ++sg) {
    const int64_t b = lo / tps;
    const int64_t hi = ((b + 1) * tps < t_last) ? (b + 1) * tps : t_last;      // this segment: tiles [lo, hi) of signal b
    // ---- clear the accumulators (all groups add into them in any order) ----
    if (!control) {
      const uint32_t q = (uint32_t)(warp_all & 3);
      for (int cg = warp_all >> 2; cg < 13; cg += kFgCompute / 128) fg_tmem_zero32(tmem_base + ((q * 32u) << 16) + (uint32_t)cg * 32u);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fg_tc_fence_before();
    __syncthreads();
    fg_tc_fence_after();

    if (control) {
      // ================= control lanes: lane s sends group s's tiles to the TMA engine and the tensor core =================
      if (tid_all - kFgCompute < kFgSubs) {
        for (int64_t tile = lo + sub; tile < hi; tile += kFgSubs) {
          const int64_t seg0 = (tile - b * tps) * TT;
          mbar_wait(bar_full, ntile_done & 1u);
          fg_tc_fence_after();
          // (a) the tile leaves for the tiled scratch image
          if (!(fa.debug & 2)) tma_store_3d(&tmap, smem_u32(s_tile), (int)(seg0 & 31), (int)(seg0 >> 5) * (F - 1), (int)b, pol_out);
          bulk_commit();
          // (b) and is accumulated into the Gram matrix: K = 16 columns = two K-steps of 8
          const uint32_t tb = smem_u32(s_tile), ob = smem_u32(s_ones);
#pragma unroll
          for (int ks = 0; ks < 2 && !(fa.debug & 1); ++ks) {
            const uint64_t d_lo = fg_desc_k_sw64(tb + ks * 32);
            const uint64_t d_hi = fg_desc_k_sw64(tb + 128 * ROWB + ks * 32);
            const uint64_t d_one = fg_desc_k_sw64(ob + ks * 32);
            fg_umma(tmem_base, d_lo, d_lo, idesc1);
            fg_umma(tmem_base + 256, d_hi, d_hi, idesc2);
            fg_umma(tmem_base + kFgPW, d_lo, d_one, idesc3);
            fg_umma(tmem_base + kFgPW + 16, d_hi, d_one, idesc3);
          }
          fg_commit(bar_free);           // arrival 1 of 2: the MMAs have retired
          bulk_wait_read<0>();           // the store has read the tile
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_free) : "memory");     // arrival 2 of 2
          ++ntile_done;
        }
        // this group's MMAs must have retired before the accumulators are read
        if (ntile_done > 0) mbar_wait(bar_free, (ntile_done - 1) & 1u);
      }
      __syncwarp();
    } else {
      // ================= compute groups =================
      const float* xb = a.x + b * a.ldx;
      (void)xb;
      float vmin = INFINITY, vmax = -INFINITY;
      if (lo + sub < hi) prefetch(lo + sub);
      for (int64_t tile = lo + sub; tile < hi; tile += kFgSubs) {
        const int64_t seg0 = (tile - b * tps) * TT;
        const bool bulk = tile_bulk(b, a.first_start + seg0 * (int64_t)a.hop);
        if (bulk) {
          mbar_wait(bar_in, in_parity);
          in_parity ^= 1u;
        } else {
          fg_bar_sync(1 + sub);      // guarded fill visible
        }
        const int tl = grp;
        const int64_t seg = seg0 + tl;
        const bool live = seg < a.nseg;
        float2 v[R0];
        {
          const float* p = s_in + (int64_t)tl * a.hop + 2 * tg;
          if ((a.hop & 1) == 0) {
#pragma unroll
            for (int r = 0; r < R0; ++r) v[r] = *reinterpret_cast<const float2*>(p + 2 * r * G);
          } else {
#pragma unroll
            for (int r = 0; r < R0; ++r) v[r] = make_float2(p[2 * r * G], p[2 * r * G + 1]);
          }
        }
        fg_bar_sync(1 + sub);        // every thread of the group holds its samples: the span may be overwritten
        if (tile + kFgSubs < hi) prefetch(tile + kFgSubs);
        // ---- detrend ----
        if (a.detrend != SPECGPU_DETREND_NONE) {
          float sx = 0.f, sc = 0.f;
          const float cb = (float)(2 * tg) - 0.5f * (float)(N - 1);
#pragma unroll
          for (int r = 0; r < R0; ++r) {
            const float c0 = cb + (float)(2 * r * G);
            sx += v[r].x + v[r].y;
            sc += c0 * v[r].x + (c0 + 1.0f) * v[r].y;
          }
          group_sum2<G>(sx, sc, s_red, tid);
          const float mean = sx * (1.0f / (float)N);
          const float slope = (a.detrend == SPECGPU_DETREND_LINEAR) ? sc * (12.0f / ((float)N * ((float)N * (float)N - 1.0f))) : 0.f;
#pragma unroll
          for (int r = 0; r < R0; ++r) {
            const float c0 = cb + (float)(2 * r * G);
            v[r].x -= mean + slope * c0;
            v[r].y -= mean + slope * (c0 + 1.0f);
          }
        }
#pragma unroll
        for (int r = 0; r < R0; ++r) {
          const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * (tg + r * G));
          v[r].x *= w.x;
          v[r].y *= w.y;
        }
        fft_group<C::LOG2M, true>(v, line, s_tw, tg);
        // the previous tile of this group must have left the tile buffer (store read + MMAs retired)
        if (ntile_done > 0) mbar_wait(bar_free, (ntile_done - 1) & 1u);
        // ---- untangle + log-PSD into the swizzled tile (see stft_kernel) ----
        const float pscale1 = 0.25f * a.scale, pscale2 = 0.5f * a.scale;
        float rmin = INFINITY, rmax = -INFINITY;
        auto bin_pair = [&](int k, float2 zk, float2 zm, int offk, int offm, auto generic_c) {
          constexpr bool GENERIC = decltype(generic_c)::value;
          const float2 e = make_float2(zk.x + zm.x, zk.y - zm.y);
          const float2 o = make_float2(zk.y + zm.y, zm.x - zk.x);
          const float2 wo = cmul(s_twn[k], o);
          const float2 xk = cadd(e, wo);
          const float2 xm = csub(e, wo);
          const bool two = GENERIC || (M - k) != k;
          const float ps = (GENERIC || k != 0) ? pscale2 : pscale1;
          float pk = xk.x * xk.x + xk.y * xk.y;
          float pm = xm.x * xm.x + xm.y * xm.y;
          if (FASTLOG) {
            pk = log2_normal(fmaf(pk, ps, a.eps));
            pm = log2_normal(fmaf(pm, ps, a.eps));
          } else {
            pk = __log2f(fmaf(pk, ps, a.eps));
            pm = __log2f(fmaf(pm, ps, a.eps));
          }
          rmin = fminf(rmin, fminf(pk, pm));
          rmax = fmaxf(rmax, fmaxf(pk, pm));
          *reinterpret_cast<float*>(s_tile + offk) = pk;
          if (two) *reinterpret_cast<float*>(s_tile + offm) = pm;
        };
        auto tile_off = [&](int row) {
          const int o = row * ROWB + tl * E;
          return o ^ ((o >> 3) & SWMASK);
        };
        {
          const int offk0 = tile_off(tg), offm0 = tile_off(M - tg);
          constexpr int J = R0 / 2;
          const int partner = (lane & ~(G - 1)) | ((G - tg) & (G - 1));
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const float2 zk = v[fft_out_reg(C::LOG2M, j)];
            float2 zm = v[fft_out_reg(C::LOG2M, (R0 - j) % R0)];
            const float2 snd = v[fft_out_reg(C::LOG2M, R0 - 1 - j)];
            const float rx = __shfl_sync(0xffffffffu, snd.x, partner);
            const float ry = __shfl_sync(0xffffffffu, snd.y, partner);
            if (tg != 0) zm = make_float2(rx, ry);
            if (j == 0) bin_pair(tg, zk, zm, offk0, offm0, std::false_type{});
            else bin_pair(tg + j * G, zk, zm, offk0 + j * JSTEP, offm0 - j * JSTEP, std::true_type{});
          }
          if (tg == 0) {
            const float2 zh = v[fft_out_reg(C::LOG2M, R0 / 2)];
            bin_pair(M / 2, zh, zh, tile_off(M / 2), tile_off(M / 2), std::false_type{});
          }
        }
        if (live) {
          vmin = fminf(vmin, rmin);
          vmax = fmaxf(vmax, rmax);
        } else {
          // dead segment (past the end of the record): its column must be exact zeros for the Gram matrix
          auto zero_row = [&](int row) { *reinterpret_cast<float*>(s_tile + tile_off(row)) = 0.f; };
          for (int j = 0; j < R0 / 2; ++j) {
            zero_row(tg + j * G);
            zero_row(M - tg - j * G);
          }
          if (tg == 0) zero_row(M / 2);
        }
        fence_proxy_async();         // tile writes -> visible to the TMA engine and the tensor core
        __syncwarp();                // (also: the FFT line of this warp's groups may be reused)
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full) : "memory");
        ++ntile_done;
      }
      // ---- per-signal min / max of this group's tiles ----
      vmin = warp_min(vmin);
      vmax = warp_max(vmax);
      if (lane == 0) {
        s_red[2 * warp] = vmin;
        s_red[2 * warp + 1] = vmax;
      }
      fg_bar_sync(1 + sub);
      if (tid == 0) {
        for (int w = 1; w < kFgSubThreads / 32; ++w) {
          vmin = fminf(vmin, s_red[2 * w]);
          vmax = fmaxf(vmax, s_red[2 * w + 1]);
        }
        if (vmin <= vmax) {          // (a group without tiles in this segment has nothing to contribute)
          atomicMax(a.minmax + 2 * b, minmax_word_min(a.minmax_gen, vmin));
          atomicMax(a.minmax + 2 * b + 1, minmax_word_max(a.minmax_gen, vmax));
        }
      }
    }
    fg_tc_fence_before();
    __syncthreads();               // every tile of the segment has been accumulated (the control lanes waited for it)
    fg_tc_fence_after();
    // ---- epilogue (the 24 compute warps): TMEM -> partial[sg][128][388]; each lane owns one accumulator row ----
    if (!control) {
      float* part = part0 + (size_t)sg * 128 * kFgPP;
      const int q = warp_all & 3;
      for (int cg = warp_all >> 2; cg < kFgPW / 32; cg += kFgCompute / 128) {
        uint32_t r[32];
        fg_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cg * 32u, r);
        float4* dst = reinterpret_cast<float4*>(part + (size_t)(q * 32 + lane) * kFgPP + cg * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
      }
      if ((warp_all >> 2) == 0) {      // row sums: each of the 16 columns of a group holds the same sum
        uint32_t r[32];
        fg_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kFgPW, r);
        part[(size_t)(q * 32 + lane) * kFgPP + kFgPW] = __uint_as_float(r[0]);
        part[(size_t)(q * 32 + lane) * kFgPP + kFgPW + 1] = __uint_as_float(r[16]);
      }
    }
    fg_tc_fence_before();
    __syncthreads();               // everybody has read the accumulators before they are cleared for the next signal
    fg_tc_fence_after();
    lo = hi;
  }
  if (control && tid_all - kFgCompute < kFgSubs) bulk_wait<0>();    // the last tensor stores must be complete
  __syncwarp();
  __syncthreads();
  if (warp_all == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#endif  // !SPECGPU_EMULATE

bool stft_gram_supported(int log2n) {
#if defined(SPECGPU_EMULATE)
  (void)log2n;
  return false;
#else
  return log2n == 9;
#endif
}

size_t stft_gram_partial_bytes(int64_t B, int64_t nseg, int num_sms) {
  (void)nseg;
  return (size_t)(std::max<int64_t>(B, num_sms) + 8) * 2 * 128 * 388 * sizeof(float) + 256;
}

// STFT (log-PSD into the TILED scratch image `a.out`, a.ld_out = -ntile) + Gram partials of the raw image.
// Returns 0 and the (nchunk, per) geometry gram_eig_kernel needs, 1 if this configuration cannot use the fused kernel.
int launch_stft_gram(const StftArgs& a, int64_t B, float* partial_ws, int num_sms, int64_t* nchunk, int64_t* per,
                     cudaStream_t stream) {
#if defined(SPECGPU_EMULATE)
  (void)a; (void)B; (void)partial_ws; (void)num_sms; (void)nchunk; (void)per; (void)stream;
  return 1;
#else
  using C = StftCfg<9>;
  constexpr int TT = 16;
  if (B == 0 || a.nseg == 0) return 0;
  if (a.ld_out >= 0 || -a.ld_out < ceil_div(a.nseg, kTileCols)) return 1;
  FusedArgs fa{};
  fa.st = a;
  StftArgs& args = fa.st;
  int64_t tiles = ceil_div(a.nseg, TT);
  tiles = ceil_div(tiles, kTileCols / TT) * (kTileCols / TT);        // cover every 32-column tile (dead columns are zeros)
  args.tiles_per_signal = tiles;
  args.ntiles = tiles * B;
  if (args.ntiles >= ((int64_t)1 << 31)) return 1;
  const int64_t span = (int64_t)(TT - 1) * a.hop + C::N;
  args.span = (int)span;
  args.stage_in = 1;
  args.bulk_ok = ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && a.hop % 4 == 0 && a.first_start % 4 == 0 && (span * 4) % 16 == 0) ? 1 : 0;
  const FgSmem L = fg_smem_layout((int)span);
  int max_smem = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (L.total > max_smem) return 1;
  // the tile leaves through a tensor store into [B][ntile][256][32]
  TensorMap tmap{};
  const uint64_t nt = (uint64_t)(-a.ld_out);
  const int rows_out = C::F - 1;
  if (!make_tensor_map_f32_3d(&tmap, a.out, kTileCols, nt * rows_out, (uint64_t)B, kTileCols, nt * rows_out * kTileCols, TT, rows_out, 64))
    return 1;
  args.tma_out = 1;
  args.tma_rows = rows_out;
  args.tma_nbox = 1;
  // contiguous tile ranges, at most two signals per CTA
  fa.per = std::min<int64_t>(std::max<int64_t>(ceil_div(args.ntiles, num_sms), 1), tiles);
  const int64_t grid = ceil_div(args.ntiles, fa.per);
  fa.partial = partial_ws;
  if (const char* env = std::getenv("SPECGPU_FG_DEBUG")) fa.debug = std::atoi(env);
  *nchunk = tiles;
  *per = fa.per;
  const bool fast = a.eps >= 1.17549435e-38f;
  auto kern = fast ? stft_gram_kernel<true> : stft_gram_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (e != cudaSuccess) return (int)e;
  SPECGPU_LAUNCH(kern, (unsigned)grid, kFgThreads, L.total, stream, fa, tmap);
  return (int)cudaGetLastError();
#endif
}

}  // namespace specgpu
