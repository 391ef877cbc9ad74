// K3a: Gram matrix G = S S^T on the 5th-generation tensor cores (tcgen05, TF32 operands, FP32
// accumulators in tensor memory) for rows in {128, 256} -- the one dense contraction of the path.
//
// The batch is cut into 32-column chunks, numbered g = b * nchunk + c; CTA i of a persistent grid owns
// the contiguous range [i*per, (i+1)*per), which touches at most two matrices ("segments").  Warp roles:
//   4 loader warps     cp.async (LDGSTS) the [rows x 32] fp32 slab of a chunk straight into a shared-memory ring stage
//                      in the UMMA canonical K-major SWIZZLE_128B layout (row r at r*128 B, 16-byte chunk c at
//                      c ^ (r & 7)), zero-filling columns past the end, and signal an mbarrier through
//                      cp.async.mbarrier.arrive; they hold nothing in registers and never fence, so the whole ring
//                      (6 x 32 KB) is in flight per SM -- the kernel runs one CTA per SM because of TMEM;
//   12 transform warps apply the min-max normalisation of the log image (one FMA; so the pipeline needs no separate
//                      normalise pass) and round to TF32 like cvt.rna (two integer ops) IN PLACE, fence.proxy.async,
//                      and arrive on a named barrier (an
//                      mbarrier.arrive.release or the proxy fence compile to MEMBAR.ALL.CTA; that is harmless here
//                      because these threads have no global loads in flight);
//   1 issuer warp      one elected thread issues tcgen05.mma.kind::tf32 with BOTH operands described on the same slab:
//     D1[128 x rows] += slab[0:128]   . slab[0:rows]^T     (G00 | G01)
//     D2[128 x 128 ] += slab[128:256] . slab[128:256]^T    (G11; rows == 256 only; G10 = G01^T)
//                      so the symmetric product costs 3/4 of the MMA work; tcgen05.commit releases the stage.
// After the last chunk of a segment all 16 worker warps read the accumulators back with tcgen05.ld, transpose them
// through per-warp shared tiles and write the (CTA, segment) partial as 128-byte rows.  gram_reduce_kernel sums the
// partials of a matrix in a fixed order (deterministic) and mirrors G10.
//
// The emulation build (tests only, no tensor cores on a CPU) replaces the kernel body by a scalar
// loop with the same TF32 operand rounding and the same partial layout.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace specgpu {

#if defined(SPECGPU_EMULATE)
#define SPECGPU_GRID_CONSTANT
#else
#define SPECGPU_GRID_CONSTANT __grid_constant__
#endif

constexpr int kGtcStages = 6;             // 32 KB slabs in the ring (rows = 256): 192 KB of loads in flight per SM
constexpr int kGtcChunk = 32;            // K elements per slab (128 bytes of tf32 per row)
constexpr int kGtcWarps = 16;             // worker warps: loaders + transformers (all of them run the epilogue)
constexpr int kGtcLoadWarps = 4;
constexpr int kGtcThreads = (kGtcWarps + 1) * 32;   // + the MMA issuer warp

struct GramTcArgs {
  const float* S;
  int64_t B, cols, ld;
  int64_t nchunk;           // chunks per matrix
  int64_t per;              // chunks per CTA
  const MinMaxWord* minmax; // optional [B][2] (min, max) words (common.cuh): operands are (x - min) / (max - min)
  int vec16;                // base and pitch allow 16-byte copies
  float l2_pin;             // fraction of S the producer asked L2 to keep (same address-hash policy on these loads); 0: none
  float* partial;           // [grid][2][128][PW] with PW = rows + (rows == 256 ? 128 : 0)
};

// A partial is [128][pitch]: the accumulators (logical width 384 = [G00 | G01 | G11] for 256 rows, `rows` for 128) and, in
// the next two columns, the row sums of rows r and 128 + r (the TMA-fed kernel's raw-operand route; unused otherwise).
__host__ __device__ inline int gram_tc_partial_width(int rows) { return rows == 256 ? 384 : rows; }
__host__ __device__ inline int gram_tc_partial_pitch(int rows) { return gram_tc_partial_width(rows) + 4; }

// Round to TF32 (10 mantissa bits), nearest with ties away from zero -- what cvt.rna.tf32.f32 does -- as two integer
// instructions on the sign-magnitude bit pattern (the cvt expands to about four).
__device__ __forceinline__ float round_tf32(float x) {
  unsigned u = __float_as_uint(x);
  u = (u + 0x1000u) & 0xffffe000u;
  return __uint_as_float(u);
}

#if !defined(SPECGPU_EMULATE)
// ---- PTX wrappers ---------------------------------------------------------------------------------
// Hardware named barriers for the producer -> MMA-issuer hand-off.  An mbarrier.arrive has release semantics and
// compiles to MEMBAR.ALL.CTA, which drains the producers' global loads that are still in flight for the NEXT slabs
// (measured: it serialised every slab on the memory latency); bar.arrive / bar.sync order shared memory like
// __syncthreads() without touching outstanding loads.
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major, 1), SBO >> 4 in
// [32,46) = 1024 B between 8-row groups, version 1 in [46,48), layout SWIZZLE_128B (2) in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (InstrDescriptor): D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10), both K-major,
// N >> 3 @17, M >> 4 @24.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
#endif  // !SPECGPU_EMULATE

template <int ROWS>
__global__ void __launch_bounds__(kGtcThreads, 1) gram_tc_kernel(GramTcArgs a) {
  constexpr int PW = (ROWS == 256) ? 384 : ROWS;   // accumulator columns
  constexpr int PP = PW + 4;                        // row pitch of a partial
  const int tid = threadIdx.x;
  const int64_t total = a.B * a.nchunk;
  const int64_t g0 = (int64_t)blockIdx.x * a.per;
  const int64_t g1 = (g0 + a.per < total) ? g0 + a.per : total;
  if (g0 >= g1) return;                          // uniform over the CTA
  const int64_t b_first = g0 / a.nchunk;
  const int nsegs = (int)((g1 - 1) / a.nchunk - b_first) + 1;     // 1 or 2 (per <= nchunk)
  float* part0 = a.partial + (size_t)blockIdx.x * 2 * 128 * PP;

#if defined(SPECGPU_EMULATE)
  // scalar stand-in with identical operand rounding, work split and partial layout
  for (int sg = 0; sg < nsegs; ++sg) {
    const int64_t b = b_first + sg;
    const int64_t lo = (g0 > b * a.nchunk ? g0 : b * a.nchunk) - b * a.nchunk;
    const int64_t hi = (g1 < (b + 1) * a.nchunk ? g1 : (b + 1) * a.nchunk) - b * a.nchunk;
    const float* Sb = a.S + b * ROWS * a.ld;
    float e_mn = 0.f, e_den = 1.f;
    if (a.minmax != nullptr) {
      e_mn = minmax_get_min(a.minmax, b);
      e_den = minmax_get_max(a.minmax, b) - e_mn;
    }
    for (int i = tid; i < 128 * PW; i += kGtcThreads) {
      const int r = i / PW, c = i % PW;
      const int ra = (c < ROWS) ? r : 128 + r;
      const int rb = (c < ROWS) ? c : c - ROWS + 128;
      float acc = 0.f;
      for (int64_t k = lo * kGtcChunk; k < hi * kGtcChunk && k < a.cols; ++k) {
        float xa = Sb[(int64_t)ra * a.ld + k], xb = Sb[(int64_t)rb * a.ld + k];
        if (a.minmax != nullptr) {       // same single-FMA normalisation as the device path
          xa = fmaf(xa, 1.0f / e_den, -e_mn / e_den);
          xb = fmaf(xb, 1.0f / e_den, -e_mn / e_den);
        }
        acc += round_tf32(xa) * round_tf32(xb);
      }
      part0[(size_t)sg * 128 * PP + (size_t)r * PP + c] = acc;
    }
  }
#else
  SPECGPU_DYN_SMEM(smem);   // SWIZZLE_128B atoms need 1024-byte alignment (re-aligned below)
  constexpr int SLAB = ROWS * 128;  // bytes per stage
  __shared__ __align__(8) uint64_t s_loaded[kGtcStages];   // the loader threads' cp.async copies have landed
  __shared__ __align__(8) uint64_t s_empty[kGtcStages];    // the tensor core has consumed the slab
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (ROWS == 256) ? 512 : 128;
  constexpr int kLoadThreads = kGtcLoadWarps * 32, kXformThreads = (kGtcWarps - kGtcLoadWarps) * 32;
  constexpr int kSegBarrier = 1 + kGtcStages;       // named barrier: every worker warp is done with a segment

  if (tid == 0) {
    for (int i = 0; i < kGtcStages; ++i) {
      mbar_init(smem_u32(&s_loaded[i]), kLoadThreads);
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGtcWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  unsigned char* slabs = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int ncols = (int)a.cols;
  const bool do_norm = a.minmax != nullptr;

  int64_t g = g0;                                   // first chunk of the current segment
  for (int sg = 0; sg < nsegs; ++sg) {
    const int64_t b = b_first + sg;
    const int64_t gend = (g1 < (b + 1) * a.nchunk) ? g1 : (b + 1) * a.nchunk;
    if (warp < kGtcLoadWarps) {
      // ================= loaders: global fp32 -> cp.async -> swizzled shared slab (raw values) =================
      // Nothing is held in registers and these warps never execute a fence (a fence is a MEMBAR.ALL.CTA that would
      // drain the copies in flight), so the whole ring can be in flight.  Row r of a slab lives at
      // r*128 + ((chunk16 ^ (r & 7)) << 4); this thread copies column `lane` of the rows warp, warp + 4, ...
      if (a.vec16) {
        // 16-byte copies (base and pitch are multiples of 4 floats): a warp instruction covers 4 rows x 128 bytes, lane =
        // (row sub-index, 16-byte chunk); a quarter of the copy instructions.  The padded tail of a row may be read
        // (ld >= round_up(cols, 4)); whatever it holds is zeroed by the transform warps.
        const int sub = lane >> 3, c16 = lane & 7;
        const int r0 = 4 * warp + sub;                        // rows r0 + 16 i; (r0 + 16 i) & 7 == r0 & 7
        const uint32_t voff = (uint32_t)(r0 * 128 + ((c16 ^ (r0 & 7)) << 4));
        const int64_t vstride = 16 * a.ld;
        // same address-hash policy as the producer of S: the pinned fraction stays pinned, the rest keeps streaming
        const uint64_t pol = a.l2_pin > 0.f ? l2_policy_pin_fraction(a.l2_pin) : l2_policy_evict_normal();
        for (int64_t gg = g; gg < gend; ++gg) {
          const int64_t ci = gg - g0;
          const int stage = (int)(ci % kGtcStages);
          const uint32_t use = (uint32_t)(ci / kGtcStages);
          if (use > 0) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1);
          const int kcol = (int)(gg - b * a.nchunk) * kGtcChunk + 4 * c16;
          const int nbytes = kcol < ncols ? 16 : 0;
          const float* q = a.S + (b * ROWS + r0) * a.ld + (kcol < ncols ? kcol : 0);
          const uint32_t dst = smem_u32(slabs + stage * SLAB) + voff;
#pragma unroll 8
          for (int i = 0; i < ROWS / 16; ++i, q += vstride)
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst + i * (16 * 128)), "l"(q), "r"(nbytes), "l"(pol) : "memory");
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_loaded[stage])) : "memory");
        }
      } else {
      const uint32_t sw0 = (uint32_t)((((lane >> 2) ^ (warp & 7)) << 4) + ((lane & 3) << 2));
      const uint32_t sw1 = (uint32_t)((((lane >> 2) ^ ((warp + kGtcLoadWarps) & 7)) << 4) + ((lane & 3) << 2));
      const int64_t stride = (int64_t)kGtcLoadWarps * a.ld;
      for (int64_t gg = g; gg < gend; ++gg) {
        const int64_t ci = gg - g0;
        const int stage = (int)(ci % kGtcStages);
        const uint32_t use = (uint32_t)(ci / kGtcStages);
        if (use > 0) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1);
        const int kcol = (int)(gg - b * a.nchunk) * kGtcChunk + lane;
        const int nbytes = kcol < ncols ? 4 : 0;            // src-size 0: the destination is zero-filled
        const float* q = a.S + (b * ROWS + warp) * a.ld + (kcol < ncols ? kcol : 0);
        const uint32_t dst = smem_u32(slabs + stage * SLAB + warp * 128);
#pragma unroll 8
        for (int i = 0; i < ROWS / kGtcLoadWarps; ++i, q += stride)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + i * (kGtcLoadWarps * 128) + ((i & 1) ? sw1 : sw0)),
                       "l"(q), "r"(nbytes)
                       : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_loaded[stage])) : "memory");
      }
      }
    } else if (warp < kGtcWarps) {
      // ================= transform warps: (normalise) + round to TF32, in place =================
      const int t = tid - kLoadThreads;
      float mn = 0.f, den = 1.f;
      if (do_norm) {
        mn = minmax_get_min(a.minmax, b);
        den = minmax_get_max(a.minmax, b) - mn;
      }
      const float nscale = do_norm ? 1.0f / den : 1.0f;        // x -> (x - mn) / den as x * nscale + noff
      const float noff = do_norm ? -mn / den : 0.0f;
      for (int64_t gg = g; gg < gend; ++gg) {
        const int64_t ci = gg - g0;
        const int stage = (int)(ci % kGtcStages);
        const uint32_t use = (uint32_t)(ci / kGtcStages);
        mbar_wait(smem_u32(&s_loaded[stage]), use & 1);
        unsigned char* slab = slabs + stage * SLAB;
        const int k0 = (int)(gg - b * a.nchunk) * kGtcChunk;
        // 16-byte pieces: piece e = (row e / 8, physical chunk e % 8); its logical chunk is (e % 8) ^ (row & 7).
        // The operands are rounded to 10 mantissa bits, so the normalisation is a single FMA here (the exactly
        // rounded division is kept for the image the pipeline writes, not for the Gram operands).
        if (k0 + kGtcChunk <= ncols) {
#pragma unroll 2
          for (int e = t; e < ROWS * 8; e += kXformThreads) {
            float4 x = *reinterpret_cast<float4*>(slab + e * 16);
            x.x = round_tf32(fmaf(x.x, nscale, noff));
            x.y = round_tf32(fmaf(x.y, nscale, noff));
            x.z = round_tf32(fmaf(x.z, nscale, noff));
            x.w = round_tf32(fmaf(x.w, nscale, noff));
            *reinterpret_cast<float4*>(slab + e * 16) = x;
          }
        } else {   // last chunk of a matrix: columns past the end were zero-filled and must stay zero
          for (int e = t; e < ROWS * 8; e += kXformThreads) {
            const int row = e >> 3, pc = e & 7;
            const int k = k0 + ((pc ^ (row & 7)) << 2);
            float4 x = *reinterpret_cast<float4*>(slab + e * 16);
            x.x = (k + 0 < ncols) ? round_tf32(fmaf(x.x, nscale, noff)) : 0.f;
            x.y = (k + 1 < ncols) ? round_tf32(fmaf(x.y, nscale, noff)) : 0.f;
            x.z = (k + 2 < ncols) ? round_tf32(fmaf(x.z, nscale, noff)) : 0.f;
            x.w = (k + 3 < ncols) ? round_tf32(fmaf(x.w, nscale, noff)) : 0.f;
            *reinterpret_cast<float4*>(slab + e * 16) = x;
          }
        }
        fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
        named_bar_arrive(1 + stage, kXformThreads + 32);
      }
    } else {
      // ================= MMA issuer warp: all lanes join the named barrier, lane 0 issues =================
      const uint32_t idesc1 = umma_idesc_tf32(128, ROWS);
      const uint32_t idesc2 = umma_idesc_tf32(128, 128);
      for (int64_t gg = g; gg < gend; ++gg) {
        const int64_t ci = gg - g0;
        const int stage = (int)(ci % kGtcStages);
        named_bar_sync(1 + stage, kXformThreads + 32);      // the transform warps have rewritten and fenced this slab
        if (lane == 0) {
          tc_fence_after();
          const uint32_t base = smem_u32(slabs + stage * SLAB);
#pragma unroll
          for (int ks = 0; ks < kGtcChunk / 8; ++ks) {   // K = 8 tf32 (32 bytes) per instruction
            const uint32_t accumulate = (gg > g || ks > 0) ? 1u : 0u;
            const uint64_t d_lo = umma_desc_k_sw128(base + ks * 32);
            umma_tf32(tmem_base, d_lo, d_lo, idesc1, accumulate);
            if (ROWS == 256) {
              const uint64_t d_hi = umma_desc_k_sw128(base + 128 * 128 + ks * 32);
              umma_tf32(tmem_base + 256, d_hi, d_hi, idesc2, accumulate);
            }
          }
          umma_commit(smem_u32(&s_empty[stage]));   // slab may be refilled once these MMAs retire
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(smem_u32(&s_accum));            // accumulators of this segment complete
      __syncwarp();
    }
    if (warp < kGtcWarps) {
      // ================= epilogue (all worker warps): TMEM -> registers -> partial[sg][128][PW] =================
      mbar_wait(smem_u32(&s_accum), (uint32_t)sg & 1);
      tc_fence_after();
      float* part = part0 + (size_t)sg * 128 * PP;
      const int q = warp & 3;                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
      constexpr int NCG = PW / 32;            // 32-column groups, dealt round-robin to the warps of a quarter
      // tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns); a per-warp [32][33] tile transposes it so
      // that the partial is written as 128-byte rows.  The tiles live in the slab ring: every MMA of the segment has
      // retired and the loaders (which run this epilogue too) have not started on the next segment.
      float* tr = reinterpret_cast<float*>(slabs) + warp * (32 * 33);
      for (int cg = warp >> 2; cg < NCG; cg += kGtcWarps / 4) {
        const int c = cg * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        float* dst = part + (size_t)(q * 32) * PP + c + lane;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) dst[(size_t)r * PP] = tr[r * 33 + lane];
        __syncwarp();
      }
      tc_fence_before();
      // every worker warp has read its part of TMEM and is done with its transpose tile before the next segment's
      // copies land in the ring and its first MMA overwrites the accumulators
      if (sg + 1 < nsegs) named_bar_sync(kSegBarrier, kGtcWarps * 32);
    }
    g = gend;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGtcWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
#endif
}

// ------------------------------------------------------------------------------------------------------
// TMA-fed variant (row-pitched images: 16-byte aligned rows).  Nobody touches the operands in registers or rewrites
// them in shared memory: one thread issues cp.async.bulk.tensor loads of [rows x 32] boxes straight into the ring
// in the SWIZZLE_128B K-major layout (columns past the end are zero-filled by the TMA unit), one thread issues the
// MMAs on the RAW fp32 bits (kind::tf32 reads the upper 19 bits), and the min-max normalisation is applied
// algebraically afterwards: with r = rowsums and T columns,
//     (L - m 1 1^T)(L - m 1 1^T)^T = L L^T - m (r 1^T + 1 r^T) + m^2 T 1 1^T,
// where the row sums come out of the same pipeline as two 16-column MMAs against a slab of ones (so the identity
// holds exactly for the truncated operands).  gram_eig_kernel applies the correction while it sums the partials.
// ------------------------------------------------------------------------------------------------------
// A stage of the TMA-fed ring holds kGtmChunk = 64 columns as two [rows x 32] boxes: a row then contributes 256
// contiguous bytes per visit instead of 128 (the images' rows are 15.7 KB apart, and DRAM pages like longer bursts).
constexpr int kGtmChunk = 64;
constexpr int kGtmStages = 3;

struct GramTmaArgs {
  int64_t B, cols, nchunk, per;
  int tiled;               // S is the tiled scratch image (common.cuh): tile t of matrix b is the box at row t * rows
  int debug;               // timing ablations (SPECGPU_GRAM_DEBUG): 1 = no MMAs, 2 = no epilogue stores (results invalid)
  float l2_pin;
  float* partial;
};


template <int ROWS>
__global__ void __launch_bounds__(kGtcThreads, 1) gram_tma_kernel(const GramTmaArgs a, const float* S, int64_t ld,
                                                                  const SPECGPU_GRID_CONSTANT TensorMap tmap) {
  constexpr int PW = (ROWS == 256) ? 384 : ROWS;
  constexpr int PP = PW + 4;
  const int tid = threadIdx.x;
  // CTA i = part (i % k) of matrix i / k, k = CTAs per matrix: chunks [part * per, min((part + 1) * per, nchunk)) of that
  // matrix.  Expressed in the global chunk numbering g = b * nchunk + c the loops below share with the cp.async kernel.
  const int64_t kparts = (a.nchunk + a.per - 1) / a.per;
  const int64_t b_first = blockIdx.x / kparts;
  const int64_t g0 = b_first * a.nchunk + (blockIdx.x % kparts) * a.per;
  const int64_t g1 = (g0 + a.per < (b_first + 1) * a.nchunk) ? g0 + a.per : (b_first + 1) * a.nchunk;
  pdl_trigger();                                 // one resident wave (common.cuh: programmatic dependent launch)
  if (g0 >= g1) {                                // uniform over the CTA
    pdl_wait();
    return;
  }
  constexpr int nsegs = 1;
  float* part0 = a.partial + (size_t)blockIdx.x * 128 * PP;
#if defined(SPECGPU_EMULATE)
  // scalar stand-in: truncated-to-TF32 raw operands, same work split and partial layout (row sums in columns PW, PW+1)
  auto trunc = [](float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); };
  for (int sg = 0; sg < nsegs; ++sg) {
    const int64_t b = b_first + sg;
    const int64_t lo = (g0 > b * a.nchunk ? g0 : b * a.nchunk) - b * a.nchunk;
    const int64_t hi = (g1 < (b + 1) * a.nchunk ? g1 : (b + 1) * a.nchunk) - b * a.nchunk;
    auto at = [&](int r, int64_t k) { return S[img_off(b, r, k, ROWS, ld)]; };
    for (int i = tid; i < 128 * (PW + 2); i += kGtcThreads) {
      const int r = i / (PW + 2), c = i % (PW + 2);
      float acc = 0.f;
      if (c < PW) {
        const int ra = (c < ROWS) ? r : 128 + r;
        const int rb = (c < ROWS) ? c : c - ROWS + 128;
        for (int64_t k = lo * kGtmChunk; k < hi * kGtmChunk && k < a.cols; ++k)
          acc += trunc(at(ra, k)) * trunc(at(rb, k));
      } else if (c == PW || ROWS == 256) {
        const int ra = (c == PW) ? r : 128 + r;
        for (int64_t k = lo * kGtmChunk; k < hi * kGtmChunk && k < a.cols; ++k) acc += trunc(at(ra, k));
      }
      part0[(size_t)sg * 128 * PP + (size_t)r * PP + c] = acc;
    }
  }
#else
  (void)S;
  (void)ld;
  SPECGPU_DYN_SMEM(smem);   // SWIZZLE_128B atoms need 1024-byte alignment (re-aligned below)
  constexpr int BOX = ROWS * 128;   // bytes of one [rows x 32] box
  constexpr int SLAB = 2 * BOX;     // bytes per stage (two boxes)
  constexpr int kGtcStages = kGtmStages;   // (shadows the cp.async kernel's ring depth)
  __shared__ __align__(8) uint64_t s_full[kGtcStages];     // the TMA boxes of this stage have landed
  __shared__ __align__(8) uint64_t s_empty[kGtcStages];    // the tensor core has consumed the slab
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (ROWS == 256) ? 512 : 128;
  constexpr int kSegBarrier = 1;                    // named barrier: every worker warp is done with a segment
  // Row sums of the (truncated) operands for the algebraic normalisation: the worker warps have nothing to do during the
  // main loop, so warps 1 .. ROWS/32 read every landed stage back from shared memory, one thread per image row
  // (MMAs against a slab of ones did the same on the tensor pipe but cost as much as a third of the Gram MMAs).
  constexpr int kRsWarps = ROWS / 32;
  __shared__ float s_rowsum[ROWS];

  if (tid == 0) {
    for (int i = 0; i < kGtcStages; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), 1 + kRsWarps);      // the MMA commit + one arrival per row-sum warp
    }
    mbar_init(smem_u32(&s_accum), 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == kGtcWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  unsigned char* slabs = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_wait();      // barriers, tensor memory and the descriptor prefetch overlapped the tail of the STFT; its image is now complete

  int64_t g = g0;                                   // first chunk of the current segment
  for (int sg = 0; sg < nsegs; ++sg) {
    const int64_t b = b_first + sg;
    const int64_t gend = (g1 < (b + 1) * a.nchunk) ? g1 : (b + 1) * a.nchunk;
    if (warp == 0 && lane == 0) {
      // ================= producer: one TMA box per chunk =================
      const uint64_t pol = a.l2_pin > 0.f ? l2_policy_pin_fraction(a.l2_pin) : l2_policy_evict_normal();
      for (int64_t gg = g; gg < gend; ++gg) {
        const int64_t ci = gg - g0;
        const int stage = (int)(ci % kGtcStages);
        const uint32_t use = (uint32_t)(ci / kGtcStages);
        if (use > 0) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1);
        if (a.debug & 4) {     // ablation: no loads (the MMAs run on whatever the ring holds)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[stage])) : "memory");
          continue;
        }
        mbar_arrive_expect_tx(smem_u32(&s_full[stage]), SLAB);
        const int kcol = (int)(gg - b * a.nchunk) * kGtmChunk;
        if (a.tiled) {    // two contiguous [rows x 32] tiles (a tile past the last one is out of bounds: zero-filled)
          const int t0 = kcol >> 5;
          tma_load_3d(smem_u32(slabs + stage * SLAB), &tmap, 0, t0 * ROWS, (int)b, smem_u32(&s_full[stage]), pol);
          tma_load_3d(smem_u32(slabs + stage * SLAB + BOX), &tmap, 0, (t0 + 1) * ROWS, (int)b, smem_u32(&s_full[stage]), pol);
        } else {
          tma_load_3d(smem_u32(slabs + stage * SLAB), &tmap, kcol, 0, (int)b, smem_u32(&s_full[stage]), pol);
          tma_load_3d(smem_u32(slabs + stage * SLAB + BOX), &tmap, kcol + 32, 0, (int)b, smem_u32(&s_full[stage]), pol);
        }
      }
    } else if (warp == kGtcWarps) {
      // ================= MMA issuer warp: lane 0 issues =================
      const uint32_t idesc1 = umma_idesc_tf32(128, ROWS);
      const uint32_t idesc2 = umma_idesc_tf32(128, 128);
      if (lane == 0) {
        for (int64_t gg = g; gg < gend; ++gg) {
          const int64_t ci = gg - g0;
          const int stage = (int)(ci % kGtcStages);
          const uint32_t use = (uint32_t)(ci / kGtcStages);
          mbar_wait(smem_u32(&s_full[stage]), use & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kGtmChunk / 8; ++ks) {   // K = 8 tf32 (32 bytes) per instruction; 4 steps per box
            const uint32_t base = smem_u32(slabs + stage * SLAB + (ks >> 2) * BOX);
            const uint32_t accumulate = (gg > g || ks > 0) ? 1u : 0u;
            if (a.debug & 1) continue;
            const uint64_t d_lo = umma_desc_k_sw128(base + (ks & 3) * 32);
            umma_tf32(tmem_base, d_lo, d_lo, idesc1, accumulate);
            if (ROWS == 256) {
              const uint64_t d_hi = umma_desc_k_sw128(base + 128 * 128 + (ks & 3) * 32);
              umma_tf32(tmem_base + 256, d_hi, d_hi, idesc2, accumulate);
            }
          }
          umma_commit(smem_u32(&s_empty[stage]));   // slab may be refilled once these MMAs retire
        }
        umma_commit(smem_u32(&s_accum));            // accumulators of this segment complete
      }
      __syncwarp();
    } else if (warp >= 1 && warp <= kRsWarps) {
      // ================= row sums over every stage as it lands =================
      // warp w covers rows 32 (w-1) .. +31; 8 lanes share a row (one 16-byte chunk each), so a warp load reads 4 whole
      // rows = 512 contiguous bytes (conflict-free); lane (row sub-index s = lane / 8) accumulates rows 4 i + s, i < 8.
      const int rbase = (warp - 1) * 32, sub = lane >> 3, ch = lane & 7;
      float rs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rs[i] = 0.f;
      for (int64_t gg = g; gg < gend; ++gg) {
        const int64_t ci = gg - g0;
        const int stage = (int)(ci % kGtcStages);
        const uint32_t use = (uint32_t)(ci / kGtcStages);
        mbar_wait(smem_u32(&s_full[stage]), use & 1);
        // the 16-byte chunks of a row are swizzled within the row: the order does not matter for a sum.  Truncate like
        // the tensor core (kind::tf32 drops the low 13 mantissa bits) so that the identity in the header holds exactly
        // for what the MMAs accumulate.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const unsigned char* bx = slabs + stage * SLAB + h * BOX + (rbase + sub) * 128 + ch * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(bx + i * 4 * 128);
            rs[i] += (__uint_as_float(__float_as_uint(v.x) & 0xffffe000u) + __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)) +
                     (__uint_as_float(__float_as_uint(v.z) & 0xffffe000u) + __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[stage])) : "memory");
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = rs[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (ch == 0) s_rowsum[rbase + 4 * i + sub] = v;
      }
    }
    if (warp < kGtcWarps) {
      __syncwarp();
      // ================= epilogue (all worker warps): TMEM -> registers -> partial[sg][128][PP] =================
      mbar_wait(smem_u32(&s_accum), (uint32_t)sg & 1);
      tc_fence_after();
      float* part = part0 + (size_t)sg * 128 * PP;
      const int q = warp & 3;                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
      constexpr int NCG = PW / 32;            // 32-column groups, dealt round-robin to the warps of a quarter
      // tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns).  The partial [128][PP] is assembled in
      // the (now idle) ring in its global layout -- lane = row, 16-byte stores at a pitch of PP floats = 4 banks mod 32:
      // a quarter-warp covers all 32 banks, conflict-free -- and leaves as ONE contiguous bulk copy.  (Storing straight
      // from the registers, eight 16-byte pieces per lane 1552 bytes apart, cost 8.3 us of the kernel's 41.)
      float* stage_out = reinterpret_cast<float*>(slabs);
      // every worker warp is here: the MMAs have retired (s_accum) AND the row-sum warps have read the last stages and
      // published s_rowsum -- only now may the ring be overwritten
      named_bar_sync(kSegBarrier + 1, kGtcWarps * 32);
      for (int cg = warp >> 2; cg < NCG && !(a.debug & 2); cg += kGtcWarps / 4) {
        const int c = cg * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        float4* dst = reinterpret_cast<float4*>(stage_out + (size_t)(q * 32 + lane) * PP + c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
      if (warp < 4) {
        float4 tail;
        tail.x = s_rowsum[warp * 32 + lane];
        tail.y = (ROWS == 256) ? s_rowsum[128 + warp * 32 + lane] : 0.f;
        tail.z = 0.f;
        tail.w = 0.f;
        *reinterpret_cast<float4*>(stage_out + (size_t)(warp * 32 + lane) * PP + PW) = tail;
      }
      fence_proxy_async();                                   // this thread's staging writes -> visible to the bulk copy
      named_bar_sync(kSegBarrier + 2, kGtcWarps * 32);
      if (tid == 0 && !(a.debug & 2)) {
        bulk_s2g(part, smem_u32(stage_out), (uint32_t)(128 * PP * sizeof(float)));
        bulk_commit();
        bulk_wait_read<0>();                                 // the ring may be refilled / the CTA may exit
      }
      tc_fence_before();
      // every worker warp has read its part of TMEM and is done with its transpose tile before the next segment's
      // boxes land in the ring and its first MMA overwrites the accumulators
      if (sg + 1 < nsegs) named_bar_sync(kSegBarrier, kGtcWarps * 32);
    }
    g = gend;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGtcWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
#endif
}

// G[b] = sum of the (CTA, segment) partials that cover matrix b, in CTA order; G10 mirrored from G01.
// One thread per four consecutive partial elements (pr, pc..pc+3): float4 reads along pc, float4 writes of the
// direct blocks; the mirrored block is the only scattered write (1/4 of a 256 KB matrix).
__global__ void gram_reduce_kernel(const float* partial, int rows, int64_t nchunk, int64_t per, float* G) {
  const int PW = gram_tc_partial_width(rows), PP = gram_tc_partial_pitch(rows);
  const int PW4 = PW / 4;
  const int64_t b = blockIdx.y;
  const int i0 = (int)((b * nchunk) / per), i1 = (int)(((b + 1) * nchunk - 1) / per);
  float* Gb = G + b * (int64_t)rows * rows;
  const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= 128 * PW4) return;
  const int pr = e4 / PW4, pc = (e4 - pr * PW4) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t bstart = b * nchunk;
  for (int cta = i0; cta <= i1; ++cta) {
    // a CTA whose range starts before this matrix began in matrix b-1: matrix b is its second segment
    const int sg = ((int64_t)cta * per < bstart) ? 1 : 0;
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)(cta * 2 + sg) * 128 + pr) * PP + pc));
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  if (pc < rows) {
    *reinterpret_cast<float4*>(Gb + pr * rows + pc) = s;          // D1 = [G00 | G01]
    if (pc >= 128) {                                              // G10 = G01^T
      Gb[(pc + 0) * rows + pr] = s.x;
      Gb[(pc + 1) * rows + pr] = s.y;
      Gb[(pc + 2) * rows + pr] = s.z;
      Gb[(pc + 3) * rows + pr] = s.w;
    }
  } else {
    *reinterpret_cast<float4*>(Gb + (128 + pr) * rows + 128 + (pc - rows)) = s;   // D2 = G11
  }
}

struct GramTcGeom {
  int64_t nchunk, per, grid;
};

static GramTcGeom gram_tc_geom(int64_t B, int64_t cols, int num_sms, int chunk = kGtcChunk) {
  GramTcGeom g;
  g.nchunk = ceil_div(cols, chunk);
  if (chunk == kGtmChunk) {
    // TMA-fed kernel: a whole number of CTAs per matrix (3 for 40 matrices on 148 SMs), so that no CTA straddles two
    // matrices (one epilogue per CTA) and every matrix has the same, small number of partials to sum
    const int64_t k = std::max<int64_t>(1, std::min<int64_t>(num_sms / std::max<int64_t>(B, 1), g.nchunk));
    g.per = ceil_div(g.nchunk, k);
    g.grid = B * ceil_div(g.nchunk, g.per);
    return g;
  }
  const int64_t total = B * g.nchunk;
  g.per = std::min<int64_t>(std::max<int64_t>(ceil_div(total, num_sms), 1), g.nchunk);   // <= nchunk: at most 2 segments per CTA
  g.grid = ceil_div(total, g.per);
  return g;
}

size_t gram_tc_workspace_bytes(int64_t B, int64_t rows) {
  // grid <= max(B, num_sms) + 1 CTAs, two partial slots each
  return (size_t)(std::max<int64_t>(B, 160) + 8) * 2 * 128 * gram_tc_partial_pitch((int)rows) * sizeof(float) + 256;
}

bool gram_tc_supported(int64_t rows) { return rows == 128 || rows == 256; }

void gram_tc_geometry(int64_t B, int64_t cols, int num_sms, int tma, int64_t* nchunk, int64_t* per) {
  const GramTcGeom g = gram_tc_geom(B, cols, std::min(num_sms, 160), tma ? kGtmChunk : kGtcChunk);
  *nchunk = g.nchunk;
  *per = g.per;
}

// Raw-operand Gram partials through the TMA-fed kernel (see gram_tma_kernel); returns 1 when the layout does not allow a
// tensor map (rows not 16-byte aligned), 0 on success, otherwise a CUDA error.  *raw_out tells the consumer
// (launch_gram_eig) that the partials are of the un-normalised operands and carry row sums.
int launch_gram_tma(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, float* partial_ws, int num_sms,
                    cudaStream_t stream, float l2_pin) {
  if (B == 0) return 0;
  const GramTcGeom g = gram_tc_geom(B, cols, std::min(num_sms, 160), kGtmChunk);
  TensorMap tmap{};
#if !defined(SPECGPU_EMULATE)
  if (ld < 0) {
    const uint64_t nt = (uint64_t)(-ld);
    if (!make_tensor_map_f32_3d(&tmap, S, kTileCols, nt * rows, (uint64_t)B, kTileCols, nt * rows * kTileCols, 32, (uint32_t)rows, 128))
      return 1;
  } else if (!make_tensor_map_f32_3d(&tmap, S, (uint64_t)cols, (uint64_t)rows, (uint64_t)B, (uint64_t)ld, (uint64_t)rows * ld, 32,
                                     (uint32_t)rows, 128)) {
    return 1;
  }
#endif
  GramTmaArgs a{};
  a.tiled = ld < 0 ? 1 : 0;
  a.B = B;
  a.cols = cols;
  a.nchunk = g.nchunk;
  a.per = g.per;
  a.l2_pin = l2_pin;
  a.partial = partial_ws;
  if (const char* env = std::getenv("SPECGPU_GRAM_DEBUG")) a.debug = std::atoi(env);
  // the ring; the epilogue assembles the [128][PW + 4] partial in it (a little larger than the ring for 256 rows)
  const size_t smem = std::max((size_t)kGtmStages * 2 * rows * 128, (size_t)128 * ((rows == 256 ? 384 : rows) + 4) * sizeof(float)) + 1024;
  if (rows == 256) {
    cudaError_t e = cudaFuncSetAttribute(gram_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH_PDL(gram_tma_kernel<256>, (unsigned)g.grid, kGtcThreads, smem, stream, 1, a, S, ld, tmap);
  } else {
    cudaError_t e = cudaFuncSetAttribute(gram_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH_PDL(gram_tma_kernel<128>, (unsigned)g.grid, kGtcThreads, smem, stream, 1, a, S, ld, tmap);
  }
  return (int)cudaGetLastError();
}

int launch_gram_tc(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const MinMaxWord* minmax,
                   float* partial_ws, float* G, int num_sms, cudaStream_t stream, float l2_pin) {
  if (B == 0) return 0;
  if (ld < 0) return (int)cudaErrorInvalidValue;     // the cp.async-fed kernel reads row-major images only
  const GramTcGeom g = gram_tc_geom(B, cols, std::min(num_sms, 160));
  GramTcArgs a{};
  a.S = S;
  a.B = B;
  a.cols = cols;
  a.ld = ld;
  a.nchunk = g.nchunk;
  a.per = g.per;
  a.minmax = minmax;
  a.vec16 = (((reinterpret_cast<uintptr_t>(S) & 15) == 0) && (ld % 4 == 0) && (ld >= ((cols + 3) & ~(int64_t)3))) ? 1 : 0;
  a.partial = partial_ws;
  a.l2_pin = l2_pin;
  const size_t smem = std::max((size_t)kGtcStages * rows * 128, (size_t)kGtcWarps * 32 * 33 * sizeof(float)) + 1024;
  if (rows == 256) {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(gram_tc_kernel<256>, (unsigned)g.grid, kGtcThreads, smem, stream, a);
  } else {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(gram_tc_kernel<128>, (unsigned)g.grid, kGtcThreads, smem, stream, a);
  }
  int e = (int)cudaGetLastError();
  if (e) return e;
  if (G == nullptr) return 0;      // the partials are consumed by launch_gram_eig
  SPECGPU_LAUNCH(gram_reduce_kernel, dim3((unsigned)ceil_div(32 * gram_tc_partial_width((int)rows), 256), (unsigned)B), 256,
                 0, stream, (const float*)partial_ws, (int)rows, g.nchunk, g.per, G);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
