// K3a: Gram matrix G = S S^T on the 5th-generation tensor cores (tcgen05, TF32 operands, FP32
// accumulators in tensor memory) for rows in {128, 256} -- the one dense contraction of the path.
//
// The batch is cut into 32-column chunks, numbered g = b * nchunk + c; CTA i of a persistent grid owns
// the contiguous range [i*per, (i+1)*per), which touches at most two matrices ("segments").  Per chunk the
// producer warps read the [rows x 32] fp32 slab (row-contiguous 128 B runs, three slabs in flight in
// registers), optionally apply the min-max normalisation of the log image
// (so that the pipeline needs no separate normalise pass), round it to TF32
// (cvt.rna) and store it to shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (row r at r*128 B, 16-byte chunk c at (c ^ (r & 7))); a 4-deep ring (named barriers producer -> issuer,
// mbarriers tensor core -> producer) hands slabs to one elected thread that issues tcgen05.mma.kind::tf32 with BOTH operands described on the same slab:
//     D1[128 x rows] += slab[0:128]   . slab[0:rows]^T     (G00 | G01)
//     D2[128 x 128 ] += slab[128:256] . slab[128:256]^T    (G11; rows == 256 only; G10 = G01^T)
// so the symmetric product costs 3/4 of the MMA work.  tcgen05.commit releases the slab; after the
// last chunk of a segment the accumulators are read back with tcgen05.ld and written as that
// (CTA, segment) partial.  gram_reduce_kernel sums the partials of a matrix in a fixed order
// (deterministic) and mirrors G10.
//
// The emulation build (tests only, no tensor cores on a CPU) replaces the kernel body by a scalar
// loop with the same TF32 operand rounding and the same partial layout.
#include "kernels.h"

namespace specgpu {

constexpr int kGtcStages = 4;
constexpr int kGtcChunk = 32;            // K elements per slab (128 bytes of tf32 per row)
constexpr int kGtcProducerWarps = 16;
constexpr int kGtcGroups = 2;             // producer groups; group k produces the chunks k, k + kGtcGroups, ... of a CTA
constexpr int kGtcThreads = (kGtcProducerWarps + 1) * 32;

struct GramTcArgs {
  const float* S;
  int64_t B, cols, ld;
  int64_t nchunk;           // chunks per matrix
  int64_t per;              // chunks per CTA
  const unsigned* minmax;   // optional [B][2] ordered-uint (min, max): operands are (x - min) / (max - min)
  float* partial;           // [grid][2][128][PW] with PW = rows + (rows == 256 ? 128 : 0)
};

__host__ __device__ inline int gram_tc_partial_width(int rows) { return rows == 256 ? 384 : rows; }

__device__ __forceinline__ float round_tf32(float x) {
#if defined(SPECGPU_EMULATE)
  // cvt.rna.tf32.f32: round to nearest, ties away from zero, keep 10 mantissa bits
  unsigned u = __float_as_uint(x);
  u = (u + 0x1000u) & 0xffffe000u;
  return __uint_as_float(u);
#else
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#endif
}

#if !defined(SPECGPU_EMULATE)
// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// A load the compiler may not sink towards its first use: asm volatile keeps its program order relative to the other
// volatile asm statements (mbarrier waits/arrives), so three slabs really are in flight per thread.
__device__ __forceinline__ float ldg_pinned(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
  constexpr int PW = (ROWS == 256) ? 384 : ROWS;
  const int tid = threadIdx.x;
  const int64_t total = a.B * a.nchunk;
  const int64_t g0 = (int64_t)blockIdx.x * a.per;
  const int64_t g1 = (g0 + a.per < total) ? g0 + a.per : total;
  if (g0 >= g1) return;                          // uniform over the CTA
  const int64_t b_first = g0 / a.nchunk;
  const int nsegs = (int)((g1 - 1) / a.nchunk - b_first) + 1;     // 1 or 2 (per <= nchunk)
  float* part0 = a.partial + (size_t)blockIdx.x * 2 * 128 * PW;

#if defined(SPECGPU_EMULATE)
  // scalar stand-in with identical operand rounding, work split and partial layout
  for (int sg = 0; sg < nsegs; ++sg) {
    const int64_t b = b_first + sg;
    const int64_t lo = (g0 > b * a.nchunk ? g0 : b * a.nchunk) - b * a.nchunk;
    const int64_t hi = (g1 < (b + 1) * a.nchunk ? g1 : (b + 1) * a.nchunk) - b * a.nchunk;
    const float* Sb = a.S + b * ROWS * a.ld;
    float e_mn = 0.f, e_den = 1.f;
    if (a.minmax != nullptr) {
      e_mn = ordered_to_float(a.minmax[2 * b]);
      e_den = ordered_to_float(a.minmax[2 * b + 1]) - e_mn;
    }
    for (int i = tid; i < 128 * PW; i += kGtcThreads) {
      const int r = i / PW, c = i % PW;
      const int ra = (c < ROWS) ? r : 128 + r;
      const int rb = (c < ROWS) ? c : c - ROWS + 128;
      float acc = 0.f;
      for (int64_t k = lo * kGtcChunk; k < hi * kGtcChunk && k < a.cols; ++k) {
        float xa = Sb[(int64_t)ra * a.ld + k], xb = Sb[(int64_t)rb * a.ld + k];
        if (a.minmax != nullptr) {
          xa = div_by(xa - e_mn, e_den, 1.0f / e_den);
          xb = div_by(xb - e_mn, e_den, 1.0f / e_den);
        }
        acc += round_tf32(xa) * round_tf32(xb);
      }
      part0[(size_t)sg * 128 * PW + i] = acc;
    }
  }
#else
  SPECGPU_DYN_SMEM(smem);   // SWIZZLE_128B atoms need 1024-byte alignment (re-aligned below)
  constexpr int SLAB = ROWS * 128;  // bytes per stage
  __shared__ __align__(8) uint64_t s_empty[kGtcStages];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (ROWS == 256) ? 512 : 128;

  if (tid == 0) {
    for (int i = 0; i < kGtcStages; ++i) {
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGtcProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  unsigned char* slabs = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);

  if (warp < kGtcProducerWarps) {
    // ================= producers: global fp32 -> (normalise) -> TF32 -> swizzled shared slab =================
    // The 16 producer warps form kGtcStages groups of 4; group k owns ring stage k and produces the chunks
    // k, k+4, k+8, ... of this CTA.  The proxy fence that must follow the shared-memory stores compiles to
    // MEMBAR.ALL.CTA, which also drains the thread's global loads in flight, so a thread can only keep the loads
    // of ONE chunk in flight; with four groups working on four different chunks the SM still has 3-4 slabs
    // (96-128 KB) in flight, which is what bounds this kernel (one CTA per SM because of TMEM).
    constexpr int WPG = kGtcProducerWarps / kGtcGroups;   // warps per group
    constexpr int RPW = ROWS / WPG;                        // rows per warp per chunk (lane = column)
    const int grp = warp / WPG, wg = warp % WPG;
    const bool do_norm = a.minmax != nullptr;
    const int64_t stride = (int64_t)WPG * a.ld;            // between this thread's consecutive rows (r = wg + WPG*i)
    // row r of a slab lives at r*128 + ((chunk16 ^ (r & 7)) << 4); r & 7 takes the two values wg and wg + 4
    // (WPG = 4) or the single value wg & 7 (WPG = 8)
    const uint32_t sw_even = (uint32_t)((((lane >> 2) ^ (wg & 7)) << 4) + ((lane & 3) << 2));
    const uint32_t sw_odd = (uint32_t)((((lane >> 2) ^ ((wg + WPG) & 7)) << 4) + ((lane & 3) << 2));
    const int ncols = (int)a.cols;
    int64_t g = g0;                                  // first chunk of the current segment
    for (int sg = 0; sg < nsegs; ++sg) {
      const int64_t b = b_first + sg;
      const int64_t gend = (g1 < (b + 1) * a.nchunk) ? g1 : (b + 1) * a.nchunk;
      float mn = 0.f, den = 1.f;
      if (do_norm) {
        mn = ordered_to_float(a.minmax[2 * b]);
        den = ordered_to_float(a.minmax[2 * b + 1]) - mn;
      }
      const float inv = 1.0f / den;
      // this group's chunks of the segment: CTA-relative index ci = grp (mod kGtcGroups)
      int64_t gg = g + ((grp - (int)((g - g0) % kGtcGroups)) + kGtcGroups) % kGtcGroups;
      for (; gg < gend; gg += kGtcGroups) {
        const int c = (int)(gg - b * a.nchunk);           // chunk inside the matrix
        const int stage = (int)((gg - g0) % kGtcStages);
        const uint32_t use = (uint32_t)((gg - g0) / kGtcStages);
        unsigned char* const my_slab = slabs + stage * SLAB + wg * 128;
        const int kcol = c * kGtcChunk + lane;
        const bool kok = kcol < ncols;
        const float* q = a.S + (b * ROWS + wg) * a.ld + kcol;
        float v[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i, q += stride) v[i] = kok ? __ldg(q) : 0.f;
        if (use > 0) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1);     // loads already in flight
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          float x = v[i];
          if (do_norm) x = kok ? div_by(x - mn, den, inv) : 0.f;
          *reinterpret_cast<float*>(my_slab + i * (WPG * 128) + ((i & 1) ? sw_odd : sw_even)) = round_tf32(x);
        }
        fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
        named_bar_arrive(1 + stage, WPG * 32 + 32);
      }
      g = gend;
      // ================= epilogue: TMEM -> registers -> partial[sg][128][PW] =================
      mbar_wait(smem_u32(&s_accum), (uint32_t)sg & 1);
      tc_fence_after();
      float* part = part0 + (size_t)sg * 128 * PW;
      const int q = warp & 3;                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
      constexpr int NCG = PW / 32;            // 32-column groups, dealt round-robin to the warps of a quarter
      // tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns); a per-warp [32][33] shared tile
      // transposes it so that the partial is written as 128-byte rows instead of 32 scattered 16-byte pieces
      float* tr = reinterpret_cast<float*>(slabs + kGtcStages * SLAB) + warp * (32 * 33);
      for (int cg = warp >> 2; cg < NCG; cg += kGtcProducerWarps / 4) {
        const int c = cg * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        float* dst = part + (size_t)(q * 32) * PW + c + lane;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) dst[(size_t)r * PW] = tr[r * 33 + lane];
        __syncwarp();
      }
      tc_fence_before();
      // every producer warp has read its part of TMEM before any slab of the next segment can be completed (the
      // issuer overwrites the accumulators with its first MMA)
      if (sg + 1 < nsegs) named_bar_sync(1 + kGtcStages, kGtcProducerWarps * 32);
    }
  } else {
    // ================= MMA issuer warp: all lanes join the named barrier, lane 0 issues =================
    const uint32_t idesc1 = umma_idesc_tf32(128, ROWS);
    const uint32_t idesc2 = umma_idesc_tf32(128, 128);
    int64_t g = g0;
    for (int sg = 0; sg < nsegs; ++sg) {
      const int64_t b = b_first + sg;
      const int64_t gend = (g1 < (b + 1) * a.nchunk) ? g1 : (b + 1) * a.nchunk;
      const int64_t gstart = g;
      for (; g < gend; ++g) {
        const int64_t ci = g - g0;
        const int stage = (int)(ci % kGtcStages);
        named_bar_sync(1 + stage, (kGtcProducerWarps / kGtcGroups) * 32 + 32);   // this stage's producer group has stored and fenced the slab
        if (lane == 0) {
          tc_fence_after();
          const uint32_t base = smem_u32(slabs + stage * SLAB);
#pragma unroll
          for (int ks = 0; ks < kGtcChunk / 8; ++ks) {   // K = 8 tf32 (32 bytes) per instruction
            const uint32_t accumulate = (g > gstart || ks > 0) ? 1u : 0u;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGtcProducerWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
#endif
}

// G[b] = sum of the (CTA, segment) partials that cover matrix b, in CTA order; G10 mirrored from G01.
// One thread per four consecutive partial elements (pr, pc..pc+3): float4 reads along pc, float4 writes of the
// direct blocks; the mirrored block is the only scattered write (1/4 of a 256 KB matrix).
__global__ void gram_reduce_kernel(const float* partial, int rows, int64_t nchunk, int64_t per, float* G) {
  const int PW = gram_tc_partial_width(rows);
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
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)(cta * 2 + sg) * 128 + pr) * PW + pc));
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

static GramTcGeom gram_tc_geom(int64_t B, int64_t cols, int num_sms) {
  GramTcGeom g;
  g.nchunk = ceil_div(cols, kGtcChunk);
  const int64_t total = B * g.nchunk;
  g.per = std::min<int64_t>(std::max<int64_t>(ceil_div(total, num_sms), 1), g.nchunk);   // <= nchunk: at most 2 segments per CTA
  g.grid = ceil_div(total, g.per);
  return g;
}

size_t gram_tc_workspace_bytes(int64_t B, int64_t rows) {
  // grid <= max(B, num_sms) + 1 CTAs, two partial slots each
  return (size_t)(std::max<int64_t>(B, 160) + 1) * 2 * 128 * gram_tc_partial_width((int)rows) * sizeof(float) + 256;
}

bool gram_tc_supported(int64_t rows) { return rows == 128 || rows == 256; }

int launch_gram_tc(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const unsigned* minmax,
                   float* partial_ws, float* G, int num_sms, cudaStream_t stream) {
  if (B == 0) return 0;
  const GramTcGeom g = gram_tc_geom(B, cols, std::min(num_sms, 160));
  GramTcArgs a{};
  a.S = S;
  a.B = B;
  a.cols = cols;
  a.ld = ld;
  a.nchunk = g.nchunk;
  a.per = g.per;
  a.minmax = minmax;
  a.partial = partial_ws;
  const size_t smem = (size_t)kGtcStages * rows * 128 + 1024 + (size_t)kGtcProducerWarps * 32 * 33 * sizeof(float);
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
  SPECGPU_LAUNCH(gram_reduce_kernel, dim3((unsigned)ceil_div(32 * gram_tc_partial_width((int)rows), 256), (unsigned)B), 256,
                 0, stream, (const float*)partial_ws, (int)rows, g.nchunk, g.per, G);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
