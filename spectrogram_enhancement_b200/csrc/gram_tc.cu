// K3a: Gram matrix G = S S^T on the 5th-generation tensor cores (tcgen05, TF32 operands, FP32
// accumulators in tensor memory) for rows in {128, 256} -- the one dense contraction of the path.
//
// Work unit = (matrix b, split s): a contiguous range of 32-column chunks of S[b].  Per chunk the
// producer warps read the [rows x 32] fp32 slab (row-contiguous 128 B runs, a whole slab in flight in
// registers before the first store), optionally apply the min-max normalisation of the log image
// (so that the pipeline needs no separate normalise pass), round it to TF32
// (cvt.rna) and store it to shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (row r at r*128 B, 16-byte chunk c at (c ^ (r & 7))); a 4-deep mbarrier ring hands slabs to one
// elected thread that issues tcgen05.mma.kind::tf32 with BOTH operands described on the same slab:
//     D1[128 x rows] += slab[0:128]   . slab[0:rows]^T     (G00 | G01)
//     D2[128 x 128 ] += slab[128:256] . slab[128:256]^T    (G11; rows == 256 only; G10 = G01^T)
// so the symmetric product costs 3/4 of the MMA work.  tcgen05.commit releases the slab; after the
// last chunk the accumulators are read back with tcgen05.ld and written as a per-unit partial.
// gram_reduce_kernel sums the split partials in a fixed order (deterministic) and mirrors G10.
//
// The emulation build (tests only, no tensor cores on a CPU) replaces the kernel body by a scalar
// loop with the same TF32 operand rounding and the same partial layout.
#include "kernels.h"

namespace specgpu {

constexpr int kGtcStages = 4;
constexpr int kGtcChunk = 32;            // K elements per slab (128 bytes of tf32 per row)
constexpr int kGtcProducerWarps = 8;
constexpr int kGtcThreads = (kGtcProducerWarps + 1) * 32;

struct GramTcArgs {
  const float* S;
  int64_t cols, ld;
  int nsplit;
  const unsigned* minmax;   // optional [B][2] ordered-uint (min, max): operands are (x - min) / (max - min)
  float* partial;   // [B][nsplit][128][PW] with PW = rows + (rows == 256 ? 128 : 0)
};

__host__ __device__ inline int gram_tc_partial_width(int rows) { return rows == 256 ? 384 : rows; }

__host__ __device__ inline void gram_tc_range(int64_t cols, int nsplit, int s, int64_t* c0, int64_t* c1) {
  const int64_t nchunk = (cols + kGtcChunk - 1) / kGtcChunk;
  const int64_t per = (nchunk + nsplit - 1) / nsplit;
  *c0 = (int64_t)s * per;
  *c1 = (*c0 + per < nchunk) ? *c0 + per : nchunk;
  if (*c0 > nchunk) *c0 = nchunk;
}

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
  const int64_t b = blockIdx.x / a.nsplit;
  const int s = blockIdx.x % a.nsplit;
  int64_t c0, c1;
  gram_tc_range(a.cols, a.nsplit, s, &c0, &c1);
  const float* Sb = a.S + b * ROWS * a.ld;
  float* part = a.partial + ((size_t)b * a.nsplit + s) * 128 * PW;
  const int tid = threadIdx.x;

#if defined(SPECGPU_EMULATE)
  // scalar stand-in with identical operand rounding and output layout
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
    for (int64_t k = c0 * kGtcChunk; k < c1 * kGtcChunk && k < a.cols; ++k) {
      float xa = Sb[(int64_t)ra * a.ld + k], xb = Sb[(int64_t)rb * a.ld + k];
      if (a.minmax != nullptr) {
        xa = __fdiv_rn(xa - e_mn, e_den);
        xb = __fdiv_rn(xb - e_mn, e_den);
      }
      acc += round_tf32(xa) * round_tf32(xb);
    }
    part[i] = acc;
  }
#else
  SPECGPU_DYN_SMEM(smem);   // 1024-byte aligned: required by SWIZZLE_128B
  constexpr int SLAB = ROWS * 128;  // bytes per stage
  __shared__ __align__(8) uint64_t s_full[kGtcStages];
  __shared__ __align__(8) uint64_t s_empty[kGtcStages];
  __shared__ __align__(8) uint64_t s_accum;
  __shared__ uint32_t s_tmem;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (ROWS == 256) ? 512 : 128;

  if (tid == 0) {
    for (int i = 0; i < kGtcStages; ++i) {
      mbar_init(smem_u32(&s_full[i]), kGtcProducerWarps * 32);
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
  const int64_t nch = c1 - c0;
  unsigned char* slabs = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned

  if (warp < kGtcProducerWarps) {
    // ================= producers: global fp32 -> (normalise) -> TF32 -> swizzled shared slab =================
    float mn = 0.f, den = 1.f;
    const bool do_norm = a.minmax != nullptr;
    if (do_norm) {
      mn = ordered_to_float(a.minmax[2 * b]);
      den = ordered_to_float(a.minmax[2 * b + 1]) - mn;
    }
    constexpr int RPW = ROWS / kGtcProducerWarps;   // rows per warp per chunk (lane = column)
    constexpr int BATCH = 16;                        // loads in flight per thread
    for (int64_t ci = 0; ci < nch; ++ci) {
      const int stage = (int)(ci % kGtcStages);
      const uint32_t use = (uint32_t)(ci / kGtcStages);
      unsigned char* slab = slabs + stage * SLAB;
      const int64_t k = (c0 + ci) * kGtcChunk + lane;
      const bool kok = k < a.cols;
      const float* src = Sb + k;
#pragma unroll
      for (int r0 = 0; r0 < RPW; r0 += BATCH) {
        float v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int r = warp + (r0 + i) * kGtcProducerWarps;
          v[i] = kok ? __ldg(src + (int64_t)r * a.ld) : 0.f;
        }
        if (r0 == 0 && use > 0) mbar_wait(smem_u32(&s_empty[stage]), (use - 1) & 1);   // loads already in flight
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int r = warp + (r0 + i) * kGtcProducerWarps;
          float x = v[i];
          if (do_norm) x = kok ? __fdiv_rn(x - mn, den) : 0.f;
          *reinterpret_cast<float*>(slab + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)) = round_tf32(x);
        }
      }
      fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(smem_u32(&s_full[stage]));
    }
    // ================= epilogue: TMEM -> registers -> partial[128][PW] =================
    if (nch > 0) {
      mbar_wait(smem_u32(&s_accum), 0);
      tc_fence_after();
    }
    const int q = warp & 3;                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int row = q * 32 + lane;          // TMEM lane == accumulator row
    constexpr int NCG = PW / 32;            // 32-column groups, dealt round-robin to the two warps of a quarter
    for (int cg = warp >> 2; cg < NCG; cg += kGtcProducerWarps / 4) {
      const int c = cg * 32;
      uint32_t v[32];
      if (nch > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      float4* dst = reinterpret_cast<float4*>(part + (size_t)row * PW + c);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                             __uint_as_float(v[4 * i + 3]));
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t idesc1 = umma_idesc_tf32(128, ROWS);
    const uint32_t idesc2 = umma_idesc_tf32(128, 128);
    for (int64_t ci = 0; ci < nch; ++ci) {
      const int stage = (int)(ci % kGtcStages);
      const uint32_t use = (uint32_t)(ci / kGtcStages);
      mbar_wait(smem_u32(&s_full[stage]), use & 1);
      tc_fence_after();
      const uint32_t base = smem_u32(slabs + stage * SLAB);
#pragma unroll
      for (int ks = 0; ks < kGtcChunk / 8; ++ks) {   // K = 8 tf32 (32 bytes) per instruction
        const uint64_t d_lo = umma_desc_k_sw128(base + ks * 32);
        umma_tf32(tmem_base, d_lo, d_lo, idesc1, (ci > 0 || ks > 0) ? 1u : 0u);
        if (ROWS == 256) {
          const uint64_t d_hi = umma_desc_k_sw128(base + 128 * 128 + ks * 32);
          umma_tf32(tmem_base + 256, d_hi, d_hi, idesc2, (ci > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&s_empty[stage]));   // slab may be refilled once these MMAs retire
    }
    if (nch > 0) umma_commit(smem_u32(&s_accum));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGtcProducerWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
#endif
}

// G[b] = sum over splits of the unit partials; G10 mirrored from G01.
__global__ void gram_reduce_kernel(const float* partial, int rows, int nsplit, float* G) {
  const int PW = gram_tc_partial_width(rows);
  const int64_t b = blockIdx.y;
  const float* pb = partial + (size_t)b * nsplit * 128 * PW;
  float* Gb = G + b * (int64_t)rows * rows;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * rows; i += gridDim.x * blockDim.x) {
    const int r = i / rows, c = i % rows;
    int pr, pc;
    if (r < 128) {
      pr = r;
      pc = c;                       // D1 = [G00 | G01]
    } else if (c >= 128) {
      pr = r - 128;
      pc = rows + (c - 128);        // D2 = G11
    } else {
      pr = c;
      pc = r;                       // G10 = G01^T
    }
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += pb[((size_t)k * 128 + pr) * PW + pc];
    Gb[i] = s;
  }
}

static int gram_tc_pick_split(int64_t B, int64_t cols, int num_sms) {
  const int64_t nchunk = ceil_div(cols, kGtcChunk);
  int best = 1;
  double best_cost = 1e300;
  for (int ns = 1; ns <= 8 && ns <= nchunk; ++ns) {
    const int64_t per = ceil_div(nchunk, ns);
    const int64_t waves = ceil_div(B * ns, num_sms);
    const double cost = (double)waves * ((double)per + 6.0);   // ~6 chunk-times of prologue/epilogue per unit
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = ns;
    }
  }
  return best;
}

size_t gram_tc_workspace_bytes(int64_t B, int64_t rows) {
  return (size_t)B * 8 * 128 * gram_tc_partial_width((int)rows) * sizeof(float) + 256;
}

bool gram_tc_supported(int64_t rows) { return rows == 128 || rows == 256; }

int launch_gram_tc(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, const unsigned* minmax,
                   float* partial_ws, float* G, int num_sms, cudaStream_t stream) {
  if (B == 0) return 0;
  GramTcArgs a{};
  a.S = S;
  a.cols = cols;
  a.ld = ld;
  a.nsplit = gram_tc_pick_split(B, cols, num_sms);
  a.minmax = minmax;
  a.partial = partial_ws;
  const size_t smem = (size_t)kGtcStages * rows * 128 + 1024;
  if (rows == 256) {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(gram_tc_kernel<256>, (unsigned)(B * a.nsplit), kGtcThreads, smem, stream, a);
  } else {
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH(gram_tc_kernel<128>, (unsigned)(B * a.nsplit), kGtcThreads, smem, stream, a);
  }
  int e = (int)cudaGetLastError();
  if (e) return e;
  SPECGPU_LAUNCH(gram_reduce_kernel, dim3((unsigned)ceil_div(rows * rows, 256 * 4), (unsigned)B), 256, 0, stream,
                 (const float*)partial_ws, (int)rows, a.nsplit, G);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
