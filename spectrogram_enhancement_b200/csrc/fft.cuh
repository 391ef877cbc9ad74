// Register/shared-memory FFT building blocks.
//
// A length-M complex transform (M = nperseg/2: the real segment is packed as M complex numbers) is
// done by a group of G = M/R0 threads as a Stockham autosort FFT: every pass each thread holds R0
// complex values in registers, performs R0/R radix-R butterflies on them, and exchanges through a
// padded shared-memory line.  Radices per size (first pass = R0):
//   M:   4    8    16    32     64    128     256      512      1024      2048       4096
//        4    8    16   8,4    8,8   16,8   16,16   16,8,4    16,8,8   16,16,8   16,16,16
#pragma once
#include "common.cuh"

namespace specgpu {

__host__ __device__ constexpr int fft_radix_at(int log2m, int idx) {
  constexpr int T[13][3] = {{1, 1, 1},  {1, 1, 1},   {4, 1, 1},   {8, 1, 1},   {16, 1, 1},
                            {8, 4, 1},  {8, 8, 1},   {16, 8, 1},  {16, 16, 1}, {16, 8, 4},
                            {16, 8, 8}, {16, 16, 8}, {16, 16, 16}};
  return T[log2m][idx];
}
__host__ __device__ constexpr int fft_num_passes(int log2m) {
  return (fft_radix_at(log2m, 1) == 1) ? 1 : ((fft_radix_at(log2m, 2) == 1) ? 2 : 3);
}

// Padded index into a shared-memory FFT line: one float2 of padding per 16 keeps the strided
// first-pass stores and last-pass loads off a single bank.
__host__ __device__ __forceinline__ constexpr int fft_pad(int i) { return i + (i >> 4); }

// cos/sin(2*pi*k/16), k = 0..7 -- every twiddle of an in-register radix <= 16 butterfly.
__device__ __forceinline__ float2 w16(int k) {
  constexpr float C[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                          0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
  constexpr float S[8] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                          1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
  return make_float2(C[k], -S[k]);  // exp(-2*pi*i*k/16)
}

// Forward DFT of R values held in registers (natural order in, natural order out), in packed (re, im) arithmetic:
// 8 / 28 / 84 two-lane instructions for R = 4 / 8 / 16 (the scalar form of the radix-16 butterfly is ~170).
template <int R>
__device__ __forceinline__ void dft_reg(float2 (&v)[R]) {
  if constexpr (R == 2) {
    float2 a = v[0], b = v[1];
    v[0] = add2(a, b);
    v[1] = sub2(a, b);
  } else if constexpr (R == 4) {
    float2 t0 = add2(v[0], v[2]), t1 = sub2(v[0], v[2]);
    float2 t2 = add2(v[1], v[3]), t3 = sub2(v[1], v[3]);
    v[0] = add2(t0, t2);
    v[2] = sub2(t0, t2);
    v[1] = sub_i(t1, t3);
    v[3] = add_i(t1, t3);
  } else if constexpr (R > 4) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      e[k] = v[2 * k];
      o[k] = v[2 * k + 1];
    }
    dft_reg<R / 2>(e);
    dft_reg<R / 2>(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      if (k == 0) {
        v[k] = add2(e[k], o[k]);
        v[k + R / 2] = sub2(e[k], o[k]);
      } else if (4 * k == R) {
        v[k] = sub_i(e[k], o[k]);
        v[k + R / 2] = add_i(e[k], o[k]);
      } else {
        const float2 w = w16(k * (16 / R));
        v[k] = cfma(w, o[k], e[k]);
        v[k + R / 2] = cfms(w, o[k], e[k]);
      }
    }
  }
}

// One Stockham pass (radix R, p = product of the radices of earlier passes) on the R0 register values of
// thread `tg` of a G-thread group.  Registers v[q*R + r] hold element r of butterfly i_q = tg + q*G.
//   load : v <- line[i_q + r*M/R] * W_(pR)^(k*r),  k = i_q mod p
//   store: line[(i_q-k)*R + k + r*p] <- DFT_R(v)
// The twiddles of a pass come from their own table tw[r*p + k] = exp(-2 pi i k r / (p R)): for one r the threads of a
// group read consecutive entries (no bank conflicts; indexing one table exp(-2 pi i j / M) with j = k r (M/(pR)) made
// the even-r loads 2- to 8-way conflicted and cost 20 % of the kernel's shared-memory wavefronts).
template <int M, int R0, int R, int P>
struct FftPass {
  static constexpr int G = M / R0;
  static constexpr int NB = R0 / R;
  static __device__ __forceinline__ void load(float2 (&v)[R0], const float2* line, const float2* tw, int tg) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int i = tg + q * G;
      const int k = i & (P - 1);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 a = line[fft_pad(i + r * (M / R))];
        if (r > 0 && P > 1) a = cmul(a, tw[r * P + k]);
        v[q * R + r] = a;
      }
    }
  }
  static __device__ __forceinline__ void butterflies(float2 (&v)[R0]) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      float2 u[R];
#pragma unroll
      for (int r = 0; r < R; ++r) u[r] = v[q * R + r];
      dft_reg<R>(u);
#pragma unroll
      for (int r = 0; r < R; ++r) v[q * R + r] = u[r];
    }
  }
  static __device__ __forceinline__ void store(const float2 (&v)[R0], float2* line, int tg) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int i = tg + q * G;
      const int k = i & (P - 1);
      const int j = (i - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) line[fft_pad(j + r * P)] = v[q * R + r];
    }
  }
};

// Barrier among the threads that share one FFT line.  A group of G <= 32 threads lives inside one warp (G divides
// 32 and groups are warp-aligned), so a warp barrier is enough and the other warps of the CTA keep running;
// larger groups span warps and need the CTA barrier.
template <int G>
__device__ __forceinline__ void fft_group_sync() {
  if constexpr (G <= 32) __syncwarp();
  else __syncthreads();
}

// Number of float2 entries of the per-pass twiddle tables of a length-M transform: pass 1 (R1 x R0) then pass 2 (R2 x R0 R1).
__host__ __device__ constexpr int fft_twiddle_count(int log2m) {
  const int r0 = fft_radix_at(log2m, 0), r1 = fft_radix_at(log2m, 1), r2 = fft_radix_at(log2m, 2);
  return (r1 > 1 ? r0 * r1 : 0) + (r2 > 1 ? r0 * r1 * r2 : 0);
}
// Radix of the last pass, and the register that holds output element m of thread tg after it:
// thread tg ends up with Z[tg + G m], m < R0, in v[(m % NB) * RL + m / NB] with NB = R0 / RL.
__host__ __device__ constexpr int fft_last_radix(int log2m) {
  return fft_radix_at(log2m, 2) > 1 ? fft_radix_at(log2m, 2) : (fft_radix_at(log2m, 1) > 1 ? fft_radix_at(log2m, 1) : fft_radix_at(log2m, 0));
}
__host__ __device__ constexpr int fft_out_reg(int log2m, int m) {
  const int rl = fft_last_radix(log2m), nb = fft_radix_at(log2m, 0) / rl;
  return (m % nb) * rl + m / nb;
}

// Full transform.  On entry v holds the first-pass inputs of thread tg (element r = z[tg + r*G]).
// KEEP_REGS = false: on exit line[fft_pad(k)] = Z[k], k < M, and every thread of the group has passed a group barrier.
// KEEP_REGS = true : the outputs of the last pass stay in registers -- thread tg holds Z[tg + G m] in
//                    v[fft_out_reg(LOG2M, m)] -- and the line is only used between passes.
// All threads of the CTA must call this together (groups of more than 32 threads use __syncthreads()).
// HALF_LINE (two-pass sizes whose groups live inside one warp, KEEP_REGS only): the exchange between the two passes goes
// through a line of M + M/16 FLOATS -- real parts first, then imaginary parts through the same words.  Same number of
// shared-memory wavefronts as the float2 line (a 64-bit warp access takes two), twice the LSU instructions, half the
// shared memory: that half pays for the second output tile of the STFT kernel.  With a line stride of M + M/16 floats
// (= 16 mod 32 for M = 256) the groups of a warp fall on disjoint banks.
template <int LOG2M, bool KEEP_REGS = false, bool HALF_LINE = false, bool ABL_TW = false /* timing ablation, see stft.cu */>
__device__ __forceinline__ void fft_group(float2 (&v)[fft_radix_at(LOG2M, 0)], float2* line, const float2* tw, int tg) {
  constexpr int M = 1 << LOG2M;
  constexpr int R0 = fft_radix_at(LOG2M, 0);
  constexpr int R1 = fft_radix_at(LOG2M, 1);
  constexpr int R2 = fft_radix_at(LOG2M, 2);
  constexpr int G = M / R0;
  // pass 0: a single radix-R0 butterfly straight from registers
  FftPass<M, R0, R0, 1>::butterflies(v);
  if constexpr (R1 == 1 && KEEP_REGS) return;
  if constexpr (HALF_LINE) {
    static_assert(KEEP_REGS && R1 > 1 && R2 == 1 && G <= 32, "half line: two passes, group inside a warp, outputs in registers");
    float* lf = reinterpret_cast<float*>(line);
    constexpr int NB = R0 / R1;
    const int j0 = tg * R0;                       // pass-0 store index of element r: (i - k) R + k + r p with p = 1, k = 0
    float2 u[R0];
#pragma unroll
    for (int r = 0; r < R0; ++r) lf[fft_pad(j0 + r)] = v[r].x;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int r = 0; r < R1; ++r) u[q * R1 + r].x = lf[fft_pad(tg + q * G + r * (M / R1))];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R0; ++r) lf[fft_pad(j0 + r)] = v[r].y;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int k = (tg + q * G) & (R0 - 1);
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        u[q * R1 + r].y = lf[fft_pad(tg + q * G + r * (M / R1))];
        const float2 w = ABL_TW ? make_float2(1.f - (float)(tg + r) * 1e-3f, (float)(tg + r) * 1e-3f) : tw[(r > 0 ? r : 1) * R0 + k];
        v[q * R1 + r] = (r > 0) ? cmul(u[q * R1 + r], w) : u[q * R1 + r];
      }
    }
    FftPass<M, R0, R1, R0>::butterflies(v);
    return;
  }
  FftPass<M, R0, R0, 1>::store(v, line, tg);
  fft_group_sync<G>();
  if constexpr (R1 > 1) {
    FftPass<M, R0, R1, R0>::load(v, line, tw, tg);
    FftPass<M, R0, R1, R0>::butterflies(v);
    if constexpr (R2 == 1 && KEEP_REGS) return;
    fft_group_sync<G>();
    FftPass<M, R0, R1, R0>::store(v, line, tg);
    fft_group_sync<G>();
  }
  if constexpr (R2 > 1) {
    FftPass<M, R0, R2, R0 * R1>::load(v, line, tw + R0 * R1, tg);
    FftPass<M, R0, R2, R0 * R1>::butterflies(v);
    if constexpr (KEEP_REGS) return;
    fft_group_sync<G>();
    FftPass<M, R0, R2, R0 * R1>::store(v, line, tg);
    fft_group_sync<G>();
  }
}

}  // namespace specgpu
