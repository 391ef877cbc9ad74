// K3: SVD range-truncation denoise (spec_denoising/denoising_by_svd.ipynb:155-229, 280-281).
//
// For a rows x cols matrix S (rows <= cols) the left singular vectors are the eigenvectors of the
// rows x rows Gram matrix G = S S^T and s_k = sqrt(lambda_k);   U[:,a:b] diag(s[a:b]) Vh[a:b,:] equals
// U_r U_r^T S, so V is never formed.  Stages:
//   gram      G = S S^T                      (gram_tc.cu: tcgen05 TF32 for rows in {128,256}; here: SIMT fp32)
//   eig       power iteration (leading pair only)  or  cluster-resident one-sided Jacobi (all pairs)
//   plan      singular values, median, Gavish-Donoho count, the reference's start/stop bookkeeping
//   project   out = U_r U_r^T S  (or S - U_c U_c^T S when the complement is smaller), optional clip
#include "kernels.h"
#include "ptx.cuh"

#include <type_traits>

#include "cluster.cuh"

namespace specgpu {

// ======================================================================================================
// Gram matrix, SIMT fp32 (any shape).  64x64 output tile per CTA, split along K; partial sums are
// accumulated with float atomics into a zeroed G, lower triangle mirrored from the upper.
// ======================================================================================================
constexpr int kGramThreads = 256;
// (tile, K block) are template parameters: 64 x 64 outputs with 4 x 4 per thread, K blocks of 32.

// minmax != nullptr: S holds an un-normalised log image, operands are (x - min) / (max - min) exactly as the
// rank-1 projection writes them.  only_flagged != nullptr: only matrices whose plan[b][3] != 0 (leading pair not
// converged) are computed; that launch uses ksplit == 1 and stores its tiles directly (no atomics, no zeroed G).
template <class T, int kGramTile, int kGramKB>
__global__ void __launch_bounds__(kGramThreads) gram_simt_kernel(const float* S, int rows, int64_t cols, int64_t ld,
                                                                 int ksplit, T* G, const MinMaxWord* minmax,
                                                                 const int32_t* only_flagged) {
  // operands are staged in the accumulation type: converting float -> double once per staged element instead of once per
  // use (8 conversions per 16 DFMA in the inner loop made the float64 Gram conversion-bound: 1.04 ms for 40 x [256 x 3905])
  __shared__ T sa[kGramTile][kGramKB + 1];
  __shared__ T sb[kGramTile][kGramKB + 1];
  const int64_t b = blockIdx.z;
  pdl_trigger();
  pdl_wait();
  if (only_flagged != nullptr && only_flagged[b * 4 + 3] == 0) return;   // uniform over the CTA
  float mn = 0.f, den = 1.f;
  if (minmax != nullptr) {
    mn = minmax_get_min(minmax, b);
    den = minmax_get_max(minmax, b) - mn;
  }
  const float inv = 1.0f / den;
  const int nt = (rows + kGramTile - 1) / kGramTile;
  // blockIdx.x enumerates tile pairs (ti <= tj)
  int ti = 0, rem = blockIdx.x;
  while (rem >= nt - ti) {
    rem -= nt - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const int64_t kchunk = ((cols + ksplit - 1) / ksplit + kGramKB - 1) / kGramKB * kGramKB;
  const int64_t k0 = (int64_t)blockIdx.y * kchunk;
  const int64_t k1 = (k0 + kchunk < cols) ? k0 + kchunk : cols;
  const int tid = threadIdx.x;
  constexpr int R = kGramTile / 16;        // outputs per thread and direction
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, R x R outputs each
  T acc[R][R] = {};
  for (int64_t k = k0; k < k1; k += kGramKB) {
    for (int i = tid; i < kGramTile * kGramKB; i += kGramThreads) {
      const int r = i / kGramKB, c = i % kGramKB;
      const int ra = ti * kGramTile + r, rb = tj * kGramTile + r;
      float xa = (ra < rows && k + c < k1) ? S[img_off(b, ra, k + c, rows, ld)] : 0.f;
      float xb = (rb < rows && k + c < k1) ? S[img_off(b, rb, k + c, rows, ld)] : 0.f;
      if (minmax != nullptr) {
        xa = (ra < rows && k + c < k1) ? div_by(xa - mn, den, inv) : 0.f;
        xb = (rb < rows && k + c < k1) ? div_by(xb - mn, den, inv) : 0.f;
      }
      sa[r][c] = (T)xa;
      sb[r][c] = (T)xb;
    }
    __syncthreads();
#pragma unroll (R == 4 ? 8 : 2)
    for (int c = 0; c < kGramKB; ++c) {
      T av[R], bv[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        av[i] = sa[ty + 16 * i][c];
        bv[i] = sb[tx + 16 * i][c];
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] += av[i] * bv[j];
    }
    __syncthreads();
  }
  T* Gb = G + b * (int64_t)rows * rows;
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = ti * kGramTile + ty + 16 * i, c = tj * kGramTile + tx + 16 * j;
      if (r < rows && c < rows) {
        if (ti != tj || c >= r) {
          if (ksplit == 1) {        // this CTA owns the whole sum
            Gb[(int64_t)r * rows + c] = acc[i][j];
            if (r != c) Gb[(int64_t)c * rows + r] = acc[i][j];
          } else {
            atomicAdd(Gb + (int64_t)r * rows + c, acc[i][j]);
            if (r != c) atomicAdd(Gb + (int64_t)c * rows + r, acc[i][j]);
          }
        }
      }
    }
}

// g_f64: accumulate and store G in double (the full-decomposition route; squaring the condition number in fp32 would
// blur cuts between close singular values), else float.
int launch_gram_simt(const float* S, int64_t B, int64_t rows, int64_t cols, int64_t ld, void* G, int g_f64, cudaStream_t stream,
                     const MinMaxWord* minmax, const int32_t* only_flagged) {
  if (B == 0 || rows == 0) return 0;
  // (128 x 128 tiles with 8 x 8 outputs per thread were tried for the float64 route: 194 registers, one CTA per SM, the
  // unpipelined staging exposed -- 1.32 ms against 0.87 ms for 40 x [256 x 3905]; the template stays, the launch does not)
  const bool big = false;
  const int nt = (int)ceil_div(rows, big ? 128 : 64);
  (void)big;
  const int npairs = nt * (nt + 1) / 2;
  int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(cols, 512), 16));
  if (only_flagged != nullptr) {
    ksplit = 1;     // rare repair path: direct stores, nothing to zero
  } else {
    cudaError_t e = cudaMemsetAsync(G, 0, (size_t)B * rows * rows * (g_f64 ? sizeof(double) : sizeof(float)), stream);
    if (e != cudaSuccess) return (int)e;
  }
  const dim3 grid((unsigned)npairs, (unsigned)ksplit, (unsigned)B);
  if (g_f64)
    SPECGPU_LAUNCH_PDL((gram_simt_kernel<double, 64, 32>), grid, kGramThreads, 0, stream, 1, S, (int)rows, cols, ld, ksplit, (double*)G, minmax, only_flagged);
  else
    SPECGPU_LAUNCH_PDL((gram_simt_kernel<float, 64, 32>), grid, kGramThreads, 0, stream, 1, S, (int)rows, cols, ld, ksplit, (float*)G, minmax, only_flagged);
  return (int)cudaGetLastError();
}

// ======================================================================================================
// Leading eigenpair by power iteration: one CTA per matrix, G held in registers (n <= 256) so that an
// iteration costs one pass over shared memory only.  Writes U[:,0], lam[0] and plan = {1, n, -1, status}.
// ======================================================================================================
constexpr int kPowThreads = 1024;
constexpr int kPowMaxIter = 200;

__global__ void __launch_bounds__(kPowThreads) eig_power_kernel(const float* G, int n, int max_iter, float* U, float* lam, int32_t* plan) {
  __shared__ __align__(16) float sx[256];
  __shared__ __align__(16) float sy[256];
  const int64_t b = blockIdx.x;
  const float* Gb = G + b * (int64_t)n * n;
  const int tid = threadIdx.x, lane = tid & 31;
  const int row = tid >> 2, q = tid & 3;  // 4 threads per row; thread q holds columns q*4 + 16*i + {0..3}
  float4 g[16];
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(Gb) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = q * 4 + 16 * i;
      g[i] = (row < n && c < n) ? __ldg(reinterpret_cast<const float4*>(Gb + (int64_t)row * n + c))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = q * 4 + 16 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n) {
        const float* p = Gb + (int64_t)row * n + c;
        if (c + 0 < n) v.x = p[0];
        if (c + 1 < n) v.y = p[1];
        if (c + 2 < n) v.z = p[2];
        if (c + 3 < n) v.w = p[3];
      }
      g[i] = v;
    }
  }
  // Start vector: the column of G with the largest diagonal entry (a first power step from the unit vector of the
  // most energetic row).  A fixed vector such as (1, ..., 1) can be exactly orthogonal to the leading eigenvector
  // (every column of the matrix summing to zero) and the iteration would then sit in the wrong subspace.
  if (tid < 256) sy[tid] = (tid < n) ? Gb[(int64_t)tid * n + tid] : -INFINITY;
  __syncthreads();
  {
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = sy[lane + 32 * i];
      if (d > best) {
        best = d;
        arg = lane + 32 * i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    __syncthreads();
    if (tid < 256) sx[tid] = (tid < n) ? Gb[(int64_t)arg * n + tid] : 0.f;
    __syncthreads();
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += sx[lane + 32 * i] * sx[lane + 32 * i];
    ss = warp_sum(ss);
    const float sinv = (ss > 0.f && ss < INFINITY) ? rsqrtf(ss) : 0.f;
    __syncthreads();
    if (tid < 256) sx[tid] *= sinv;
  }
  __syncthreads();
  float lambda = 0.f, prev_delta = INFINITY;
  int status = 1;
  for (int it = 0; it < max_iter; ++it) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(sx + q * 4 + 16 * i);
      acc += g[i].x * xv.x + g[i].y * xv.y + g[i].z * xv.z + g[i].w * xv.w;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0) sy[row] = acc;
    __syncthreads();
    // every warp redundantly reduces |y|^2, x.y and |y/|y| - x|^2 (uniform control flow, no extra barriers)
    float yy = 0.f, xy = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = sy[lane + 32 * i], x = sx[lane + 32 * i];
      yy += y * y;
      xy += x * y;
    }
    yy = warp_sum(yy);
    xy = warp_sum(xy);
    // a zero or non-finite iterate (G x == 0: x has fallen into the null space; NaN input) is NOT convergence: leave
    // status = 1 so that the full solver redoes this matrix (yy is identical in every warp: uniform exit)
    if (!(yy > 0.f) || !(yy < INFINITY)) break;
    const float inv = rsqrtf(yy);
    float dd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = sy[lane + 32 * i] * inv - sx[lane + 32 * i];
      dd += d * d;
    }
    dd = warp_sum(dd);
    lambda = xy;  // Rayleigh quotient x^T G x with |x| = 1
    __syncthreads();
    if (tid < 256) sx[tid] = sy[tid] * inv;
    __syncthreads();
    if (dd < 1e-13f || (dd < 1e-10f && dd >= prev_delta)) {
      status = 0;
      break;
    }
    prev_delta = dd;
  }
  if (tid < n) U[b * (int64_t)n * n + (int64_t)tid * n] = sx[tid];
  if (tid == 0) {
    lam[b * n] = lambda;
    // the leading-pair route always serves the default range [1, n): plan = {start, stop, num_sing, status}
    plan[b * 4 + 0] = 1;
    plan[b * 4 + 1] = n;
    plan[b * 4 + 2] = -1;
    plan[b * 4 + 3] = status;
  }
}

int launch_eig_power(const float* G, int64_t B, int n, int max_iter, float* U, float* lam, int32_t* plan, cudaStream_t stream) {
  if (B == 0) return 0;
  if (n > 256) return -1;
  if (max_iter <= 0) max_iter = kPowMaxIter;
  SPECGPU_LAUNCH(eig_power_kernel, (unsigned)B, kPowThreads, 0, stream, G, n, max_iter, U, lam, plan);
  return (int)cudaGetLastError();
}

// ======================================================================================================
// Split-K reduce of the tensor-core Gram partials fused with the leading-pair power iteration: one thread-block
// cluster per matrix, CTA r owns rows [32 r, 32 r + 32) of G.  Each CTA sums its rows over the (CTA, segment)
// partials of gram_tc.cu in a fixed order (deterministic) straight into registers -- G never exists in global
// memory -- and the iteration exchanges the 32 new entries of y per CTA through distributed shared memory, one
// cluster barrier per step.  Replaces gram_reduce + eig_power (two launches, a 36 MB round trip and a 1-CTA-per-
// matrix kernel that spent most of its 17 us loading G) on the default route.
//   partial layout (gram_tc.cu): [cta][2][128][PW], PW = 384: D1 = [G00 | G01] in columns 0..255 of rows 0..127,
//   D2 = G11 in columns 256..383; PW = 128 for 128 rows (D1 = G only).  G10 = G01^T is transposed through shared memory.
// ======================================================================================================
// 256 threads per CTA (8 per row of G, 32 columns each): 320 CTAs of 40 matrices then fit the 148 SMs in ONE wave (three
// or more per SM); with 512-thread CTAs at two per SM the last three clusters ran as a second wave and doubled the kernel.
constexpr int kGeThreads = 256, kGeRows = 32, kGeTpr = 8;

struct GramEigArgs {
  const float* partial;
  int64_t nchunk, per;     // geometry of the gram_tc launch that wrote the partials
  const MinMaxWord* raw_minmax;   // != nullptr: partials of the raw image + row sums, see launch_gram_eig
  int64_t cols;
  int per_matrix;          // partials are [b * k + part] (launch_gram_tma), else [cta][2] over global chunk ranges
  int max_iter;
  float* U;                // [B][n][n]: column 0 receives u0
  float* lam;              // [B][n]
  int32_t* plan;           // [B][4] = {1, n, -1, status}
  int32_t* flagged;        // optional [B]: copy of status that survives the repair pass (which rewrites plan[b][3])
  int debug = 0;           // timing ablations of gram_eig1_kernel (SPECGPU_EIG1_DEBUG): 1 = no partial loads, 2 = loads only
};

template <int N>
__global__ void __launch_bounds__(kGeThreads, 3) gram_eig_kernel(GramEigArgs a) {
  constexpr int CL = N / kGeRows;                 // CTAs per cluster
  constexpr int PL = (N == 256) ? 384 : N;        // accumulator columns of a partial row
  constexpr int PW = PL + 4;                      // its pitch (columns PL, PL + 1: row sums of rows r, 128 + r)
  constexpr int NJ = N / (4 * kGeTpr);            // float4 per thread: columns q*4 + 32 j + {0..3}
  constexpr int CS = 4 * kGeTpr;                  // column stride between a thread's float4
  pdl_trigger();
  pdl_wait();
  SPECGPU_DYN_SMEM(smem);
  float* sx = reinterpret_cast<float*>(smem);     // [N] current iterate (every CTA holds all of it)
  float* sy = sx + N;                             // [2][N] G x, double buffered across iterations
  float* sdiag = sy + 2 * N;                      // [CL][2] (largest diagonal, its row) per CTA
  float* sT = sdiag + 2 * CL + 8;                 // [32][132] transposed G10 rows (N == 256, ranks >= CL/2)
  float* sr = sT + kGeRows * 132;                 // [N] row sums of the raw image (raw_minmax route)
  float* scand = sr + N;                          // [CL][N] every CTA's candidate start row
  const int rank = SPECGPU_CLUSTER_RANK();
  const int64_t b = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31;
  const int rl = tid / kGeTpr, q = tid % kGeTpr;
  const int row0 = rank * kGeRows;
  const int kparts = (int)((a.nchunk + a.per - 1) / a.per);
  const int i0 = a.per_matrix ? (int)(b * kparts) : (int)((b * a.nchunk) / a.per);
  const int i1 = a.per_matrix ? i0 + kparts - 1 : (int)(((b + 1) * a.nchunk - 1) / a.per);
  const int64_t bstart = b * a.nchunk;
  // partial of CTA `cta` that belongs to this matrix
  auto part_of = [&](int cta) -> const float* {
    if (a.per_matrix) return a.partial + (size_t)cta * 128 * PW;
    // a CTA whose range starts before this matrix began in matrix b-1: matrix b is its second segment
    const int sg = ((int64_t)cta * a.per < bstart) ? 1 : 0;
    return a.partial + (size_t)(cta * 2 + sg) * 128 * PW;
  };
  const bool lower = (N == 256) && row0 >= 128;   // rows of [G10 | G11]
  float4 g[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lower) {
    for (int i = tid; i < kGeRows * 132; i += kGeThreads) sT[i] = 0.f;
    __syncthreads();
  }
  // The partials are summed in CTA order (deterministic) but loaded kGeBatch at a time: all loads of a batch are in flight
  // before the first add (one memory round trip per batch instead of one per partial).
  constexpr int kGeBatch = 1;
  for (int c0 = i0; c0 <= i1; c0 += kGeBatch) {
    const float* base[kGeBatch];
#pragma unroll
    for (int u = 0; u < kGeBatch; ++u) {
      const int cta = (c0 + u <= i1) ? c0 + u : i1;
      base[u] = part_of(cta);
    }
    if (!lower) {
      float4 v[kGeBatch][NJ];
#pragma unroll
      for (int u = 0; u < kGeBatch; ++u)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          v[u][j] = __ldg(reinterpret_cast<const float4*>(base[u] + (size_t)(row0 + rl) * PW + q * 4 + CS * j));
#pragma unroll
      for (int u = 0; u < kGeBatch; ++u)
        if (c0 + u <= i1) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            g[j].x += v[u][j].x; g[j].y += v[u][j].y; g[j].z += v[u][j].z; g[j].w += v[u][j].w;
          }
        }
    } else {
      float4 v[kGeBatch][NJ / 2], t[kGeBatch][1024 / kGeThreads];
#pragma unroll
      for (int u = 0; u < kGeBatch; ++u) {
#pragma unroll
        for (int j = NJ / 2; j < NJ; ++j)    // G11 from D2
          v[u][j - NJ / 2] = __ldg(reinterpret_cast<const float4*>(base[u] + (size_t)(row0 - 128 + rl) * PW + 256 + q * 4 + CS * (j - NJ / 2)));
        // G10[row][c] = D1[c][row]: 128-byte runs D1[c][row0 .. row0 + 31], accumulated transposed in shared memory (every
        // cell is owned by one thread for all partials: no atomics); 128 * 8 float4 = two per thread
#pragma unroll
        for (int h = 0; h < 1024 / kGeThreads; ++h) {
          const int i = tid + h * kGeThreads;
          t[u][h] = __ldg(reinterpret_cast<const float4*>(base[u] + (size_t)(i >> 3) * PW + row0 + (i & 7) * 4));
        }
      }
#pragma unroll
      for (int u = 0; u < kGeBatch; ++u)
        if (c0 + u <= i1) {
#pragma unroll
          for (int j = NJ / 2; j < NJ; ++j) {
            g[j].x += v[u][j - NJ / 2].x; g[j].y += v[u][j - NJ / 2].y; g[j].z += v[u][j - NJ / 2].z; g[j].w += v[u][j - NJ / 2].w;
          }
#pragma unroll
          for (int h = 0; h < 1024 / kGeThreads; ++h) {
            const int i = tid + h * kGeThreads;
            const int c = i >> 3, e4 = i & 7;
            sT[(e4 * 4 + 0) * 132 + c] += t[u][h].x;
            sT[(e4 * 4 + 1) * 132 + c] += t[u][h].y;
            sT[(e4 * 4 + 2) * 132 + c] += t[u][h].z;
            sT[(e4 * 4 + 3) * 132 + c] += t[u][h].w;
          }
        }
    }
  }
  if (lower) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NJ / 2; ++j) g[j] = *reinterpret_cast<const float4*>(sT + rl * 132 + q * 4 + CS * j);
  }
  float lam_scale = 1.0f;
  if (a.raw_minmax != nullptr) {
    // partials of the RAW image L: (L - m)(L - m)^T = L L^T - m (r 1^T + 1 r^T) + m^2 T 1 1^T  (the common factor
    // 1 / (max - min)^2 of the normalised image does not change eigenvectors and is only applied to lambda)
    if (tid < N) {
      float rs = 0.f;
      for (int c0 = i0; c0 <= i1; c0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cta = (c0 + u <= i1) ? c0 + u : i1;
          v[u] = __ldg(part_of(cta) + (size_t)(tid & 127) * PW + PL + (tid >> 7));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u <= i1) rs += v[u];
      }
      sr[tid] = rs;
    }
    __syncthreads();
    const float m = minmax_get_min(a.raw_minmax, b);
    const float den = minmax_get_max(a.raw_minmax, b) - m;
    lam_scale = 1.0f / (den * den);
    const float ri = sr[row0 + rl];
    const float m2t = m * m * (float)a.cols;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 rc = *reinterpret_cast<const float4*>(sr + q * 4 + CS * j);
      g[j].x = g[j].x - m * (ri + rc.x) + m2t;
      g[j].y = g[j].y - m * (ri + rc.y) + m2t;
      g[j].z = g[j].z - m * (ri + rc.z) + m2t;
      g[j].w = g[j].w - m * (ri + rc.w) + m2t;
    }
  }
  // ---- start vector: the row of G with the largest diagonal entry, cluster-wide (see eig_power_kernel) ----
  {
    // the diagonal entry of row row0 + rl sits in the thread whose columns contain it
    const int dc = row0 + rl;
    float dval = -INFINITY;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c0 = q * 4 + CS * j;
      if (dc >= c0 && dc < c0 + 4) {
        const int e = dc - c0;
        dval = e == 0 ? g[j].x : (e == 1 ? g[j].y : (e == 2 ? g[j].z : g[j].w));
      }
    }
    // reduce over the threads of the row, then over the 32 rows
#pragma unroll
    for (int o = kGeTpr / 2; o > 0; o >>= 1) dval = fmaxf(dval, __shfl_xor_sync(0xffffffffu, dval, o));
    if (q == 0) sy[rl] = dval;
    __syncthreads();
    if (tid < 32) {
      float best = sy[tid];
      int arg = row0 + tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) {
          best = ob;
          arg = oa;
        }
      }
      if (tid < CL) {   // tell every CTA of the cluster
        float* remote = SPECGPU_MAP_SHARED(sdiag, tid);
        remote[2 * rank] = best;
        remote[2 * rank + 1] = __int_as_float(arg);
      }
      if (tid == 0) sdiag[2 * CL] = __int_as_float(arg);     // this CTA's own candidate row
    }
    __syncthreads();
    {
      // every CTA also ships its candidate row (row == column by symmetry) to all peers in the same step, so that one
      // cluster barrier settles both the choice and the start vector
      const int mine = __float_as_int(sdiag[2 * CL]);
      if (rl == mine - row0) {
#pragma unroll
        for (int r = 0; r < CL; ++r) {
          float* remote = SPECGPU_MAP_SHARED(scand, r) + rank * N;
#pragma unroll
          for (int j = 0; j < NJ; ++j) *reinterpret_cast<float4*>(remote + q * 4 + CS * j) = g[j];
        }
      }
    }
    SPECGPU_CLUSTER_SYNC();
    float best = sdiag[0];
    int arg = __float_as_int(sdiag[1]);
    int owner = 0;
#pragma unroll
    for (int r = 1; r < CL; ++r) {
      const float ob = sdiag[2 * r];
      const int oa = __float_as_int(sdiag[2 * r + 1]);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
        owner = r;
      }
    }
    if (tid < N) sx[tid] = scand[owner * N + tid];
    __syncthreads();
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < N / 32; ++i) ss += sx[lane + 32 * i] * sx[lane + 32 * i];
    ss = warp_sum(ss);
    const float sinv = (ss > 0.f && ss < INFINITY) ? rsqrtf(ss) : 0.f;
    __syncthreads();
    if (tid < N) sx[tid] *= sinv;
    __syncthreads();
  }
  float lambda = 0.f, prev_delta = INFINITY;
  int status = 1;
  for (int it = 0; it < a.max_iter; ++it) {
    float* syc = sy + (it & 1) * N;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 xv = *reinterpret_cast<const float4*>(sx + q * 4 + CS * j);
      acc += g[j].x * xv.x + g[j].y * xv.y + g[j].z * xv.z + g[j].w * xv.w;
    }
#pragma unroll
    for (int o = kGeTpr / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (q < CL) {   // thread q of the row sends y[row] to CTA q
      float* remote = SPECGPU_MAP_SHARED(syc, q);
      remote[row0 + rl] = acc;
    }
    SPECGPU_CLUSTER_SYNC();
    // every warp of every CTA redundantly reduces |y|^2, x.y and |y/|y| - x|^2 (identical values: uniform control flow)
    float yy = 0.f, xy = 0.f;
#pragma unroll
    for (int i = 0; i < N / 32; ++i) {
      const float y = syc[lane + 32 * i], x = sx[lane + 32 * i];
      yy += y * y;
      xy += x * y;
    }
    yy = warp_sum(yy);
    xy = warp_sum(xy);
    if (!(yy > 0.f) || !(yy < INFINITY)) break;   // zero / non-finite iterate: not convergence (status stays 1)
    const float inv = rsqrtf(yy);
    float dd = 0.f;
#pragma unroll
    for (int i = 0; i < N / 32; ++i) {
      const float d = syc[lane + 32 * i] * inv - sx[lane + 32 * i];
      dd += d * d;
    }
    dd = warp_sum(dd);
    lambda = xy;  // Rayleigh quotient x^T G x with |x| = 1
    __syncthreads();
    if (tid < N) sx[tid] = syc[tid] * inv;
    __syncthreads();
    if (dd < 1e-13f || (dd < 1e-10f && dd >= prev_delta)) {
      status = 0;
      break;
    }
    prev_delta = dd;
  }
  // nobody may leave while a peer can still write into its shared memory
  SPECGPU_CLUSTER_SYNC();
  if (tid < kGeRows) a.U[b * (int64_t)N * N + (int64_t)(row0 + tid) * N] = sx[row0 + tid];
  if (rank == 0 && tid == 0) {
    a.lam[b * N] = lambda * lam_scale;
    a.plan[b * 4 + 0] = 1;
    a.plan[b * 4 + 1] = N;
    a.plan[b * 4 + 2] = -1;
    a.plan[b * 4 + 3] = status;
    if (a.flagged != nullptr) a.flagged[b] = status;
  }
}

template <int N>
static int launch_gram_eig_t(const GramEigArgs& a, int64_t B, cudaStream_t stream) {
  constexpr int CL = N / kGeRows;
  const size_t smem = (size_t)(3 * N + 2 * CL + 8 + kGeRows * 132 + N + CL * N) * sizeof(float);
  SPECGPU_LAUNCH_PDL(gram_eig_kernel<N>, (unsigned)(B * CL), kGeThreads, smem, stream, CL, a);
  return (int)cudaGetLastError();
}

// ======================================================================================================
// The same step for 256-row matrices whose partials come one set per matrix (launch_gram_tma), as ONE CTA per matrix:
// the symmetric three quarters of G -- the partial layout itself, [128][388] = [G00 | G01 | G11 | row sums] -- are summed
// over the matrix's partials straight into 194 KB of shared memory (every load of a batch in flight before the first
// add), and the power iteration runs from there: row dots for G00 | G01 and G11, column sums for G10 = G01^T.  The
// min-max normalisation is never applied to the entries; it enters every product algebraically,
//   G' x = G x - m (r (1.x) + 1 (r.x)) + m^2 T (1.x) 1.
// No cluster, no distributed shared memory, no cluster barriers: the cluster kernel above spends 18.6 of its 20 us
// outside the (two, typically) iterations -- launching 320 clustered CTAs, three dependent rounds of partial loads, four
// cluster barriers.
// ======================================================================================================
#if !defined(SPECGPU_EMULATE)
constexpr int kGe1Threads = 1024, kGe1PP = 388, kGe1PW = 384;

__global__ void __launch_bounds__(kGe1Threads, 1) gram_eig1_kernel(GramEigArgs a) {
  constexpr int N = 256, H = 128, PP = kGe1PP, PW = kGe1PW;
  pdl_trigger();
  SPECGPU_DYN_SMEM(smem);
  float* P = reinterpret_cast<float*>(smem);     // [128][388]
  float* sx = P + H * PP;                        // [256] iterate
  float* sy = sx + N;                            // [256] G x
  float* sr = sy + N;                            // [256] row sums of the raw image
  float* scol = sr + N;                          // [8][128] partial column sums of G01
  float* sred = scol + 8 * H;                    // [64] reductions
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x;
  const int kparts = (int)((a.nchunk + a.per - 1) / a.per);
  const float4* part = reinterpret_cast<const float4*>(a.partial + (size_t)b * kparts * H * PP);
  constexpr int NV = H * PP / 4;                 // float4 per partial (12416)
  pdl_wait();
  // ---- P = sum of the partials, in partial order (deterministic); every load of a trip (3 positions x all partials) is
  //      in flight before the first add: five round trips for three partials instead of twelve ----
  auto sum_partials = [&](auto kp_c) {
    constexpr int KP = decltype(kp_c)::value;
    for (int v0 = tid; v0 < NV; v0 += 3 * kGe1Threads) {
      float4 w[KP][3];
#pragma unroll
      for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int v = v0 + u * kGe1Threads;
          w[k][u] = (v < NV) ? __ldg(part + (size_t)k * NV + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        float4 acc = w[0][u];
#pragma unroll
        for (int k = 1; k < KP; ++k) {
          acc.x += w[k][u].x; acc.y += w[k][u].y; acc.z += w[k][u].z; acc.w += w[k][u].w;
        }
        const int v = v0 + u * kGe1Threads;
        if (v < NV) reinterpret_cast<float4*>(P)[v] = acc;
      }
    }
  };
  if (a.debug & 1) {
  } else if (kparts == 3) {
    sum_partials(std::integral_constant<int, 3>{});
  } else if (kparts == 2) {
    sum_partials(std::integral_constant<int, 2>{});
  } else if (kparts == 1) {
    sum_partials(std::integral_constant<int, 1>{});
  } else if (kparts == 4) {
    sum_partials(std::integral_constant<int, 4>{});
  } else {
    for (int v = tid; v < NV; v += kGe1Threads) {
      float4 acc = __ldg(part + v);
      for (int k = 1; k < kparts; ++k) {
        const float4 w = __ldg(part + (size_t)k * NV + v);
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
      reinterpret_cast<float4*>(P)[v] = acc;
    }
  }
  __syncthreads();
  if (a.debug & 2) return;
  float m = 0.f, m2t = 0.f, lam_scale = 1.f;
  if (a.raw_minmax != nullptr) {
    m = minmax_get_min(a.raw_minmax, b);
    const float den = minmax_get_max(a.raw_minmax, b) - m;
    lam_scale = 1.0f / (den * den);
    m2t = m * m * (float)a.cols;
  }
  if (tid < N) sr[tid] = (a.raw_minmax != nullptr) ? P[(tid & (H - 1)) * PP + PW + (tid >> 7)] : 0.f;
  __syncthreads();
  // ---- start vector: the row of G' with the largest diagonal entry (ties: the lowest index) ----
  if (tid < N) {
    const float g = (tid < H) ? P[tid * PP + tid] : P[(tid - H) * PP + 2 * H + (tid - H)];
    sy[tid] = g - 2.f * m * sr[tid] + m2t;
  }
  __syncthreads();
  if (warp == 0) {
    float best = -INFINITY;
    int arg = 0;
    for (int i = lane; i < N; i += 32) {
      const float v = sy[i];
      if (v > best) {
        best = v;
        arg = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (lane == 0) sred[0] = __int_as_float(arg);
  }
  __syncthreads();
  {
    const int arow = __float_as_int(sred[0]);
    if (tid < N) {
      float g;     // G[arow][tid] out of the symmetric storage
      if (arow < H) g = (tid < H) ? P[arow * PP + tid] : P[arow * PP + tid];                    // [G00 | G01] row
      else g = (tid < H) ? P[tid * PP + H + (arow - H)] : P[(arow - H) * PP + 2 * H + (tid - H)];   // G01^T column | G11 row
      sx[tid] = g - m * (sr[arow] + sr[tid]) + m2t;
    }
    __syncthreads();
    float ss = 0.f;
    for (int i = lane; i < N; i += 32) ss += sx[i] * sx[i];
    ss = warp_sum(ss);
    const float sinv = (ss > 0.f && ss < INFINITY) ? rsqrtf(ss) : 0.f;
    __syncthreads();
    if (tid < N) sx[tid] *= sinv;
    __syncthreads();
  }
  float lambda = 0.f, prev_delta = INFINITY;
  int status = 1;
  for (int it = 0; it < a.max_iter; ++it) {
    // s1 = 1.x, s2 = r.x (every warp redundantly: identical values, uniform control flow)
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < N; i += 32) {
      s1 += sx[i];
      s2 += sr[i] * sx[i];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    // (a) rows 4 w .. 4 w + 3 of [G00 | G01] and of G11: lane l covers columns 4 l .. 4 l + 3 (+ 128)
    {
      const float4 xa = *reinterpret_cast<const float4*>(sx + 4 * lane);
      const float4 xb = *reinterpret_cast<const float4*>(sx + H + 4 * lane);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int i = 4 * warp + rr;
        const float4 g0 = *reinterpret_cast<const float4*>(P + i * PP + 4 * lane);
        const float4 g1 = *reinterpret_cast<const float4*>(P + i * PP + H + 4 * lane);
        const float4 g2 = *reinterpret_cast<const float4*>(P + i * PP + 2 * H + 4 * lane);
        float top = g0.x * xa.x + g0.y * xa.y + g0.z * xa.z + g0.w * xa.w + g1.x * xb.x + g1.y * xb.y + g1.z * xb.z + g1.w * xb.w;
        float bot = g2.x * xb.x + g2.y * xb.y + g2.z * xb.z + g2.w * xb.w;
        top = warp_sum(top);
        bot = warp_sum(bot);
        if (lane == 0) {
          sy[i] = top;
          sy[H + i] = bot;
        }
      }
    }
    // (b) G10 x_top = G01^T x_top: thread (c, part) sums rows 16 part .. +15 of column c of G01
    {
      const int c = tid & (H - 1), pt = tid >> 7;
      float cs = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) cs = fmaf(P[(16 * pt + i) * PP + H + c], sx[16 * pt + i], cs);
      scol[pt * H + c] = cs;
    }
    __syncthreads();
    if (tid < N) {
      float y = sy[tid];
      if (tid >= H) {
        const int c = tid - H;
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) y += scol[pt * H + c];
      }
      sy[tid] = y - m * (sr[tid] * s1 + s2) + m2t * s1;
    }
    __syncthreads();
    float yy = 0.f, xy = 0.f;
    for (int i = lane; i < N; i += 32) {
      const float y = sy[i], x = sx[i];
      yy += y * y;
      xy += x * y;
    }
    yy = warp_sum(yy);
    xy = warp_sum(xy);
    if (!(yy > 0.f) || !(yy < INFINITY)) break;   // zero / non-finite iterate: not convergence (status stays 1)
    const float inv = rsqrtf(yy);
    float dd = 0.f;
    for (int i = lane; i < N; i += 32) {
      const float d = sy[i] * inv - sx[i];
      dd += d * d;
    }
    dd = warp_sum(dd);
    lambda = xy;
    __syncthreads();
    if (tid < N) sx[tid] = sy[tid] * inv;
    __syncthreads();
    if (dd < 1e-13f || (dd < 1e-10f && dd >= prev_delta)) {
      status = 0;
      break;
    }
    prev_delta = dd;
  }
  __syncthreads();
  if (tid < N) a.U[b * (int64_t)N * N + (int64_t)tid * N] = sx[tid];
  if (tid == 0) {
    a.lam[b * N] = lambda * lam_scale;
    a.plan[b * 4 + 0] = 1;
    a.plan[b * 4 + 1] = N;
    a.plan[b * 4 + 2] = -1;
    a.plan[b * 4 + 3] = status;
    if (a.flagged != nullptr) a.flagged[b] = status;
  }
}
#endif

int launch_gram_eig(const float* partial, int64_t nchunk, int64_t per, int64_t B, int n, int max_iter, float* U, float* lam,
                    int32_t* plan, cudaStream_t stream, const MinMaxWord* raw_minmax, int64_t cols, int per_matrix,
                    int32_t* flagged) {
  if (B == 0) return 0;
  GramEigArgs a{partial, nchunk, per, raw_minmax, cols, per_matrix, max_iter > 0 ? max_iter : kPowMaxIter, U, lam, plan, flagged};
#if !defined(SPECGPU_EMULATE)
  static const bool one_cta = !(std::getenv("SPECGPU_GRAM_EIG1") && std::getenv("SPECGPU_GRAM_EIG1")[0] == '0');
  if (n == 256 && per_matrix && one_cta) {
    if (const char* env = std::getenv("SPECGPU_EIG1_DEBUG")) a.debug = std::atoi(env);
    const size_t smem = ((size_t)128 * kGe1PP + 3 * 256 + 8 * 128 + 64) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(gram_eig1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    SPECGPU_LAUNCH_PDL(gram_eig1_kernel, (unsigned)B, kGe1Threads, smem, stream, 1, a);
    return (int)cudaGetLastError();
  }
#endif
  if (n == 256) return launch_gram_eig_t<256>(a, B, stream);
  if (n == 128) return launch_gram_eig_t<128>(a, B, stream);
  return -1;
}

// ======================================================================================================
// Full symmetric eigen-decomposition: one-sided (Hestenes) Jacobi on the columns of G, resident in the
// shared memory of a thread-block cluster.  The n_pad columns are cut into 2*CL blocks of cb columns;
// CTA r of the cluster holds two blocks ("top", "bot") plus one spare slot.  A sweep is 2*CL-1
// block-rounds of the round-robin tournament; between block-rounds the blocks move one position round
// the ring through distributed shared memory (two cluster barriers).  Within a block-round every CTA
// orthogonalises its 2*cb columns pairwise (one warp per pair): the full local tournament in the first
// block-round of a sweep, only the cross pairs (top_i, bot_j) afterwards.  At convergence column k is
// lambda_k * u_k; eig_sort_kernel orders the pairs by descending lambda.
// ======================================================================================================
constexpr int kJacThreads = 1024;
constexpr int kJacMaxSweeps = 40;
template <class T>
struct JacTol {
  static constexpr float value = 1e-6f;
};
template <>
struct JacTol<double> {
  static constexpr double value = 1e-14;
};

template <class T>
struct JacobiArgs {
  const T* G;         // [B][n][n]
  int n, n_pad, cb, cl;
  const int32_t* plan;  // skip matrices whose plan[b][3] == 0 when skip_converged
  int skip_converged;
  T* Ucols;           // [B][n_pad][n_pad] : row k = (unsorted) eigenvector k (contiguous)
  T* lam_raw;         // [B][n_pad]
  int32_t* status;    // [B] sweeps used (negative: hit the cap)
};

// Orthogonalise columns x, y (length n, one warp).  Returns true if a rotation was applied.
template <class T>
__device__ __forceinline__ bool jacobi_pair(T* x, T* y, int n, int lane) {
  T a = 0, bq = 0, g = 0;
  for (int i = lane; i < n; i += 32) {
    const T xv = x[i], yv = y[i];
    a += xv * xv;
    bq += yv * yv;
    g += xv * yv;
  }
  a = warp_sum(a);
  bq = warp_sum(bq);
  g = warp_sum(g);
  // |g| > tol sqrt(a b)  <=>  g^2 > tol^2 a b   (no square root on the path every pair takes)
  if (!(g * g > ((T)JacTol<T>::value * (T)JacTol<T>::value) * (a * bq))) return false;
  const T zeta = (bq - a) / ((T)2 * g);
  const T az = zeta < 0 ? -zeta : zeta;
  const T t = (zeta < 0 ? (T)-1 : (T)1) / (az + (T)sqrt((double)((T)1 + zeta * zeta)));
  const T c = (T)rsqrt((double)((T)1 + t * t)), s = c * t;
  for (int i = lane; i < n; i += 32) {
    const T xv = x[i], yv = y[i];
    x[i] = c * xv - s * yv;
    y[i] = s * xv + c * yv;
  }
  return true;
}

template <class T>
__global__ void __launch_bounds__(kJacThreads) eig_jacobi_kernel(JacobiArgs<T> a) {
  pdl_trigger();
  pdl_wait();
  SPECGPU_DYN_SMEM(smem);
  const int n = a.n, np = a.n_pad, cb = a.cb, CL = a.cl;
  const int rank = SPECGPU_CLUSTER_RANK();
  const int64_t b = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kJacThreads / 32;
  const size_t slot_floats = (size_t)cb * np;
  T* slots = reinterpret_cast<T*>(smem);
  int* s_flag = reinterpret_cast<int*>(slots + 3 * slot_floats);  // [CL] per-rank "rotated" flags + [1] local
  // Slot roles.  Rank 0 never re-labels (top 0, bot 1, spare 2); every rank >= 1 applies the same
  // permutation after each ring step, so all of them share (p_top, p_bot, p_spare), which every rank
  // (rank 0 included) tracks in order to address its neighbours' slots.
  int p_top = 0, p_bot = 1, p_spare = 2;
  const bool skip = a.skip_converged && a.plan[b * 4 + 3] == 0;   // uniform over the whole cluster

  if (!skip) {
    // block ids: rank r starts with blocks 2r (top) and 2r+1 (bot); column j of G == row j (symmetric)
    const T* Gb = a.G + b * (int64_t)n * n;
    for (int s = 0; s < 2; ++s) {
      T* dst = slots + (size_t)s * slot_floats;   // initially top = slot 0, bot = slot 1 everywhere
      const int col0 = (2 * rank + s) * cb;
      for (int i = tid; i < cb * np; i += kJacThreads) {
        const int c = col0 + i / np, r = i % np;
        dst[i] = (c < n && r < n) ? Gb[(int64_t)c * n + r] : (T)0;
      }
    }
  }
  __syncthreads();
  int sweeps_used = 0;
  bool converged = skip;
  if (!skip) {
    for (int sweep = 0; sweep < kJacMaxSweeps && !converged; ++sweep) {
      if (tid == 0) s_flag[CL] = 0;
      __syncthreads();
      bool rotated = false;
      for (int br = 0; br < 2 * CL - 1; ++br) {
        const int top = (rank == 0) ? 0 : p_top, bot = (rank == 0) ? 1 : p_bot;
        T* Tp = slots + top * slot_floats;
        T* Bt = slots + bot * slot_floats;
        if (br == 0) {
          // full tournament over the m = 2*cb local columns (player m-1 fixed)
          const int m = 2 * cb;
          for (int r = 0; r < m - 1; ++r) {
            for (int i = warp; i < cb; i += NW) {
              int p, q;
              if (i == 0) {
                p = m - 1;
                q = r;
              } else {
                p = (r + i) % (m - 1);
                q = (r - i + (m - 1)) % (m - 1);
              }
              T* x = (p < cb) ? Tp + (size_t)p * np : Bt + (size_t)(p - cb) * np;
              T* y = (q < cb) ? Tp + (size_t)q * np : Bt + (size_t)(q - cb) * np;
              rotated |= jacobi_pair(x, y, np, lane);
            }
            __syncthreads();
          }
        } else {
          for (int r = 0; r < cb; ++r) {
            for (int i = warp; i < cb; i += NW) {
              T* x = Tp + (size_t)i * np;
              T* y = Bt + (size_t)((i + r) % cb) * np;
              rotated |= jacobi_pair(x, y, np, lane);
            }
            __syncthreads();
          }
        }
        if (CL > 1) {
          // ---- move the blocks one step round the ring (see header comment) ----
          // phase 1: rank 0 sends bot, ranks 1..CL-2 send top, to the right neighbour's spare slot
          if (rank < CL - 1) {
            const T* srcp = slots + (rank == 0 ? bot : top) * slot_floats;
            T* dstp = SPECGPU_MAP_SHARED(slots + p_spare * slot_floats, rank + 1);
            for (size_t i = tid; i < slot_floats; i += kJacThreads) dstp[i] = srcp[i];
          }
          SPECGPU_CLUSTER_SYNC();
          // phase 2: ranks 1..CL-1 send their old bot to the left neighbour's freed slot
          //          (rank 0: its old bot slot; rank i-1 >= 1: its old top slot)
          if (rank >= 1) {
            const T* srcp = slots + bot * slot_floats;
            const int left_free = (rank - 1 == 0) ? 1 : p_top;
            T* dstp = SPECGPU_MAP_SHARED(slots + left_free * slot_floats, rank - 1);
            for (size_t i = tid; i < slot_floats; i += kJacThreads) dstp[i] = srcp[i];
          }
          SPECGPU_CLUSTER_SYNC();
          {
            const int old_top = p_top, old_bot = p_bot;
            p_top = p_spare;     // received in phase 1
            p_bot = old_top;     // rank CL-1: its own old top; others: received in phase 2 into the old top slot
            p_spare = old_bot;   // sent away in phase 2
          }
          // rank 0: top fixed, bot slot refilled in phase 2, spare untouched
        }
      }
      // ---- did anybody rotate during this sweep? ----
      if (rotated && lane == 0) s_flag[CL] = 1;
      __syncthreads();
      int any = s_flag[CL];
      if (CL > 1) {
        if (tid < CL) {
          int* remote = SPECGPU_MAP_SHARED(s_flag, tid);
          remote[rank] = any;
        }
        SPECGPU_CLUSTER_SYNC();
        any = 0;
        for (int r = 0; r < CL; ++r) any |= s_flag[r];
        SPECGPU_CLUSTER_SYNC();
      }
      __syncthreads();   // everybody has read the flag before thread 0 clears it for the next sweep
      sweeps_used = sweep + 1;
      converged = (any == 0);
    }
    // ---- write lambda_k = |a_k|, u_k = a_k / |a_k| for the 2*cb local columns ----
    for (int s = 0; s < 2; ++s) {
      const int slot = (rank == 0) ? s : (s == 0 ? p_top : p_bot);
      const T* src = slots + slot * slot_floats;
      for (int i = warp; i < cb; i += NW) {
        const T* x = src + (size_t)i * np;
        T nn = 0;
        for (int r = lane; r < np; r += 32) nn += x[r] * x[r];
        nn = warp_sum(nn);
        const T nrm = (T)sqrt((double)nn);
        const T inv = nrm > 0 ? (T)1 / nrm : (T)0;
        // any slot order is fine: eig_sort_kernel orders by lambda
        const int64_t k = (int64_t)(2 * rank + s) * cb + i;
        T* u = a.Ucols + (b * np + k) * np;
        for (int r = lane; r < np; r += 32) u[r] = x[r] * inv;
        if (lane == 0) a.lam_raw[b * np + k] = nrm;
      }
    }
    if (rank == 0 && tid == 0) a.status[b] = converged ? sweeps_used : -sweeps_used;
  }
}

// Order eigenpairs by descending lambda (rank by counting; ties by index) and lay U out as
// U[b][i][k] (row-major n x n, column k = k-th largest).  One CTA per matrix.
template <class T>
__global__ void eig_sort_kernel(const T* Ucols, const T* lam_raw, const int32_t* jstatus, int n, int n_pad,
                                int skip_converged, float* U, float* lam, int32_t* plan) {
  const int64_t b = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  if (skip_converged && plan[b * 4 + 3] == 0) return;
  __shared__ int s_rank[512];
  const T* lr = lam_raw + b * n_pad;
  for (int k = threadIdx.x; k < n_pad; k += blockDim.x) {
    const T v = lr[k];
    int rk = 0;
    for (int j = 0; j < n_pad; ++j) {
      const T w = lr[j];
      rk += (w > v || (w == v && j < k)) ? 1 : 0;
    }
    s_rank[k] = rk;
    if (rk < n) lam[b * n + rk] = (float)v;
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < (int64_t)n_pad * n; i += blockDim.x) {
    const int k = (int)(i / n), r = (int)(i % n);
    const int rk = s_rank[k];
    if (rk < n) U[b * (int64_t)n * n + (int64_t)r * n + rk] = (float)Ucols[(b * n_pad + k) * n_pad + r];
  }
  if (threadIdx.x == 0) plan[b * 4 + 3] = (jstatus[b] > 0) ? 0 : 1;
}

struct JacobiGeom {
  int n_pad, cb, cl;
  size_t smem;
};

// fp64 columns are twice as large: twice the cluster size keeps the three slots of a CTA under 227 KB.
// (n > 256 in double would need a non-portable cluster of 16: that size stays in float.)
static bool jacobi_f64_supported(int n) { return n <= 256; }

static JacobiGeom jacobi_geom(int n, int f64) {
  JacobiGeom g;
  if (f64) g.cl = (n <= 64) ? 1 : (n <= 128 ? 2 : 4);
  else g.cl = (n <= 128) ? 1 : (n <= 256 ? 2 : 8);
  const int nb = 2 * g.cl;
  g.cb = (n + nb - 1) / nb;
  g.n_pad = g.cb * nb;
  g.smem = 3 * (size_t)g.cb * g.n_pad * (f64 ? sizeof(double) : sizeof(float)) + (g.cl + 2) * sizeof(int) + 16;
  return g;
}

size_t jacobi_workspace_bytes(int64_t B, int n) {
  const JacobiGeom g = jacobi_geom(n, 0);
  const JacobiGeom g2 = jacobi_geom(std::min(n, 256), 1);
  const size_t np = (size_t)std::max(g.n_pad, g2.n_pad);
  return (size_t)B * np * np * sizeof(double) + (size_t)B * np * sizeof(double) + (size_t)B * sizeof(int32_t) + 1024;
}

template <class T>
static int launch_eig_jacobi_t(const T* G, int64_t B, int n, int skip_converged, float* U, float* lam, int32_t* plan,
                               void* ws, cudaStream_t stream) {
  const JacobiGeom g = jacobi_geom(n, sizeof(T) == 8);
  char* w = static_cast<char*>(ws);
  T* Ucols = reinterpret_cast<T*>(w);
  w += (size_t)B * g.n_pad * g.n_pad * sizeof(T);
  T* lam_raw = reinterpret_cast<T*>(w);
  w += (((size_t)B * g.n_pad * sizeof(T)) + 255) & ~(size_t)255;
  int32_t* jstatus = reinterpret_cast<int32_t*>(w);
  JacobiArgs<T> a{G, n, g.n_pad, g.cb, g.cl, plan, skip_converged, Ucols, lam_raw, jstatus};
#ifndef SPECGPU_EMULATE
  cudaError_t e = cudaFuncSetAttribute(eig_jacobi_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
#endif
  SPECGPU_LAUNCH_PDL(eig_jacobi_kernel<T>, (unsigned)(B * g.cl), kJacThreads, g.smem, stream, g.cl, a);
  if (cudaPeekAtLastError() != cudaSuccess) return (int)cudaGetLastError();
  SPECGPU_LAUNCH_PDL(eig_sort_kernel<T>, (unsigned)B, 512, 0, stream, 1, (const T*)Ucols, (const T*)lam_raw, (const int32_t*)jstatus, n,
                 g.n_pad, skip_converged, U, lam, plan);
  return (int)cudaGetLastError();
}

bool eig_jacobi_f64_supported(int n) { return jacobi_f64_supported(n); }

int launch_eig_jacobi(const void* G, int g_f64, int64_t B, int n, int skip_converged, float* U, float* lam, int32_t* plan,
                      void* ws, cudaStream_t stream) {
  if (B == 0) return 0;
  if (n > 512) return -1;
  if (g_f64) {
    if (!jacobi_f64_supported(n)) return -1;
    return launch_eig_jacobi_t<double>((const double*)G, B, n, skip_converged, U, lam, plan, ws, stream);
  }
  return launch_eig_jacobi_t<float>((const float*)G, B, n, skip_converged, U, lam, plan, ws, stream);
}

// ======================================================================================================
// Plan: singular values, median, optimal hard threshold count, start/stop with the reference's clamps and
// Python slice semantics.  kind 0: explicit (start, stop); 1: use_optimal; 2: computeSignal (1, 2*num_sing).
// ======================================================================================================
__global__ void svd_plan_kernel(const float* lam, int n, int kind, int start, int stop, double omega, int32_t* plan,
                                float* s_out) {
  const int64_t b = blockIdx.x;
  __shared__ float s_s[512];
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  if (kind != 0 || s_out != nullptr) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const float s = sqrtf(fmaxf(lam[b * n + k], 0.f));
      s_s[k] = s;
      if (s_out != nullptr) s_out[b * n + k] = s;
    }
  }
  __syncthreads();
  int a = start, e = stop, num_sing = -1;
  if (kind != 0) {
    // np.median of the descending s: mean of the two middle values for even n
    const float med = (n & 1) ? s_s[n / 2] : __fdiv_rn(__fadd_rn(s_s[n / 2 - 1], s_s[n / 2]), 2.0f);
    // beta = np.min(shape) / np.max(shape) is an np.float64, so omega(beta) and t* = omega * median are float64 and
    // `s > t*` compares in float64 (denoising_by_svd.ipynb:210-215)
    const double t_star = __dmul_rn(omega, (double)med);
    int cnt = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) cnt += ((double)s_s[k] > t_star) ? 1 : 0;
    atomicAdd(&s_cnt, cnt);
    __syncthreads();
    num_sing = s_cnt;
    if (kind == 1) {
      a = 0;
      e = num_sing - 1;
    } else {
      a = 1;
      e = 2 * num_sing;
    }
  }
  if (threadIdx.x == 0) {
    if (a < 0) a = 0;          // "prevent some bad values from being used"
    if (e > n) e = n;
    // python slice u[:, a:e]
    if (e < 0) e = (e + n > 0) ? e + n : 0;
    if (a > n) a = n;
    plan[b * 4 + 0] = a;
    plan[b * 4 + 1] = e;
    plan[b * 4 + 2] = num_sing;
  }
}

int launch_svd_plan(const float* lam, int64_t B, int n, int kind, int start, int stop, double omega, int32_t* plan,
                    float* s_out, cudaStream_t stream) {
  if (B == 0) return 0;
  if (n > 512) return -1;
  SPECGPU_LAUNCH(svd_plan_kernel, (unsigned)B, 256, 0, stream, lam, n, kind, start, stop, omega, plan, s_out);
  return (int)cudaGetLastError();
}

// ======================================================================================================
// Projection: out = sum_{k in [a,e)} u_k (u_k^T S), or S minus the complementary sum when that is the
// shorter one.  A CTA owns a [rows x 32] column tile of S in shared memory; each warp takes one vector
// of a chunk of 8 for the coefficient pass (lane = column) and rows warp, warp+8, ... for the update.
// ======================================================================================================
constexpr int kRecThreads = 256, kRecCols = 32, kRecChunk = 8;

template <class OutT, int RPW>   // RPW = rows per warp-thread = ceil(rows / 8)
__global__ void __launch_bounds__(kRecThreads) svd_project_kernel(const float* S, int rows, int64_t cols, int64_t ld,
                                                                  const float* U, const int32_t* plan, int clip,
                                                                  OutT* out, int64_t ldo) {
  SPECGPU_DYN_SMEM(smem);
  float* tile = reinterpret_cast<float*>(smem);                   // [rows][33]
  float* s_u = tile + (size_t)rows * (kRecCols + 1);               // [kRecChunk][rows]
  float* s_w = s_u + (size_t)kRecChunk * rows;                     // [kRecChunk][32]
  const int64_t b = blockIdx.y;
  const int64_t c0 = (int64_t)blockIdx.x * kRecCols;
  const int ncol = (int)((cols - c0 < kRecCols) ? (cols - c0) : kRecCols);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Sb = S + b * rows * ld;
  const float* Ub = U + b * (int64_t)rows * rows;
  int a = plan[b * 4 + 0], e = plan[b * 4 + 1];
  if (e < a) e = a;
  const int nk = e - a;
  const bool complement = (rows - nk) < nk;   // subtract the vectors outside [a, e)
  const int nvec = complement ? rows - nk : nk;

  for (int r = warp; r < rows; r += kRecThreads / 32)
    tile[r * (kRecCols + 1) + lane] = (lane < ncol) ? Sb[(int64_t)r * ld + c0 + lane] : 0.f;

  float acc[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) acc[i] = 0.f;

  for (int v0 = 0; v0 < nvec; v0 += kRecChunk) {
    __syncthreads();   // tile ready / previous chunk consumed
    // stage up to 8 vectors: index v -> eigen-column k
    for (int i = tid; i < kRecChunk * rows; i += kRecThreads) {
      const int vv = i / rows, r = i % rows;
      const int v = v0 + vv;
      float val = 0.f;
      if (v < nvec) {
        const int k = complement ? (v < a ? v : v + nk) : a + v;
        val = Ub[(int64_t)r * rows + k];
      }
      s_u[vv * rows + r] = val;
    }
    __syncthreads();
    {  // coefficients w[vv][col] = u^T tile[:, col]
      const float* u = s_u + warp * rows;
      float w = 0.f;
      if (v0 + warp < nvec)
        for (int r = 0; r < rows; ++r) w += u[r] * tile[r * (kRecCols + 1) + lane];
      s_w[warp * 32 + lane] = w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = warp + 8 * i;
      if (r < rows) {
        float s = 0.f;
#pragma unroll
        for (int vv = 0; vv < kRecChunk; ++vv) s += s_u[vv * rows + r] * s_w[vv * 32 + lane];
        acc[i] += s;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + 8 * i;
    if (r < rows && lane < ncol) {
      float v = complement ? tile[r * (kRecCols + 1) + lane] - acc[i] : acc[i];
      if (clip && v < 0.f) v = 0.f;
      out[(b * rows + r) * ldo + c0 + lane] = (OutT)v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// The default denoiseSignal range (start = 1, stop = len(s)) removes only the leading component:
//   D = S - u0 (u0^T S).
// This streaming kernel fuses it with the min-max normalisation of the log image and the clip, so the
// pipeline reads the log image once and writes S and D once.  A CTA owns a [rows x 32] column tile held in
// registers; warp w reduces rows w, w+16, ... for the 32 coefficients, then updates the same rows.
// ------------------------------------------------------------------------------------------------------
constexpr int kR1Warps = 16, kR1Threads = kR1Warps * 32;

// STREAM (pipeline): L was left in L2 by the producer for this, its last, reader -- the loads mark their lines evict_first
// and the S / D stores stream out (evict_first) so that they do not push out lines of L that are still to be read.
template <bool STREAM>
__device__ __forceinline__ float r1_load(const float* p, uint64_t pol) {
#if !defined(SPECGPU_EMULATE)
  if (STREAM) return ld_global_hint(p, pol);
#endif
  (void)pol;
  return __ldg(p);
}
template <bool STREAM>
__device__ __forceinline__ void r1_store(float* p, float v, uint64_t pol) {
#if !defined(SPECGPU_EMULATE)
  if (STREAM) {
    st_global_hint(p, v, pol);
    return;
  }
#endif
  (void)pol;
  *p = v;
}

template <int RPW, bool NORM, bool CLIP, bool STREAM, bool TILES = false>   // RPW = rows per thread = ceil(rows / 16)
__device__ __forceinline__ void svd_rank1_tile(const int64_t b, const float* L, int rows, int64_t cols, int64_t ld,
                                               const MinMaxWord* minmax, const float* U, float* S, float* D, int64_t ldo,
                                               const R1Tiles tl) {
  uint64_t pol = 0;
#if !defined(SPECGPU_EMULATE)
  if (STREAM) pol = l2_policy_evict_first();
#endif
  SPECGPU_DYN_SMEM(smem);
  float* s_u = reinterpret_cast<float*>(smem);          // [rows]
  float* s_w = s_u + rows;                              // [16][32] partial coefficients
  const int64_t c0 = (int64_t)blockIdx.x * kRecCols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mn = 0.f, den = 1.f;
  if (NORM) {
    mn = minmax_get_min(minmax, b);
    den = minmax_get_max(minmax, b) - mn;
  }
  const float inv = 1.0f / den;
  for (int r = tid; r < rows; r += kR1Threads) s_u[r] = U[b * (int64_t)rows * rows + (int64_t)r * rows];
  // row-major source, or the tiled scratch image (ld < 0): tile blockIdx.x of matrix b is one contiguous [rows x 32] block
  const float* pl = L + img_off(b, warp, c0 + lane, rows, ld);
  float* ps = S + (b * rows + warp) * ldo + c0 + lane;
  float* pd = D + (b * rows + warp) * ldo + c0 + lane;
  const int64_t stepl = kR1Warps * (ld < 0 ? (int64_t)kTileCols : ld), stepo = kR1Warps * ldo;
  const bool write_s = S != nullptr && (NORM || S != L);
  // tile copy of D: this lane's column inside its tile (columns past the last tile are not exported)
  float* pt = nullptr;
  int64_t stept = 0;
  if (TILES && c0 + lane < (int64_t)tl.ntiles * tl.tile_w) {
    const int64_t t = (c0 + lane) / tl.tile_w;
    pt = tl.ptr + ((b * tl.ntiles + t) * rows + warp) * tl.tile_w + ((c0 + lane) - t * tl.tile_w);
    stept = (int64_t)kR1Warps * tl.tile_w;
  }
  float x[RPW];
  if (c0 + kRecCols <= cols && rows == kR1Warps * RPW) {
    // ---- full tile: no per-element predicates ----
#pragma unroll
    for (int i = 0; i < RPW; ++i) x[i] = r1_load<STREAM>(pl + i * stepl, pol);
    __syncthreads();
    float u[RPW];
    float w = 0.f;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      u[i] = s_u[warp + kR1Warps * i];
      if (NORM) x[i] = div_by(x[i] - mn, den, inv);
      w = fmaf(u[i], x[i], w);
    }
    s_w[warp * 32 + lane] = w;
    if (write_s) {
#pragma unroll
      for (int i = 0; i < RPW; ++i) r1_store<STREAM>(ps + i * stepo, x[i], pol);
    }
    __syncthreads();
    w = 0.f;
#pragma unroll
    for (int k = 0; k < kR1Warps; ++k) w += s_w[k * 32 + lane];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      float v = fmaf(-u[i], w, x[i]);
      if (CLIP) v = (v < 0.f) ? 0.f : v;   // NaN stays NaN, like hacked[hacked < 0] = 0
      r1_store<STREAM>(pd + i * stepo, v, pol);
      if (TILES && pt != nullptr) r1_store<STREAM>(pt + i * stept, v, pol);
    }
    return;
  }
  // ---- ragged tile (last columns / rows not a multiple of 16): this thread's rows are warp, warp+16, ...; nr of them
  //      are inside the matrix (warp-uniform) ----
  const bool col_ok = c0 + lane < cols;
  const int nr = col_ok ? ((rows - warp + kR1Warps - 1) / kR1Warps) : 0;
#pragma unroll
  for (int i = 0; i < RPW; ++i) x[i] = (i < nr) ? __ldg(pl + i * stepl) : 0.f;
  __syncthreads();
  float w = 0.f;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    if (NORM) x[i] = (i < nr) ? div_by(x[i] - mn, den, inv) : 0.f;
    if (i < nr) w = fmaf(s_u[warp + kR1Warps * i], x[i], w);
  }
  s_w[warp * 32 + lane] = w;
  if (write_s) {
#pragma unroll
    for (int i = 0; i < RPW; ++i)
      if (i < nr) ps[i * stepo] = x[i];
  }
  __syncthreads();
  w = 0.f;
#pragma unroll
  for (int k = 0; k < kR1Warps; ++k) w += s_w[k * 32 + lane];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    if (i < nr) {
      float v = fmaf(-s_u[warp + kR1Warps * i], w, x[i]);
      if (CLIP) v = (v < 0.f) ? 0.f : v;
      pd[i * stepo] = v;
      if (TILES && pt != nullptr) pt[i * stept] = v;
    }
  }
}

// only_flagged == nullptr: CTA (x, y) does column tile x of matrix y.  Repair pass (only_flagged != nullptr): a SMALL grid
// (a big one costs microseconds even when every CTA returns at once) whose CTAs walk the matrices and redo the flagged ones.
template <int RPW, bool NORM, bool CLIP, bool STREAM, bool TILES = false>
__global__ void __launch_bounds__(kR1Threads, 2) svd_rank1_kernel(const float* L, int rows, int64_t cols, int64_t ld,
                                                                  const MinMaxWord* minmax, const float* U, float* S,
                                                                  float* D, int64_t ldo, const int32_t* only_flagged, int64_t B,
                                                                  const R1Tiles tl) {
  pdl_wait();      // (a multi-wave grid: dependents are released when its CTAs exit)
  if (tl.info != nullptr && blockIdx.x == 0 && threadIdx.x < 4) {
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) tl.info[b * 4 + threadIdx.x] = tl.plan[b * 4 + threadIdx.x];
  }
  if (only_flagged == nullptr) {
    svd_rank1_tile<RPW, NORM, CLIP, STREAM, TILES>(blockIdx.y, L, rows, cols, ld, minmax, U, S, D, ldo, tl);
    return;
  }
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    if (only_flagged[b] == 0) continue;      // uniform over the CTA
    svd_rank1_tile<RPW, NORM, CLIP, STREAM, TILES>(b, L, rows, cols, ld, minmax, U, S, D, ldo, tl);
    __syncthreads();                         // the tile's shared scratch is reused
  }
}

template <int RPW, bool STREAM>
static void launch_rank1_ts(dim3 grid, size_t smem, cudaStream_t stream, const float* L, int rows, int64_t cols, int64_t ld,
                            const MinMaxWord* minmax, const float* U, int clip, float* S, float* D, int64_t ldo,
                            const int32_t* only_flagged, int64_t B, R1Tiles tl) {
  const bool nrm = minmax != nullptr;
#define SPECGPU_R1(NRM, CL, TL) \
  SPECGPU_LAUNCH_PDL((svd_rank1_kernel<RPW, NRM, CL, STREAM, TL>), grid, kR1Threads, smem, stream, 1, L, rows, cols, ld, minmax, U, S, D, ldo, only_flagged, B, tl)
  if (tl.ptr != nullptr) {    // the instantiations without tiles stay what they were (the extra pointer costs the plain path 2.6 us)
    if (nrm && clip) SPECGPU_R1(true, true, true);
    else if (nrm) SPECGPU_R1(true, false, true);
    else if (clip) SPECGPU_R1(false, true, true);
    else SPECGPU_R1(false, false, true);
  } else {
    if (nrm && clip) SPECGPU_R1(true, true, false);
    else if (nrm) SPECGPU_R1(true, false, false);
    else if (clip) SPECGPU_R1(false, true, false);
    else SPECGPU_R1(false, false, false);
  }
#undef SPECGPU_R1
}
template <int RPW>
static void launch_rank1_t(dim3 grid, size_t smem, cudaStream_t stream, const float* L, int rows, int64_t cols, int64_t ld,
                           const MinMaxWord* minmax, const float* U, int clip, float* S, float* D, int64_t ldo, int stream_out,
                           const int32_t* only_flagged, int64_t B, R1Tiles tl) {
  if (stream_out) launch_rank1_ts<RPW, true>(grid, smem, stream, L, rows, cols, ld, minmax, U, clip, S, D, ldo, only_flagged, B, tl);
  else launch_rank1_ts<RPW, false>(grid, smem, stream, L, rows, cols, ld, minmax, U, clip, S, D, ldo, only_flagged, B, tl);
}

int launch_svd_rank1(const float* L, int64_t B, int rows, int64_t cols, int64_t ld, const MinMaxWord* minmax, const float* U,
                     int clip, float* S, float* D, int64_t ldo, cudaStream_t stream, int stream_out, const int32_t* only_flagged,
                     R1Tiles tl) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  const size_t smem = ((size_t)rows + kR1Warps * 32) * sizeof(float);
  const dim3 grid((unsigned)ceil_div(cols, kRecCols), (unsigned)(only_flagged ? std::min<int64_t>(B, 2) : B));
  if (rows <= 64) launch_rank1_t<4>(grid, smem, stream, L, rows, cols, ld, minmax, U, clip, S, D, ldo, stream_out, only_flagged, B, tl);
  else if (rows <= 128) launch_rank1_t<8>(grid, smem, stream, L, rows, cols, ld, minmax, U, clip, S, D, ldo, stream_out, only_flagged, B, tl);
  else if (rows <= 256) launch_rank1_t<16>(grid, smem, stream, L, rows, cols, ld, minmax, U, clip, S, D, ldo, stream_out, only_flagged, B, tl);
  else return -1;
  return (int)cudaGetLastError();
}

template <class OutT>
static int launch_project_t(const float* S, int64_t B, int rows, int64_t cols, int64_t ld, const float* U,
                            const int32_t* plan, int clip, OutT* out, int64_t ldo, cudaStream_t stream) {
  const size_t smem = ((size_t)rows * (kRecCols + 1) + (size_t)kRecChunk * rows + kRecChunk * 32) * sizeof(float);
  const dim3 grid((unsigned)ceil_div(cols, kRecCols), (unsigned)B);
#define SPECGPU_PROJECT_CASE(RPW)                                                                             \
  {                                                                                                           \
    auto kern = svd_project_kernel<OutT, RPW>;                                                                \
    if (smem > 48 * 1024) {                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
      if (e != cudaSuccess) return (int)e;                                                                    \
    }                                                                                                         \
    SPECGPU_LAUNCH(kern, grid, kRecThreads, smem, stream, S, rows, cols, ld, U, plan, clip, out, ldo);        \
  }
  if (rows <= 64) SPECGPU_PROJECT_CASE(8)
  else if (rows <= 128) SPECGPU_PROJECT_CASE(16)
  else if (rows <= 256) SPECGPU_PROJECT_CASE(32)
  else if (rows <= 512) SPECGPU_PROJECT_CASE(64)
  else return -1;
#undef SPECGPU_PROJECT_CASE
  return (int)cudaGetLastError();
}

int launch_svd_project(const float* S, int64_t B, int rows, int64_t cols, int64_t ld, const float* U,
                       const int32_t* plan, int clip, void* out, int out_f64, int64_t ldo, cudaStream_t stream) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  if (out_f64) return launch_project_t<double>(S, B, rows, cols, ld, U, plan, clip, (double*)out, ldo, stream);
  return launch_project_t<float>(S, B, rows, cols, ld, U, plan, clip, (float*)out, ldo, stream);
}

}  // namespace specgpu
