// Values-first symmetric eigensolver for the `use_optimal` / `computeSignal` modes of denoiseSignal
// (spec_denoising/denoising_by_svd.ipynb:161-186, 210-217): those modes need ALL singular values (np.median, the count
// above the Gavish-Donoho threshold) but only the few leading vectors.  The cluster Jacobi solver (svd.cu) delivers all
// 256 vectors in ~20 sweeps and needs 4 SMs per matrix (two waves for 40 matrices: 16 ms); here, in float64,
//   1. tridiag_cluster_kernel  Householder tridiagonalisation with G resident in the distributed shared memory of a
//                       small cluster (rows dealt round-robin to the CTAs), ONE cluster barrier per column
//   2. bisect_kernel    all eigenvalues of the tridiagonal matrix by Sturm-count bisection, one thread per eigenvalue,
//                       division-free (scaled three-term recurrence)
//   3. (svd_plan_kernel, svd.cu: median, threshold, num_sing, start/stop)
//   4. trivec_kernel    the L <= 16 leading eigenvectors by inverse iteration on the tridiagonal matrix (pivoted LU as in
//                       dgttrf/dgttrs, one warp per vector)
//   5. triback_kernel   modified Gram-Schmidt among them, back-transformation through the Householder reflectors (staged
//                       through shared memory 16 at a time), residual check |G z - lambda z| against the ORIGINAL G,
//                       U[:, k] written; a matrix whose plan needs other vectors (trailing ones, more than 16) or whose
//                       check fails is flagged (plan[b][3] = 1) and redone by the Jacobi solver, which skips the others.
// Plain C++ under SPECGPU_EMULATE as well (no tensor-core / TMA instructions here).
#include "cluster.cuh"
#include "kernels.h"

namespace specgpu {

#ifdef SPECGPU_EMULATE
constexpr int kTcThreads = 64;      // one OS thread per CUDA thread (x cluster size) in the emulator
#else
constexpr int kTcThreads = 512;
#endif
constexpr int kTcMaxCl = 4;
constexpr int kTriMaxVec = 16;
constexpr int kTriBackThreads = 32 * kTriMaxVec;
constexpr int kTriMaxN = 256;
constexpr int kTriChunk = 16;       // reflectors staged per round of the back-transformation

__device__ __forceinline__ int hi32(double x) { return (int)(__double_as_longlong(x) >> 32); }
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// ======================================================================================================
// 1. Householder tridiagonalisation.  CTA `rank` of a cluster of CL owns rows rank, rank + CL, ... of the matrix and keeps
// them in REGISTERS for the whole kernel: thread (g, j) holds column j of the local rows g, g + NG, ... (43 doubles for
// n = 256, CL = 3, 512 threads).  Column step k, every CTA:
//   a. has x = row k (columns >= k) after all earlier updates; sigma, alpha, beta and the Householder vector v follow
//      redundantly (same data, same order of operations: bit-identical in every CTA)
//   b. w_j = sum over ITS rows i > k of A[i][j] v_i  -- by symmetry the partial sums over the CTAs (and row groups) add
//      up to (A v)_j; a thread walks down its own column: no shuffles, no shared-memory traffic for the matrix
//   c. pushes w to every CTA of the cluster; the owner of row k + 1 also pushes that row as it stands   -> cluster barrier
//   d. p = beta * sum of the partials, K = beta/2 v.p, q = p - K v; x for the next step = row k+1 - v_{k+1} q - q_{k+1} v
//   e. A[i][j] -= v_i q_j + q_i v_j on its own rows.
// One cluster barrier and six CTA barriers per column.  Push targets are double-buffered on k & 1: a CTA that is one
// step ahead writes the other half.  (v_i, q_i) of the local rows are kept as pairs in row-group order so that the inner
// loops read them with one 16-byte broadcast load.
// v_k is left in W[k][k+1..n) (the layout the back-transformation reads); d, e (e[k] couples k and k+1) and beta [B][n].
// ======================================================================================================
struct TriClArgs {
  const double* G;
  double* W;
  double* d;
  double* e;
  double* beta;
  int n;
  int cl;
};

constexpr int kTcRowCap = 86;        // local rows per CTA the register tile can hold
template <int T>
struct TcGeom {
  static constexpr int CW = T < kTriMaxN ? T : kTriMaxN;    // column lanes
  static constexpr int NG = T / CW;                          // row groups
  static constexpr int NCOL = kTriMaxN / CW;                 // columns per thread
  static constexpr int UMAX = (kTcRowCap + NG - 1) / NG;     // rows per thread
};

template <int T>
static size_t tri_cluster_smem(int n) {
  using Ge = TcGeom<T>;
  // x, v, q [n]; vq [NG][UMAX][2]; wpart [2][kTcMaxCl][NG][n]; rowbuf [2][n]; red [T/32]
  return ((size_t)3 * n + 2 * Ge::NG * Ge::UMAX + 2 * kTcMaxCl * Ge::NG * (size_t)n + 2 * (size_t)n + T / 32) * sizeof(double);
}

// all threads get the sum; two barriers, the first of which also orders whatever the caller wrote before
template <int T>
__device__ __forceinline__ double tc_block_sum(double v, double* red, int tid) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = ((tid & 31) < T / 32) ? red[tid & 31] : 0.0;
  return warp_sum(s);      // xor butterfly: every lane of every warp adds in the same order
}

template <int T>
__global__ void __launch_bounds__(T) tridiag_cluster_kernel(TriClArgs a) {
  SPECGPU_DYN_SMEM(smem_raw);
  using Ge = TcGeom<T>;
  constexpr int CW = Ge::CW, NG = Ge::NG, NCOL = Ge::NCOL, UMAX = Ge::UMAX;
  const int n = a.n, CL = a.cl;
  const int rank = (CL > 1) ? SPECGPU_CLUSTER_RANK() : 0;
  const int64_t b = blockIdx.x / CL;
  const int tid = threadIdx.x, cj = tid % CW, g = tid / CW;
  double* x = reinterpret_cast<double*>(smem_raw);
  double* v = x + n;
  double* q = v + n;
  double* vq = q + n;                                   // [NG][UMAX][2]
  double* wpart = vq + 2 * NG * UMAX;                   // [2][kTcMaxCl][NG][n]
  double* rowbuf = wpart + 2 * kTcMaxCl * NG * n;       // [2][n]
  double* red = rowbuf + 2 * n;
  const double* G = a.G + b * (int64_t)n * n;
  double* W = a.W + b * (int64_t)n * n;
  double* d = a.d + b * n;
  double* e = a.e + b * n;
  double* be = a.beta + b * n;
  const int nloc = (n - rank + CL - 1) / CL;            // my rows: i = rank + CL * lr, lr = g + NG * u
  double A[NCOL][UMAX];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) {
    const int j = cj + CW * c;
#pragma unroll
    for (int u = 0; u < UMAX; ++u) {
      const int lr = g + NG * u;
      A[c][u] = (lr < nloc && j < n) ? G[(int64_t)(rank + CL * lr) * n + j] : 0.0;
    }
  }
  double part = 0.0;
  for (int j = tid; j < n; j += T) {
    const double xj = G[j];
    x[j] = xj;
    v[j] = 0.0;
    q[j] = 0.0;
    if (j >= 1) part += xj * xj;
  }
  for (int idx = tid; idx < 2 * NG * UMAX; idx += T) vq[idx] = 0.0;
  if (CL > 1) SPECGPU_CLUSTER_SYNC();                   // every CTA of the cluster is resident before the first push
  for (int k = 0; k + 2 < n; ++k) {
    const int par = k & 1;
    // a. the reflector (part = this thread's share of sum_{j > k} x_j^2)
    const double sigma = tc_block_sum<T>(part, red, tid);
    const double x0 = x[k + 1];
    const bool refl = (sigma - x0 * x0) > 0.0;          // otherwise the column is already tridiagonal: identity
    double alpha = x0, beta = 0.0;
    if (refl) {
      alpha = (x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma);
      beta = 1.0 / (sigma - alpha * x0);                // 2 / (v^T v) with v = x - alpha e_0
    }
    for (int j = k + 1 + tid; j < n; j += T) {
      const double vj = refl ? ((j == k + 1) ? x0 - alpha : x[j]) : 0.0;
      v[j] = vj;
      if (j % CL == rank) {
        const int lr = j / CL;
        vq[((lr % NG) * UMAX + lr / NG) * 2] = vj;
      }
      if (rank == k % CL) W[(int64_t)k * n + j] = vj;
    }
    if (rank == 0 && tid == 0) {
      d[k] = x[k];
      e[k] = alpha;
      be[k] = beta;
    }
    __syncthreads();
    // b. partial matrix-vector product over my rows, c. push
    const int lr0 = (k + CL - rank) / CL;               // first local row with global index >= k + 1
    const int u0 = (lr0 - g + NG - 1) / NG;             // (lr0 >= 0, g < NG: the numerator is never below -(NG-1)+NG-1)
    const int u1 = (nloc - g + NG - 1) / NG;            // rows u0 <= u < u1 of this thread are live
    const bool own_next = rank == (k + 1) % CL;
    const int lrn = (k + 1) / CL;
    const bool my_next = own_next && (lrn % NG) == g;
    const int un = lrn / NG;
    const double* vqg = vq + g * UMAX * 2;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      const int j = cj + CW * c;
      if (j > k && j < n) {
        double acc0 = 0.0, acc1 = 0.0, rowv = 0.0;
        if (refl) {
#pragma unroll
          for (int u = 0; u < UMAX; ++u) {
            if (u >= u0 && u < u1) {
              if (u & 1) acc1 += A[c][u] * vqg[2 * u];
              else acc0 += A[c][u] * vqg[2 * u];
            }
          }
        }
        if (my_next) {
#pragma unroll
          for (int u = 0; u < UMAX; ++u)
            if (u == un) rowv = A[c][u];
        }
        const double w = acc0 + acc1;
        double* wdst = wpart + ((par * kTcMaxCl + rank) * NG + g) * n + j;
        double* rdst = rowbuf + par * n + j;
        if (CL == 1) {
          *wdst = w;
          if (my_next) *rdst = rowv;
        } else {
          for (int r = 0; r < CL; ++r) {
            *SPECGPU_MAP_SHARED(wdst, r) = w;
            if (my_next) *SPECGPU_MAP_SHARED(rdst, r) = rowv;
          }
        }
      }
    }
    if (CL > 1) SPECGPU_CLUSTER_SYNC();
    else __syncthreads();
    // d. p, K, q and the next row
    part = 0.0;
    for (int j = k + 1 + tid; j < n; j += T) {
      double p = 0.0;
      for (int r = 0; r < CL * NG; ++r) p += wpart[(par * kTcMaxCl * NG + r) * n + j];
      p *= beta;
      q[j] = p;
      part += v[j] * p;
    }
    const double K = 0.5 * beta * tc_block_sum<T>(part, red, tid);
    for (int j = k + 1 + tid; j < n; j += T) {
      const double qj = q[j] - K * v[j];
      q[j] = qj;
      if (j % CL == rank) {
        const int lr = j / CL;
        vq[((lr % NG) * UMAX + lr / NG) * 2 + 1] = qj;
      }
    }
    __syncthreads();
    const double v1 = v[k + 1], q1 = q[k + 1];
    part = 0.0;
    for (int j = k + 1 + tid; j < n; j += T) {
      const double xn = rowbuf[par * n + j] - v1 * q[j] - q1 * v[j];
      x[j] = xn;
      if (j > k + 1) part += xn * xn;                   // the next column's sigma
    }
    // e. rank-2 update of my rows
    if (refl) {
#pragma unroll
      for (int c = 0; c < NCOL; ++c) {
        const int j = cj + CW * c;
        if (j > k && j < n) {
          const double vj = v[j], qj = q[j];
#pragma unroll
          for (int u = 0; u < UMAX; ++u)
            if (u >= u0 && u < u1) A[c][u] -= vqg[2 * u] * qj + vqg[2 * u + 1] * vj;
        }
      }
    }
    // (no barrier here: the first one inside the next tc_block_sum orders x, v, q, vq against their next writers)
  }
  __syncthreads();
  // x holds row n - 2 from its diagonal on
  if (rank == 0 && tid == 0) {
    d[n - 2] = x[n - 2];
    e[n - 2] = x[n - 1];
    be[n - 2] = 0.0;
    e[n - 1] = 0.0;
    be[n - 1] = 0.0;
  }
  if (rank == (n - 1) % CL) {                           // d[n-1]: the last diagonal entry, wherever it lives
    const int lr = (n - 1) / CL;
    if (g == lr % NG && cj == (n - 1) % CW) {
      const int c_sel = (n - 1) / CW, u_sel = lr / NG;
      double val = 0.0;
#pragma unroll
      for (int c = 0; c < NCOL; ++c)
#pragma unroll
        for (int u = 0; u < UMAX; ++u)
          if (c == c_sel && u == u_sel) val = A[c][u];
      d[n - 1] = val;
    }
  }
}

// ======================================================================================================
// 2. All eigenvalues of the symmetric tridiagonal (d, e), descending: lam (float, the input of svd_plan_kernel) and lamd.
// The matrix is scaled by a power of two to norm <= 1.  Sturm count by the three-term recurrence on the leading
// principal minors, p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}: one dependent FMA per row instead of a division.  A
// minor smaller than 2^-100 of its predecessor is replaced by -2^-100 of it (LAPACK's pivmin rule in scaled units);
// every fourth row the pair (p_{i-1}, p_i) is renormalised if it has left [2^-300, 2^300].
// kBisK threads share an eigenvalue: each pass they count at the kBisK interior points that cut the bracket into
// kBisK + 1 equal parts and keep the part that holds the eigenvalue -- (kBisK + 1)^-passes = 2^-58 of the Gershgorin
// interval, i.e. below 2^-52 of the norm: the tridiagonalisation itself is no more accurate than that.
// ======================================================================================================
#ifdef SPECGPU_EMULATE
constexpr int kBisK = 2, kBisPasses = 37;
#else
constexpr int kBisK = 4, kBisPasses = 25;
#endif
// One CTA of kTriMaxN threads serves kTriMaxN / kBisK eigenvalues; a matrix is spread over kBisK CTAs (grid.y) so that 40
// matrices occupy every SM -- with one 1024-thread CTA per matrix the 25 x 256 dependent Sturm steps ran on 40 SMs'
// float64 units (0.63 ms).  Every CTA repeats the cheap Gershgorin / scaling prologue.
constexpr int kBisThreads = kTriMaxN;
constexpr int kBisEigs = kBisThreads / kBisK;

struct SturmState {
  double pa, pb;
  int cnt;
};

__device__ __forceinline__ void sturm_step(SturmState& st, double di, double e2, double x, double floor_rel) {
  double pc = (di - x) * st.pb - e2 * st.pa;
  const int eb = hi32(st.pb) & 0x7ff00000, ec = hi32(pc) & 0x7ff00000;
  if (ec + (100 << 20) < eb) pc = -floor_rel * st.pb;
  st.cnt += (int)((unsigned)(hi32(pc) ^ hi32(st.pb)) >> 31);
  st.pa = st.pb;
  st.pb = pc;
}

// number of eigenvalues of the scaled matrix below x
__device__ __forceinline__ int sturm_count(const double* sd, const double* se2, int n, double x) {
  const double floor_rel = pow2(-100);
  SturmState st;
  st.pa = 1.0;
  st.pb = sd[0] - x;
  if (fabs(st.pb) < floor_rel) st.pb = -floor_rel;
  st.cnt = st.pb < 0.0 ? 1 : 0;
  int i = 1;
  for (; i + 3 < n; i += 4) {
    sturm_step(st, sd[i], se2[i - 1], x, floor_rel);
    sturm_step(st, sd[i + 1], se2[i], x, floor_rel);
    sturm_step(st, sd[i + 2], se2[i + 1], x, floor_rel);
    sturm_step(st, sd[i + 3], se2[i + 2], x, floor_rel);
    const unsigned en = (unsigned)(hi32(st.pb) & 0x7ff00000);
    if (en - ((1023u - 300u) << 20) > (600u << 20)) {      // rare: renormalise p_i to [1, 2)
      const double sc = pow2(1023 - (int)(en >> 20));
      st.pa *= sc;
      st.pb *= sc;
    }
  }
  for (; i < n; ++i) sturm_step(st, sd[i], se2[i - 1], x, floor_rel);
  return st.cnt;
}

__global__ void __launch_bounds__(kBisThreads) bisect_kernel(const double* dall, const double* eall, int n, float* lam, double* lamd) {
  __shared__ double sd[kTriMaxN], se2[kTriMaxN], sred[2 * kBisThreads / 32];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  double lo = INFINITY, hi = -INFINITY, di = 0.0, er = 0.0;
  if (tid < n) {
    di = dall[b * n + tid];
    const double el = (tid > 0) ? fabs(eall[b * n + tid - 1]) : 0.0;
    er = (tid + 1 < n) ? fabs(eall[b * n + tid]) : 0.0;
    lo = di - el - er;                  // Gershgorin
    hi = di + el + er;
  }
  double mn = lo, mx = hi;
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((tid & 31) == 0) {
    sred[2 * (tid >> 5)] = mn;
    sred[2 * (tid >> 5) + 1] = mx;
  }
  __syncthreads();
  double gl = INFINITY, gu = -INFINITY;
  for (int w = 0; w < kBisThreads / 32; ++w) {
    gl = fmin(gl, sred[2 * w]);
    gu = fmax(gu, sred[2 * w + 1]);
  }
  const double tnorm = fmax(fabs(gl), fabs(gu));
  if (!(tnorm >= 2.3e-308) || !(tnorm < INFINITY)) {    // zero (or subnormal) matrix; non-finite input: NaN goes through
    if (tid < n) {
      const double ev = (tnorm < 2.3e-308) ? 0.0 : NAN;
      lamd[b * n + tid] = ev;
      lam[b * n + tid] = (float)ev;
    }
    return;
  }
  // power-of-two scale: tnorm * scale in [0.5, 1)
  const int ex = ((hi32(tnorm) >> 20) & 0x7ff) - 1022;
  const double scale = pow2(-ex), unscale = pow2(ex);
  if (tid < n) {
    sd[tid] = di * scale;
    const double es = er * scale;
    se2[tid] = es * es;                 // e[tid]^2 couples tid and tid + 1
  }
  __syncthreads();
  const int eig = blockIdx.y * kBisEigs + tid / kBisK, s = tid % kBisK;     // the kBisK threads of an eigenvalue sit in one warp
  const int lane0 = (tid & 31) - s;
  const double slack = 2.0 * 2.22e-16 * n + 2e-30;
  lo = gl * scale - slack;
  hi = gu * scale + slack;
  const int idx = n - 1 - (eig < n ? eig : n - 1);      // eigenvalue number idx (ascending, 0-based) = the eig-th largest
  for (int it = 0; it < kBisPasses; ++it) {
    const double w = hi - lo;
    const int cnt = sturm_count(sd, se2, n, lo + w * ((double)(s + 1) * (1.0 / (kBisK + 1))));
    double nlo = lo, nhi = hi;
    bool found = false;
#pragma unroll
    for (int t = 0; t < kBisK; ++t) {
      const int ct = __shfl_sync(0xffffffffu, cnt, lane0 + t);
      const double xt = lo + w * ((double)(t + 1) * (1.0 / (kBisK + 1)));
      if (!found) {
        if (ct > idx) {
          nhi = xt;
          found = true;
        } else {
          nlo = xt;
        }
      }
    }
    if (nlo <= nhi) {                   // (a non-monotone count pair cannot turn the bracket inside out)
      lo = nlo;
      hi = nhi;
    }
  }
  if (s == 0 && eig < n) {
    const double ev = 0.5 * (lo + hi) * unscale;
    lamd[b * n + eig] = ev;
    lam[b * n + eig] = (float)ev;
  }
}

// ======================================================================================================
// 4. Leading eigenvectors of the tridiagonal matrix by inverse iteration: block (b, l) -- one warp -- factors
// T - lambda_l I with partial pivoting (as LAPACK's dgttrf) and solves three times (dgttrs); lane 0 walks the
// recurrences in shared memory, the element-wise parts use the whole warp.
// The route is decided from plan[b] = {a, e, num_sing, status}: the vectors the projection will ask for are [a, e) or,
// when the complement is shorter, [0, a) u [e, n).  Only leading sets of at most kTriMaxVec vectors are served here.
// ======================================================================================================
struct TriVecArgs {
  const double* d;
  const double* e;
  const double* lamd;
  int n;
  int32_t* plan;
  double* Y;          // [B][kTriMaxVec][n] eigenvectors of T
  int32_t* nvec;      // [B] number of leading vectors computed (0 when flagged)
};

__device__ __forceinline__ int tri_needed_leading(int a, int e, int n, bool* ok) {
  if (e < a) e = a;
  const int nk = e - a;
  const bool complement = (n - nk) < nk;
  const int L = complement ? a : (nk > 0 ? e : 0);
  const int trailing = complement ? n - e : 0;
  *ok = trailing == 0 && L <= kTriMaxVec;
  return L;
}

__device__ __forceinline__ double warp_max_f64(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(32) trivec_kernel(TriVecArgs a) {
  __shared__ double sdl[kTriMaxN], sdd[kTriMaxN], sdu[kTriMaxN], sdu2[kTriMaxN], sx[kTriMaxN], sinv[kTriMaxN];
  __shared__ unsigned char spv[kTriMaxN];
  const int64_t b = blockIdx.x / kTriMaxVec;
  const int l = blockIdx.x % kTriMaxVec;
  const int n = a.n, lane = threadIdx.x;
  bool ok = false;
  const int L = tri_needed_leading(a.plan[b * 4 + 0], a.plan[b * 4 + 1], n, &ok);
  if (l == 0 && lane == 0) {
    a.plan[b * 4 + 3] = ok ? 0 : 1;   // not ok: the Jacobi solver redoes this matrix
    a.nvec[b] = ok ? L : 0;
  }
  if (!ok || l >= L) return;          // uniform over the warp
  const double* d = a.d + b * n;
  const double* e = a.e + b * n;
  const double lam = a.lamd[b * n + l];
  double tn = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double ei = (i + 1 < n) ? e[i] : 0.0, em = (i > 0) ? e[i - 1] : 0.0, di = d[i];
    tn = fmax(tn, fabs(di) + fabs(ei) + fabs(em));
    sdd[i] = di - lam;
    sdl[i] = ei;
    sdu[i] = ei;
    sdu2[i] = 0.0;
    // start vector with a little structure (a constant vector can be orthogonal to an eigenvector)
    sx[i] = 1.0 + 0.37 * (double)((i * 7 + l * 3) % 11) / 11.0;
  }
  const double tnorm = warp_max_f64(tn);
  const double tiny = fmax(tnorm * 2.22e-16, 1e-300);
  __syncwarp();
  if (lane == 0) {                                // dgttrf
    double dd = sdd[0];
    for (int i = 0; i + 1 < n; ++i) {
      const double dl = sdl[i];
      double dn = sdd[i + 1];
      if (fabs(dd) >= fabs(dl)) {
        if (dd == 0.0) dd = tiny;
        const double fact = dl / dd;
        sdd[i] = dd;
        sdl[i] = fact;
        dn -= fact * sdu[i];
        spv[i] = 0;
      } else {
        const double fact = dd / dl;
        sdd[i] = dl;
        sdl[i] = fact;
        const double temp = sdu[i];
        sdu[i] = dn;
        dn = temp - fact * dn;
        if (i + 2 < n) {
          sdu2[i] = sdu[i + 1];
          sdu[i + 1] = -fact * sdu[i + 1];
        }
        spv[i] = 1;
      }
      dd = dn;
    }
    sdd[n - 1] = dd;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    double dd = sdd[i];
    if (fabs(dd) < tiny) dd = (dd < 0.0) ? -tiny : tiny;
    sinv[i] = 1.0 / dd;
  }
  __syncwarp();
  bool bad = false;
  for (int it = 0; it < 3; ++it) {
    if (lane == 0) {
      double xi = sx[0];                          // dgttrs: L solve
      for (int i = 0; i + 1 < n; ++i) {
        const double xn = sx[i + 1];
        if (spv[i] == 0) {
          sx[i] = xi;
          xi = xn - sdl[i] * xi;
        } else {
          sx[i] = xn;
          xi = xi - sdl[i] * xn;
        }
      }
      double x1 = xi * sinv[n - 1], x2 = 0.0;     // U solve (du2[n-2] == 0)
      sx[n - 1] = x1;
      for (int i = n - 2; i >= 0; --i) {
        const double xv = (sx[i] - sdu[i] * x1 - sdu2[i] * x2) * sinv[i];
        sx[i] = xv;
        x2 = x1;
        x1 = xv;
      }
    }
    __syncwarp();
    double mx = 0.0;
    bool fin = true;
    for (int i = lane; i < n; i += 32) {
      const double ax = fabs(sx[i]);
      fin = fin && (ax < INFINITY);               // false for NaN as well
      mx = fmax(mx, ax);
    }
    mx = warp_max_f64(mx);
    fin = __all_sync(0xffffffffu, fin);
    if (!fin || !(mx > 0.0)) {
      bad = true;
      break;
    }
    const double inv = 1.0 / mx;
    for (int i = lane; i < n; i += 32) sx[i] *= inv;
    __syncwarp();
  }
  double nn = 0.0;
  for (int i = lane; i < n; i += 32) nn += sx[i] * sx[i];
  nn = warp_sum(nn);
  const double inv = (!bad && nn > 0.0) ? 1.0 / sqrt(nn) : NAN;     // NaN vector: the residual check flags the matrix
  double* y = a.Y + (b * kTriMaxVec + l) * (int64_t)n;
  for (int i = lane; i < n; i += 32) y[i] = bad ? NAN : sx[i] * inv;
}

// ======================================================================================================
// 5. z_k = Q y_k (reflectors applied in reverse), residual check against the original G, U[:, k] = z_k.
// One CTA of 16 warps per matrix.  Warp 0 first runs modified Gram-Schmidt over the L vectors in eigenvalue order (close
// eigenvalues give nearly parallel iterates).  Then rounds of 16 reflectors: all threads stage the rows from global
// memory (8 independent loads each), warp l applies them to z_l, which it keeps in registers.  The residual pass walks
// the rows of G with all warps, two rows in flight per warp.
// ======================================================================================================
struct TriBackArgs {
  const double* W;      // row k: v_k in W[k][k+1..n)
  const double* beta;
  const double* G;      // the original matrix
  const double* lamd;
  const double* Y;
  int n;
  const int32_t* nvec;
  int32_t* plan;
  float* U;             // [B][n][n]
};

constexpr size_t kTriBackSmem = (size_t)(kTriMaxVec + kTriChunk) * kTriMaxN * sizeof(double) + kTriChunk * sizeof(double);

__global__ void __launch_bounds__(kTriBackThreads) triback_kernel(TriBackArgs a) {
  SPECGPU_DYN_SMEM(smem_raw);
  double* sz = reinterpret_cast<double*>(smem_raw);         // [kTriMaxVec][kTriMaxN]
  double* sv = sz + kTriMaxVec * kTriMaxN;                   // [kTriChunk][kTriMaxN]
  double* sbe = sv + kTriChunk * kTriMaxN;                   // [kTriChunk]
  __shared__ int s_bad;
  const int64_t b = blockIdx.x;
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.nvec[b];
  if (L == 0) return;           // flagged, or nothing to compute
  if (tid == 0) s_bad = 0;
  const double* W = a.W + b * (int64_t)n * n;
  const double* G = a.G + b * (int64_t)n * n;
  const double* be = a.beta + b * n;
  if (warp < L) {
    const double* y = a.Y + (b * kTriMaxVec + warp) * (int64_t)n;
    for (int i = lane; i < kTriMaxN; i += 32) sz[warp * kTriMaxN + i] = (i < n) ? y[i] : 0.0;
  }
  __syncthreads();
  if (warp == 0) {
    int bad = 0;
    for (int l = 1; l < L; ++l) {
      double* yl = sz + l * kTriMaxN;
      for (int p = 0; p < l; ++p) {
        const double* yp = sz + p * kTriMaxN;
        double dot = 0.0;
        for (int i = lane; i < n; i += 32) dot += yp[i] * yl[i];
        dot = warp_sum(dot);
        for (int i = lane; i < n; i += 32) yl[i] -= dot * yp[i];
        __syncwarp();
      }
      double nn = 0.0;
      for (int i = lane; i < n; i += 32) nn += yl[i] * yl[i];
      nn = warp_sum(nn);
      if (!(nn > 1e-12)) bad = 1;                 // lost to its neighbours (or NaN): let the Jacobi solver do this matrix
      const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 0.0;
      for (int i = lane; i < n; i += 32) yl[i] *= inv;
      __syncwarp();
    }
    if (bad && lane == 0) s_bad = 1;
  }
  __syncthreads();
  double zr[kTriMaxN / 32];
#pragma unroll
  for (int u = 0; u < kTriMaxN / 32; ++u) zr[u] = (warp < L) ? sz[warp * kTriMaxN + lane + 32 * u] : 0.0;
  for (int khi = n - 3; khi >= 0; khi -= kTriChunk) {
    const int cnt = (khi + 1 < kTriChunk) ? khi + 1 : kTriChunk;      // reflectors khi, khi - 1, ..., khi - cnt + 1
    for (int idx = tid; idx < cnt * kTriMaxN; idx += kTriBackThreads) {
      const int r = idx / kTriMaxN, j = idx - r * kTriMaxN, k = khi - r;
      sv[idx] = (j > k && j < n) ? W[(int64_t)k * n + j] : 0.0;
    }
    if (tid < cnt) sbe[tid] = be[khi - tid];
    __syncthreads();
    if (warp < L) {
      for (int r = 0; r < cnt; ++r) {
        const double bk = sbe[r];
        if (bk == 0.0) continue;                  // uniform
        const double* vr = sv + r * kTriMaxN;
        double vreg[kTriMaxN / 32], dot = 0.0;
#pragma unroll
        for (int u = 0; u < kTriMaxN / 32; ++u) {
          vreg[u] = vr[lane + 32 * u];
          dot += vreg[u] * zr[u];
        }
        dot = warp_sum(dot) * bk;
#pragma unroll
        for (int u = 0; u < kTriMaxN / 32; ++u) zr[u] -= dot * vreg[u];
      }
    }
    __syncthreads();
  }
  if (warp < L) {
#pragma unroll
    for (int u = 0; u < kTriMaxN / 32; ++u) sz[warp * kTriMaxN + lane + 32 * u] = zr[u];
  }
  __syncthreads();
  // residual |G z - lambda z|_inf against the largest eigenvalue
  const double tol = 1e-9 * fmax(fabs(a.lamd[b * n]), 1e-300);
  int bad = 0;
  constexpr int NWB = kTriBackThreads / 32;
  for (int i = warp; i < n; i += 2 * NWB) {
    const bool two = i + NWB < n;
    const double* row0 = G + (int64_t)i * n;
    const double* row1 = two ? row0 + (int64_t)NWB * n : row0;
    double g0[kTriMaxN / 32], g1[kTriMaxN / 32];
#pragma unroll
    for (int u = 0; u < kTriMaxN / 32; ++u) {
      const int j = lane + 32 * u;
      g0[u] = (j < n) ? row0[j] : 0.0;
      g1[u] = (j < n) ? row1[j] : 0.0;
    }
    for (int l = 0; l < L; ++l) {
      const double* z = sz + l * kTriMaxN;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int u = 0; u < kTriMaxN / 32; ++u) {
        const double zj = z[lane + 32 * u];
        a0 += g0[u] * zj;
        a1 += g1[u] * zj;
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      const double lam = a.lamd[b * n + l];
      if (!(fabs(a0 - lam * z[i]) <= tol)) bad = 1;
      if (two && !(fabs(a1 - lam * z[i + NWB]) <= tol)) bad = 1;
    }
  }
  if (bad) s_bad = 1;
  float* Ub = a.U + b * (int64_t)n * n;
  for (int idx = tid; idx < n * L; idx += kTriBackThreads) {
    const int i = idx / L, l = idx - i * L;
    Ub[(int64_t)i * n + l] = (float)sz[l * kTriMaxN + i];
  }
  __syncthreads();
  if (tid == 0 && s_bad) a.plan[b * 4 + 3] = 1;
}

bool eig_tridiag_supported(int n) { return n >= 3 && n <= kTriMaxN; }

size_t eig_tridiag_workspace_bytes(int64_t B, int n) {
  // d, e, beta, lamd: 4 x [B][n] doubles; Y: [B][16][n] doubles; nvec: [B] ints
  return (size_t)B * n * 8 * 4 + (size_t)B * kTriMaxVec * n * 8 + (size_t)B * 4 + 1024;
}

static int launch_tridiag_cluster(const TriClArgs& a0, int64_t B, cudaStream_t stream) {
  TriClArgs a = a0;
  int cl = (a.n + kTcRowCap - 1) / kTcRowCap;     // the register tile holds kTcRowCap rows per CTA
  if (cl > kTcMaxCl) return -1;
  const size_t smem = tri_cluster_smem<kTcThreads>(a.n);
#ifdef SPECGPU_EMULATE
  if (a.n >= 8 && cl < 3) cl = 3;     // exercise the distributed path in the CPU tests as well
  a.cl = cl;
  SPECGPU_LAUNCH_CLUSTER(tridiag_cluster_kernel<kTcThreads>, (unsigned)(B * cl), kTcThreads, smem, stream, cl, a);
  return (int)cudaGetLastError();
#else
  {
    cudaError_t e = cudaFuncSetAttribute(tridiag_cluster_kernel<kTcThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  for (; cl <= kTcMaxCl; ++cl) {      // a cluster size the device refuses falls through to the next one
    a.cl = cl;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * cl));
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tridiag_cluster_kernel<kTcThreads>, a);
    if (e == cudaSuccess) return (int)cudaGetLastError();
    (void)cudaGetLastError();
    if (cl == kTcMaxCl) return (int)e;
  }
  return -1;
#endif
}

// Values: tridiagonalise G (kept intact; the reflectors go to W) and write all eigenvalues (descending) to lam / ws.
int launch_eig_tridiag_values(const double* G, double* W, int64_t B, int n, float* lam, void* ws, cudaStream_t stream) {
  if (B == 0) return 0;
  if (!eig_tridiag_supported(n)) return -1;
  double* d = static_cast<double*>(ws);
  double* e = d + B * n;
  double* beta = e + B * n;
  double* lamd = beta + B * n;
  TriClArgs ta{G, W, d, e, beta, n, 1};
  const int rc = launch_tridiag_cluster(ta, B, stream);
  if (rc != 0) return rc;
  SPECGPU_LAUNCH(bisect_kernel, dim3((unsigned)B, (unsigned)((n + kBisEigs - 1) / kBisEigs)), kBisThreads, 0, stream, (const double*)d, (const double*)e, n, lam, lamd);
  return (int)cudaGetLastError();
}

// Vectors: the leading vectors the plan asks for (plan[b][3] = 1 where this route does not apply or its checks fail).
int launch_eig_tridiag_vectors(const double* W, const double* G, int64_t B, int n, int32_t* plan, float* U, void* ws,
                               cudaStream_t stream) {
  if (B == 0) return 0;
  if (!eig_tridiag_supported(n)) return -1;
  double* d = static_cast<double*>(ws);
  double* e = d + B * n;
  double* beta = e + B * n;
  double* lamd = beta + B * n;
  double* Y = lamd + B * n;
  int32_t* nvec = reinterpret_cast<int32_t*>(Y + B * kTriMaxVec * n);
  TriVecArgs va{d, e, lamd, n, plan, Y, nvec};
  SPECGPU_LAUNCH(trivec_kernel, (unsigned)(B * kTriMaxVec), 32, 0, stream, va);
  TriBackArgs ba{W, beta, G, lamd, Y, n, nvec, plan, U};
#ifndef SPECGPU_EMULATE
  {
    cudaError_t e2 = cudaFuncSetAttribute(triback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTriBackSmem);
    if (e2 != cudaSuccess) return (int)e2;
  }
#endif
  SPECGPU_LAUNCH(triback_kernel, (unsigned)B, kTriBackThreads, kTriBackSmem, stream, ba);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
