// Values-first symmetric eigensolver for the `use_optimal` / `computeSignal` modes of denoiseSignal
// (spec_denoising/denoising_by_svd.ipynb:161-186, 210-217): those modes need ALL singular values (np.median, the count
// above the Gavish-Donoho threshold) but only the few leading vectors.  The cluster Jacobi solver (svd.cu) delivers all
// 256 vectors in ~20 sweeps and needs 4 SMs per matrix (two waves for 40 matrices: 16 ms); here, in float64,
//   1. tridiag_kernel   Householder tridiagonalisation of a work copy of G, one CTA per matrix, G streamed from L2
//   2. bisect_kernel    all eigenvalues of the tridiagonal matrix by Sturm-count bisection, one thread per eigenvalue
//   3. (svd_plan_kernel, svd.cu: median, threshold, num_sing, start/stop)
//   4. trivec_kernel    the L <= 16 leading eigenvectors by inverse iteration on the tridiagonal matrix (pivoted LU,
//                       one thread per vector), modified Gram-Schmidt among them
//   5. triback_kernel   back-transformation through the Householder reflectors, residual check |G z - lambda z| against
//                       the ORIGINAL G, U[:, k] written; a matrix whose plan needs other vectors (trailing ones, more than
//                       16) or whose check fails is flagged (plan[b][3] = 1) and redone by the Jacobi solver, which skips
//                       the others.
// Plain C++ under SPECGPU_EMULATE as well (no tensor-core / TMA instructions here).
#include "kernels.h"

namespace specgpu {

constexpr int kTriThreads = 1024;
constexpr int kTriMaxVec = 16;
constexpr int kTriBackThreads = 32 * kTriMaxVec;
constexpr int kTriMaxN = 256;

__device__ __forceinline__ double tri_block_sum(double v, double* red, int tid) {
  v = warp_sum(v);
  __syncthreads();                       // red may still be read from the previous reduction
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < kTriThreads / 32; ++w) s += red[w];
  return s;
}

// A: [B][n][n] work copy of the symmetric matrix (destroyed: row k keeps the Householder vector v_k in A[k][k+1..n)),
// d[B][n], e[B][n] (e[k] couples k and k+1), beta[B][n].
__global__ void __launch_bounds__(kTriThreads) tridiag_kernel(double* Aall, int n, double* dall, double* eall, double* ball) {
  __shared__ double sv[kTriMaxN], sp[kTriMaxN], red[kTriThreads / 32];
  const int64_t b = blockIdx.x;
  double* A = Aall + b * (int64_t)n * n;
  double* d = dall + b * n;
  double* e = eall + b * n;
  double* be = ball + b * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kTriThreads / 32;
  for (int k = 0; k + 2 < n; ++k) {
    const int m = n - k - 1;
    double* x = A + (int64_t)k * n + k + 1;          // row k right of the diagonal (== column k below it)
    double part = 0.0;
    for (int i = tid; i < m; i += kTriThreads) part += x[i] * x[i];
    const double x0 = x[0];                           // (read before the barriers of the sum: x is rewritten below)
    const double sigma = tri_block_sum(part, red, tid);
    if (tid == 0) d[k] = A[(int64_t)k * n + k];
    const double tail = sigma - x0 * x0;              // already tridiagonal in this column?
    if (!(tail > 0.0)) {                              // uniform
      if (tid == 0) {
        e[k] = x0;
        be[k] = 0.0;
      }
      __syncthreads();
      for (int i = tid; i < m; i += kTriThreads) x[i] = 0.0;      // v_k = 0: the reflector is the identity
      __syncthreads();
      continue;
    }
    const double alpha = (x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma);
    const double beta = 1.0 / (sigma - alpha * x0);   // 2 / (v^T v) with v = x - alpha e_0
    for (int i = tid; i < m; i += kTriThreads) {
      const double vi = (i == 0) ? x0 - alpha : x[i];
      sv[i] = vi;
      x[i] = vi;                                      // keep v_k for the back-transformation
    }
    if (tid == 0) {
      e[k] = alpha;
      be[k] = beta;
    }
    __syncthreads();
    // p = beta * A22 v.  A warp takes rows i, i + NW, ... two at a time with all their loads in flight before the first
    // FMA (the matrix is streamed from L2 by one SM: latency, not bandwidth, is what has to be hidden).
    for (int i = warp; i < m; i += 2 * NW) {
      const double* row0 = A + (int64_t)(k + 1 + i) * n + k + 1;
      const bool two = i + NW < m;
      const double* row1 = two ? row0 + (int64_t)NW * n : row0;
      double r0[kTriMaxN / 32], r1[kTriMaxN / 32];
#pragma unroll
      for (int u = 0; u < kTriMaxN / 32; ++u) {
        const int j = lane + 32 * u;
        r0[u] = (j < m) ? row0[j] : 0.0;
        r1[u] = (j < m) ? row1[j] : 0.0;
      }
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int u = 0; u < kTriMaxN / 32; ++u) {
        const int j = lane + 32 * u;
        const double vj = (j < m) ? sv[j] : 0.0;
        acc0 += r0[u] * vj;
        acc1 += r1[u] * vj;
      }
      acc0 = warp_sum(acc0);
      acc1 = warp_sum(acc1);
      if (lane == 0) {
        sp[i] = beta * acc0;
        if (two) sp[i + NW] = beta * acc1;
      }
    }
    __syncthreads();
    part = 0.0;
    for (int i = tid; i < m; i += kTriThreads) part += sv[i] * sp[i];
    const double K = 0.5 * beta * tri_block_sum(part, red, tid);
    for (int i = tid; i < m; i += kTriThreads) sp[i] -= K * sv[i];      // q
    __syncthreads();
    // A22 -= v q^T + q v^T  (same two-rows-at-a-time walk)
    for (int i = warp; i < m; i += 2 * NW) {
      double* row0 = A + (int64_t)(k + 1 + i) * n + k + 1;
      const bool two = i + NW < m;
      double* row1 = two ? row0 + (int64_t)NW * n : row0;
      const double v0 = sv[i], q0 = sp[i];
      const double v1 = two ? sv[i + NW] : 0.0, q1 = two ? sp[i + NW] : 0.0;
      double r0[kTriMaxN / 32], r1[kTriMaxN / 32];
#pragma unroll
      for (int u = 0; u < kTriMaxN / 32; ++u) {
        const int j = lane + 32 * u;
        r0[u] = (j < m) ? row0[j] : 0.0;
        r1[u] = (j < m && two) ? row1[j] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kTriMaxN / 32; ++u) {
        const int j = lane + 32 * u;
        if (j < m) {
          const double vj = sv[j], qj = sp[j];
          row0[j] = r0[u] - (v0 * qj + q0 * vj);
          if (two) row1[j] = r1[u] - (v1 * qj + q1 * vj);
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (n >= 2) {
      d[n - 2] = A[(int64_t)(n - 2) * n + n - 2];
      e[n - 2] = A[(int64_t)(n - 2) * n + n - 1];
      be[n - 2] = 0.0;
    }
    d[n - 1] = A[(int64_t)(n - 1) * n + n - 1];
    e[n - 1] = 0.0;
    be[n - 1] = 0.0;
  }
}

// All eigenvalues of the symmetric tridiagonal (d, e), descending: lam (float, the input of svd_plan_kernel) and lamd.
__global__ void __launch_bounds__(kTriMaxN) bisect_kernel(const double* dall, const double* eall, int n, float* lam, double* lamd) {
  __shared__ double sd[kTriMaxN], se2[kTriMaxN], sred[2 * kTriMaxN / 32];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  double lo = 0.0, hi = 0.0, emax = 0.0;
  if (tid < n) {
    const double di = dall[b * n + tid];
    const double el = (tid > 0) ? fabs(eall[b * n + tid - 1]) : 0.0;
    const double er = (tid + 1 < n) ? fabs(eall[b * n + tid]) : 0.0;
    sd[tid] = di;
    se2[tid] = er * er;                 // e[tid]^2 couples tid and tid + 1
    lo = di - el - er;                  // Gershgorin
    hi = di + el + er;
    emax = er;
  } else {
    lo = INFINITY;
    hi = -INFINITY;
  }
  // block-wide min / max
  double mn = lo, mx = hi, em = emax;
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    em = fmax(em, __shfl_xor_sync(0xffffffffu, em, o));
  }
  if ((tid & 31) == 0) {
    sred[2 * (tid >> 5)] = mn;
    sred[2 * (tid >> 5) + 1] = mx;
  }
  __syncthreads();
  double gl = INFINITY, gu = -INFINITY;
  for (int w = 0; w < kTriMaxN / 32; ++w) {
    gl = fmin(gl, sred[2 * w]);
    gu = fmax(gu, sred[2 * w + 1]);
  }
  __syncthreads();
  if ((tid & 31) == 0) sred[tid >> 5] = em;
  __syncthreads();
  double e2max = 0.0;
  for (int w = 0; w < kTriMaxN / 32; ++w) e2max = fmax(e2max, sred[w] * sred[w]);
  const double tnorm = fmax(fabs(gl), fabs(gu));
  const double pivmin = fmax(2.2250738585072014e-308 * fmax(e2max, 1.0), 1e-300);
  gl -= 2.0 * tnorm * 2.22e-16 * n + 2.0 * pivmin;
  gu += 2.0 * tnorm * 2.22e-16 * n + 2.0 * pivmin;
  if (tid >= n) return;
  // eigenvalue number idx (ascending, 0-based) = the (tid)-th largest
  const int idx = n - 1 - tid;
  lo = gl;
  hi = gu;
  for (int it = 0; it < 80; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;      // interval exhausted in floating point
    // Sturm count: number of eigenvalues < mid
    double qv = sd[0] - mid;
    if (fabs(qv) < pivmin) qv = -pivmin;
    int cnt = qv < 0.0 ? 1 : 0;
    for (int i = 1; i < n; ++i) {
      qv = sd[i] - mid - se2[i - 1] / qv;
      if (fabs(qv) < pivmin) qv = -pivmin;
      cnt += qv < 0.0 ? 1 : 0;
    }
    if (cnt > idx) hi = mid;
    else lo = mid;
  }
  const double ev = 0.5 * (lo + hi);
  lamd[b * n + tid] = ev;
  lam[b * n + tid] = (float)ev;
}

// Leading eigenvectors of the tridiagonal matrix: lane l < L does inverse iteration for eigenvalue l (pivoted LU of
// T - lambda I as in LAPACK's dgttrf/dgttrs), then modified Gram-Schmidt in order.  One warp per matrix.
// Decides the route: plan[b] = {a, e, num_sing, status}; the vectors the projection will ask for are [a, e) or, when the
// complement is shorter, [0, a) u [e, n).  Only leading sets of at most kTriMaxVec vectors are served here.
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

__global__ void __launch_bounds__(32) trivec_kernel(TriVecArgs a) {
  const int64_t b = blockIdx.x;
  const int n = a.n, lane = threadIdx.x;
  bool ok = false;
  const int L = tri_needed_leading(a.plan[b * 4 + 0], a.plan[b * 4 + 1], n, &ok);
  if (!ok) {                   // uniform over the warp
    if (lane == 0) {
      a.plan[b * 4 + 3] = 1;   // the Jacobi solver redoes this matrix
      a.nvec[b] = 0;
    }
    return;
  }
  const double* d = a.d + b * n;
  const double* e = a.e + b * n;
  double* Y = a.Y + b * (int64_t)kTriMaxVec * n;
  int bad = 0;
  if (lane < L) {
    double dl[kTriMaxN], dd[kTriMaxN], du[kTriMaxN], du2[kTriMaxN], x[kTriMaxN];
    unsigned char pv[kTriMaxN];
    const double lam = a.lamd[b * n + lane];
    double tnorm = 0.0;
    for (int i = 0; i < n; ++i) tnorm = fmax(tnorm, fabs(d[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0));
    const double tiny = fmax(tnorm * 2.22e-16, 1e-300);
    for (int i = 0; i < n; ++i) {
      dd[i] = d[i] - lam;
      dl[i] = (i + 1 < n) ? e[i] : 0.0;
      du[i] = (i + 1 < n) ? e[i] : 0.0;
      du2[i] = 0.0;
    }
    for (int i = 0; i + 1 < n; ++i) {            // dgttrf
      if (fabs(dd[i]) >= fabs(dl[i])) {
        if (dd[i] == 0.0) dd[i] = tiny;
        const double fact = dl[i] / dd[i];
        dl[i] = fact;
        dd[i + 1] -= fact * du[i];
        pv[i] = 0;
      } else {
        const double fact = dd[i] / dl[i];
        dd[i] = dl[i];
        dl[i] = fact;
        const double temp = du[i];
        du[i] = dd[i + 1];
        dd[i + 1] = temp - fact * dd[i + 1];
        if (i + 2 < n) {
          du2[i] = du[i + 1];
          du[i + 1] = -fact * du[i + 1];
        }
        pv[i] = 1;
      }
    }
    if (dd[n - 1] == 0.0) dd[n - 1] = tiny;
    if (fabs(dd[n - 1]) < tiny) dd[n - 1] = (dd[n - 1] < 0.0) ? -tiny : tiny;
    // start vector with a little structure (a constant vector can be orthogonal to an eigenvector)
    for (int i = 0; i < n; ++i) x[i] = 1.0 + 0.37 * (double)((i * 7 + lane * 3) % 11) / 11.0;
    for (int it = 0; it < 3; ++it) {
      for (int i = 0; i + 1 < n; ++i) {          // dgttrs: L solve
        if (pv[i] == 0) {
          x[i + 1] -= dl[i] * x[i];
        } else {
          const double temp = x[i];
          x[i] = x[i + 1];
          x[i + 1] = temp - dl[i] * x[i];
        }
      }
      x[n - 1] /= dd[n - 1];                     // U solve
      if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / (fabs(dd[n - 2]) < tiny ? tiny : dd[n - 2]);
      for (int i = n - 3; i >= 0; --i)
        x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / (fabs(dd[i]) < tiny ? tiny : dd[i]);
      double mx = 0.0;
      for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(x[i]));
      if (!(mx > 0.0) || !(mx < INFINITY)) {
        bad = 1;
        break;
      }
      const double inv = 1.0 / mx;
      for (int i = 0; i < n; ++i) x[i] *= inv;
    }
    double nn = 0.0;
    for (int i = 0; i < n; ++i) nn += x[i] * x[i];
    const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 0.0;
    for (int i = 0; i < n; ++i) Y[(int64_t)lane * n + i] = x[i] * inv;
  }
  __syncwarp();
  // modified Gram-Schmidt in eigenvalue order (close eigenvalues give nearly parallel iterates)
  for (int l = 1; l < L; ++l) {
    double* yl = Y + (int64_t)l * n;
    for (int p = 0; p < l; ++p) {
      const double* yp = Y + (int64_t)p * n;
      double dot = 0.0;
      for (int i = lane; i < n; i += 32) dot += yp[i] * yl[i];
      dot = warp_sum(dot);
      for (int i = lane; i < n; i += 32) yl[i] -= dot * yp[i];
      __syncwarp();
    }
    double nn = 0.0;
    for (int i = lane; i < n; i += 32) nn += yl[i] * yl[i];
    nn = warp_sum(nn);
    if (!(nn > 1e-12)) bad = 1;                   // lost to its neighbours: let the Jacobi solver do this matrix
    const double inv = (nn > 0.0) ? 1.0 / sqrt(nn) : 0.0;
    for (int i = lane; i < n; i += 32) yl[i] *= inv;
    __syncwarp();
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    a.plan[b * 4 + 3] = bad ? 1 : 0;
    a.nvec[b] = bad ? 0 : L;
  }
}

// z_k = Q y_k (reflectors applied in reverse), residual check against the original G, U[:, k] = z_k.
struct TriBackArgs {
  const double* W;      // the work copy after tridiag_kernel (row k: v_k)
  const double* beta;
  const double* G;      // the original matrix
  const double* lamd;
  const double* Y;
  int n;
  const int32_t* nvec;
  int32_t* plan;
  float* U;             // [B][n][n]
};

__global__ void __launch_bounds__(kTriBackThreads) triback_kernel(TriBackArgs a) {
  __shared__ double sz[kTriMaxVec][kTriMaxN];
  __shared__ int s_bad;
  const int64_t b = blockIdx.x;
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.nvec[b];
  if (L == 0) return;           // flagged, or nothing to compute
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const double* W = a.W + b * (int64_t)n * n;
  const double* G = a.G + b * (int64_t)n * n;
  const double* be = a.beta + b * n;
  if (warp < L) {
    double* z = sz[warp];
    const double* y = a.Y + (b * kTriMaxVec + warp) * (int64_t)n;
    for (int i = lane; i < n; i += 32) z[i] = y[i];
    __syncwarp();
    for (int k = n - 3; k >= 0; --k) {
      const double bk = be[k];
      if (bk == 0.0) continue;                    // uniform
      const int m = n - k - 1;
      const double* v = W + (int64_t)k * n + k + 1;
      double dot = 0.0;
      for (int i = lane; i < m; i += 32) dot += v[i] * z[k + 1 + i];
      dot = warp_sum(dot) * bk;
      for (int i = lane; i < m; i += 32) z[k + 1 + i] -= dot * v[i];
      __syncwarp();
    }
    // residual |G z - lambda z|_inf against the largest eigenvalue
    const double lam = a.lamd[b * n + warp], lam0 = fabs(a.lamd[b * n]);
    double rmax = 0.0;
    for (int i = 0; i < n; ++i) {
      const double* row = G + (int64_t)i * n;
      double acc = 0.0;
      for (int j = lane; j < n; j += 32) acc += row[j] * z[j];
      acc = warp_sum(acc);
      rmax = fmax(rmax, fabs(acc - lam * z[i]));
    }
    if (!(rmax <= 1e-9 * fmax(lam0, 1e-300))) s_bad = 1;
    float* Ub = a.U + b * (int64_t)n * n;
    for (int i = lane; i < n; i += 32) Ub[(int64_t)i * n + warp] = (float)z[i];
  }
  __syncthreads();
  if (tid == 0 && s_bad) a.plan[b * 4 + 3] = 1;
}

bool eig_tridiag_supported(int n) { return n >= 3 && n <= kTriMaxN; }

size_t eig_tridiag_workspace_bytes(int64_t B, int n) {
  // d, e, beta, lamd: 4 x [B][n] doubles; Y: [B][16][n] doubles; nvec: [B] ints
  return (size_t)B * n * 8 * 4 + (size_t)B * kTriMaxVec * n * 8 + (size_t)B * 4 + 1024;
}

// Values: tridiagonalise the work copy W (destroyed) and write all eigenvalues (descending) to lam (float) / ws.
int launch_eig_tridiag_values(double* W, int64_t B, int n, float* lam, void* ws, cudaStream_t stream) {
  if (B == 0) return 0;
  if (!eig_tridiag_supported(n)) return -1;
  double* d = static_cast<double*>(ws);
  double* e = d + B * n;
  double* beta = e + B * n;
  double* lamd = beta + B * n;
  SPECGPU_LAUNCH(tridiag_kernel, (unsigned)B, kTriThreads, 0, stream, W, n, d, e, beta);
  SPECGPU_LAUNCH(bisect_kernel, (unsigned)B, kTriMaxN, 0, stream, (const double*)d, (const double*)e, n, lam, lamd);
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
  SPECGPU_LAUNCH(trivec_kernel, (unsigned)B, 32, 0, stream, va);
  TriBackArgs ba{W, beta, G, lamd, Y, n, nvec, plan, U};
  SPECGPU_LAUNCH(triback_kernel, (unsigned)B, kTriBackThreads, 0, stream, ba);
  return (int)cudaGetLastError();
}

}  // namespace specgpu
