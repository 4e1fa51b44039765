// Hand-written symmetric eigensolver for the per-chunk Gram matrices (n <= 1024, a batch of them at once):
// the k largest eigenpairs of every matrix, entirely on the device, no host round trips.
//
// It stands where the LAPACK work inside the reference's svdecon stands (facemap/utils.py:751-776: sklearn
// randomized_svd -> scipy lu / qr / gesdd; the Gram form is matlab/utils/svdecon.m:16-43, `eig`).  cuSOLVER's syevd is
// host-driven at these sizes — one launch sequence per column, ~19 us per column whatever the precision, the call
// returns only when the device is done (profiles/r02_eig.md) — so 15 chunk matrices of order 999 cost 70 ms
// batched.  Here one thread-block CLUSTER owns one matrix and walks its columns without leaving the device:
//
//   A  sytrd_cluster_kernel   Householder tridiagonalisation, lower variant (LAPACK dsytd2 conventions): per column
//                             the reflector, p = tau A v over the cluster's share of the trailing columns (partial
//                             sums in shared memory, reduced across the cluster through distributed shared memory),
//                             w = p - (tau/2)(p.v) v, rank-2 update; three cluster barriers per column
//   B  bisect_kernel          the k largest eigenvalues of the tridiagonal matrix by Sturm bisection
//   C  invit_kernel           their eigenvectors by inverse iteration on T - lambda I factored with partial
//                             pivoting (LAPACK dlagtf / dlagts), one thread per eigenvalue
//   D  backtransform_kernel   Q z for every eigenvector: one warp per vector (registers), reflectors staged in
//                             shared memory
//   E  check_kernel           self-check: adjacent eigenvectors orthogonal, everything finite
//
// tests/eigh_model.py restates the arithmetic in NumPy; tests/test_eigh_cluster_gpu.py holds the kernels to it and to
// numpy.linalg.eigh (eigenvalues 1e-13 of the norm, residuals 1e-12, orthogonality 1e-10 at n = 999, k = 250).
#include <cooperative_groups.h>
#include <math.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace fm {

constexpr int EC_CLUSTER = 8;
constexpr int EC_THREADS = 512;
constexpr int EC_WARPS = EC_THREADS / 32;
constexpr int EC_NMAX = 1024;

__device__ __forceinline__ double ec_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block (every thread gets it); `red` holds one double per warp
__device__ __forceinline__ double ec_block_sum(double v, double* red) {
  v = ec_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// ---- A: tridiagonalisation ---------------------------------------------------------------------------------------
// G: [batch][n][n] symmetric, fully populated (row-major == column-major); the lower triangle is worked on:
// element (r, c), r >= c, lives at G[c * n + r].  On return column c holds, below the sub-diagonal, the reflector
// v_c[c+2:] (v_c[c+1] = 1 implicit); d [n], e [n-1], tau [n-1] per matrix.
__global__ void __cluster_dims__(EC_CLUSTER, 1, 1) __launch_bounds__(EC_THREADS, 1)
    sytrd_cluster_kernel(double* __restrict__ Gall, int n, double* __restrict__ dall, double* __restrict__ eall,
                         double* __restrict__ tauall) {
  extern __shared__ double ec_smem[];
  double* v = ec_smem;                       // [EC_NMAX] current reflector, indexed by matrix row
  double* w = v + EC_NMAX;                   // [EC_NMAX] p, then w
  double* ppart = w + EC_NMAX;               // [EC_NMAX] this CTA's partial p
  double* accw = ppart + EC_NMAX;            // [EC_WARPS][EC_NMAX] per-warp partial p
  double* red = accw + EC_WARPS * EC_NMAX;   // [32]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int mat = blockIdx.x / EC_CLUSTER;
  double* A = Gall + (size_t)mat * n * n;
  double* dd = dall + (size_t)mat * n;
  double* ee = eall + (size_t)mat * n;
  double* tt = tauall + (size_t)mat * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int colbase = rank + EC_CLUSTER * warp;          // this warp's columns: colbase + 128 q
  const int colstep = EC_CLUSTER * EC_WARPS;
  double* myacc = accw + warp * EC_NMAX;

  for (int i = 0; i + 1 < n; ++i) {
    const int r0 = i + 1;                                // first row of the reflector
    // 1. column i below the diagonal -> v, its norm -> reflector (every CTA of the cluster, redundantly)
    double ss = 0.0;
    for (int r = r0 + tid; r < n; r += EC_THREADS) {
      const double x = __ldcg(A + (size_t)i * n + r);
      v[r] = x;
      if (r > r0) ss += x * x;
    }
    ss = ec_block_sum(ss, red);
    const double alpha = v[r0];
    double beta = alpha, tau = 0.0;
    if (ss > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      tau = (beta - alpha) / beta;
      const double scale = 1.0 / (alpha - beta);
      __syncthreads();
      for (int r = r0 + 1 + tid; r < n; r += EC_THREADS) v[r] *= scale;
    }
    __syncthreads();
    if (tid == 0) v[r0] = 1.0;
    if (rank == 0 && tid == 0) {
      ee[i] = beta;
      tt[i] = tau;
      dd[i] = __ldcg(A + (size_t)i * n + i);
    }
    if (tau == 0.0) continue;                            // identical on every CTA: same data, same arithmetic
    __syncthreads();
    // 2. partial p = A_trailing v over this CTA's columns (lower triangle, each element serves p_row and p_col)
    for (int r = r0 + lane; r < n; r += 32) myacc[r] = 0.0;
    __syncwarp();
    int q0 = (r0 - colbase + colstep - 1) / colstep;
    if (q0 < 0) q0 = 0;
    for (int j = colbase + colstep * q0; j < n; j += colstep) {
      const double vj = v[j];
      const double* col = A + (size_t)j * n;
      double dot = 0.0;
      for (int r = j + lane; r < n; r += 128) {
        double a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = (r + 32 * u < n) ? __ldcg(col + r + 32 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = r + 32 * u;
          if (rr < n) {
            dot += a[u] * v[rr];
            if (rr > j) myacc[rr] += a[u] * vj;
          }
        }
      }
      dot = ec_warp_sum(dot);
      __syncwarp();
      if (lane == 0) myacc[j] += dot;
      __syncwarp();
    }
    __syncthreads();
    for (int r = r0 + tid; r < n; r += EC_THREADS) {
      double s = 0.0;
#pragma unroll
      for (int wv = 0; wv < EC_WARPS; ++wv) s += accw[wv * EC_NMAX + r];
      ppart[r] = s;
    }
    cluster.sync();                                      // #1: every CTA's partial p is in its shared memory
    // 3. this CTA reduces its slice of p over the cluster and hands it to everybody (distributed shared memory)
    {
      const int m = n - r0;
      const int L = (m + EC_CLUSTER - 1) / EC_CLUSTER;
      const int lo = r0 + rank * L, hi = min(n, lo + L);
      for (int r = lo + tid; r < hi; r += EC_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < EC_CLUSTER; ++c) s += cluster.map_shared_rank(ppart, c)[r];
        s *= tau;
#pragma unroll
        for (int c = 0; c < EC_CLUSTER; ++c) cluster.map_shared_rank(w, c)[r] = s;
      }
      if (rank == 0)                                     // everybody has read column i: keep the reflector there
        for (int r = r0 + 1 + tid; r < n; r += EC_THREADS) __stcg(A + (size_t)i * n + r, v[r]);
    }
    cluster.sync();                                      // #2: p complete everywhere
    // 4. w = p - (tau / 2)(p . v) v
    double pv = 0.0;
    for (int r = r0 + tid; r < n; r += EC_THREADS) pv += w[r] * v[r];
    pv = ec_block_sum(pv, red);
    const double half = 0.5 * tau * pv;
    for (int r = r0 + tid; r < n; r += EC_THREADS) w[r] -= half * v[r];
    __syncthreads();
    // 5. rank-2 update of this CTA's columns: A -= v w^T + w v^T (lower triangle)
    for (int j = colbase + colstep * q0; j < n; j += colstep) {
      const double vj = v[j], wj = w[j];
      double* col = A + (size_t)j * n;
      for (int r = j + lane; r < n; r += 128) {
        double a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = (r + 32 * u < n) ? __ldcg(col + r + 32 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = r + 32 * u;
          if (rr < n) __stcg(col + rr, a[u] - (v[rr] * wj + w[rr] * vj));
        }
      }
    }
    cluster.sync();                                      // #3: the next pivot column is final before anyone reads it
  }
  if (rank == 0 && tid == 0) dd[n - 1] = __ldcg(A + (size_t)(n - 1) * n + (n - 1));
}

// ---- B: bisection ------------------------------------------------------------------------------------------------
// lam_desc[mat][j] = j-th largest eigenvalue of the tridiagonal (d, e); also tnorm[mat] = max(|d|, |e|)
__global__ void __launch_bounds__(128) bisect_kernel(const double* __restrict__ dall, const double* __restrict__ eall,
                                                     int n, int k, double* __restrict__ lamdesc,
                                                     double* __restrict__ tnorm_out) {
  extern __shared__ double bs_smem[];
  double* d = bs_smem;            // [n]
  double* e2 = d + n;             // [n]
  __shared__ double s_lo[4], s_hi[4], s_e2[4], s_tn[4];
  const int mat = blockIdx.y;
  const double* dg = dall + (size_t)mat * n;
  const double* eg = eall + (size_t)mat * n;
  double lo = 1e300, hi = -1e300, e2max = 0.0, tn = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double di = dg[i];
    const double el = i > 0 ? fabs(eg[i - 1]) : 0.0, er = i + 1 < n ? fabs(eg[i]) : 0.0;
    d[i] = di;
    e2[i] = er * er;
    lo = fmin(lo, di - el - er);
    hi = fmax(hi, di + el + er);
    e2max = fmax(e2max, er * er);
    tn = fmax(tn, fmax(fabs(di), er));
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    e2max = fmax(e2max, __shfl_xor_sync(0xffffffffu, e2max, o));
    tn = fmax(tn, __shfl_xor_sync(0xffffffffu, tn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; s_e2[threadIdx.x >> 5] = e2max; s_tn[threadIdx.x >> 5] = tn;
  }
  __syncthreads();
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    lo = fmin(lo, s_lo[i]); hi = fmax(hi, s_hi[i]); e2max = fmax(e2max, s_e2[i]); tn = fmax(tn, s_tn[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) tnorm_out[mat] = tn;
  const double tiny = 2.2250738585072014e-308, eps = 2.220446049250313e-16;
  const double pivmin = fmax(tiny * fmax(1.0, e2max), tiny);
  const double tnorm = fmax(fabs(lo), fabs(hi));
  const double pad = 2.0 * tnorm * eps * n + 2.0 * pivmin;
  lo -= pad;
  hi += pad;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int want = n - 1 - j;          // count(x) <= want  <=>  x <= lambda_j
  for (int it = 0; it < 90; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (!(mid > lo && mid < hi)) break;
    double q = d[0] - mid;
    if (fabs(q) < pivmin) q = -pivmin;
    int c = q < 0.0;
    for (int i = 1; i < n; ++i) {
      q = d[i] - mid - e2[i - 1] / q;
      if (fabs(q) < pivmin) q = -pivmin;
      c += q < 0.0;
    }
    if (c <= want) lo = mid; else hi = mid;
  }
  lamdesc[(size_t)mat * k + j] = 0.5 * (lo + hi);
}

// ---- C: inverse iteration ------------------------------------------------------------------------------------------
// one thread per eigenvalue; its factorisation lives in global scratch laid out [array][row][kpad] (coalesced over
// the eigenvalues of a warp).  Z[mat][j][n] receives the unit eigenvector of T for lam_desc[j].
__device__ __forceinline__ double ec_start_value(int i, int j) {
  unsigned long long h = ((unsigned long long)i * 0x9E3779B1ull + (unsigned long long)(j + 1) * 0x85EBCA77ull) & 0xFFFFFFFFull;
  h ^= h >> 15;
  h = (h * 0x2C1B3C6Dull) & 0xFFFFFFFFull;
  h ^= h >> 12;
  return ((double)h + 0.5) / 2147483648.0 - 1.0;
}

__global__ void __launch_bounds__(64) invit_kernel(const double* __restrict__ dall, const double* __restrict__ eall,
                                                   int n, int k, int kpad, const double* __restrict__ lamdesc,
                                                   const double* __restrict__ tnorm_in, double* __restrict__ fac,
                                                   unsigned char* __restrict__ swp, double* __restrict__ Z, int iters) {
  const int mat = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const double* d = dall + (size_t)mat * n;
  const double* e = eall + (size_t)mat * n;
  const double lam = lamdesc[(size_t)mat * k + j];
  const size_t plane = (size_t)n * kpad;
  double* Aa = fac + (size_t)mat * 5 * plane + j;        // element i at [i * kpad]
  double* Bb = Aa + plane;
  double* Cc = Bb + plane;
  double* D2 = Cc + plane;
  double* X = D2 + plane;
  unsigned char* SW = swp + (size_t)mat * plane + j;
  const double eps = 2.220446049250313e-16;
  const double tol = eps * fmax(tnorm_in[mat], 1e-300);
  // factor T - lam I = P L U (dlagtf)
  double cur_a = d[0] - lam, cur_b = n > 1 ? e[0] : 0.0;
  for (int i = 0; i + 1 < n; ++i) {
    const double ci = e[i];
    const double next_a = d[i + 1] - lam;
    const double next_b = i + 2 < n ? e[i + 1] : 0.0;
    const size_t o = (size_t)i * kpad;
    if (fabs(cur_a) >= fabs(ci)) {
      const double ai = cur_a == 0.0 ? tol : cur_a;
      const double m = ci / ai;
      Aa[o] = ai; Bb[o] = cur_b; Cc[o] = m; D2[o] = 0.0; SW[o] = 0;
      cur_a = next_a - m * cur_b;
      cur_b = next_b;
    } else {
      const double m = cur_a / ci;
      Aa[o] = ci; Bb[o] = next_a; Cc[o] = m; D2[o] = next_b; SW[o] = 1;
      cur_a = cur_b - m * next_a;
      cur_b = -m * next_b;
    }
  }
  {
    const size_t o = (size_t)(n - 1) * kpad;
    Aa[o] = cur_a; Bb[o] = 0.0; D2[o] = 0.0;
  }
  for (int i = 0; i < n; ++i) X[(size_t)i * kpad] = ec_start_value(i, j);
  double scale = 1.0, ss = 0.0;
  for (int it = 0; it < iters; ++it) {
    // forward: the row operations applied to the right-hand side (scaled by the previous maximum)
    double carry = X[0] * scale;
    for (int i = 0; i + 1 < n; ++i) {
      const size_t o = (size_t)i * kpad;
      const double xn = X[o + kpad] * scale;
      const double m = Cc[o];
      if (SW[o]) {
        X[o] = xn;
        carry = carry - m * xn;
      } else {
        X[o] = carry;
        carry = xn - m * carry;
      }
    }
    X[(size_t)(n - 1) * kpad] = carry;
    // backward substitution with U = (A, B, D2), tiny pivots replaced
    double x1 = 0.0, x2 = 0.0, xmax = 0.0;
    ss = 0.0;
    for (int i = n - 1; i >= 0; --i) {
      const size_t o = (size_t)i * kpad;
      double t = X[o] - Bb[o] * x1 - D2[o] * x2;
      double piv = Aa[o];
      if (fabs(piv) < tol) piv = piv < 0.0 ? -tol : tol;
      const double xi = t / piv;
      X[o] = xi;
      xmax = fmax(xmax, fabs(xi));
      ss += xi * xi;
      x2 = x1;
      x1 = xi;
    }
    scale = xmax > 0.0 ? 1.0 / xmax : 1.0;
  }
  // unit norm: ss was accumulated on the unscaled last iterate
  double rn = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
  if (!isfinite(rn)) {        // overflow of the sum of squares: go through the max-scaled vector
    double s2 = 0.0;
    for (int i = 0; i < n; ++i) { const double xi = X[(size_t)i * kpad] * scale; s2 += xi * xi; }
    rn = scale / sqrt(s2);
  }
  double* z = Z + ((size_t)mat * k + j) * n;
  for (int i = 0; i < n; ++i) z[i] = X[(size_t)i * kpad] * rn;
}

// ---- D: back-transformation ------------------------------------------------------------------------------------------
// W[mat][k-1-j][:] = H_0 H_1 ... H_{n-2} z_j  (rows in ASCENDING order of eigenvalue, like syevd's output)
constexpr int BT_WARPS = 8;
__global__ void __launch_bounds__(BT_WARPS * 32) backtransform_kernel(const double* __restrict__ Gall, int n, int k,
                                                                      const double* __restrict__ tauall,
                                                                      const double* __restrict__ Z,
                                                                      const double* __restrict__ lamdesc,
                                                                      double* __restrict__ W, double* __restrict__ lam_asc) {
  __shared__ double vbuf[2][EC_NMAX];
  const int mat = blockIdx.y;
  const double* A = Gall + (size_t)mat * n * n;
  const double* tau = tauall + (size_t)mat * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * BT_WARPS + warp;
  const bool live = j < k;
  double z[EC_NMAX / 32];
#pragma unroll
  for (int t = 0; t < EC_NMAX / 32; ++t) {
    const int r = lane + 32 * t;
    z[t] = (live && r < n) ? Z[((size_t)mat * k + j) * n + r] : 0.0;
  }
  int buf = 0;
  for (int i = n - 2; i >= 0; --i) {
    const double ti = tau[i];
    if (ti == 0.0) continue;                             // uniform over the block
    const int r0 = i + 1;
    for (int r = r0 + 1 + threadIdx.x; r < n; r += BT_WARPS * 32) vbuf[buf][r] = __ldcg(A + (size_t)i * n + r);
    if (threadIdx.x == 0) vbuf[buf][r0] = 1.0;
    __syncthreads();
    const double* v = vbuf[buf];
    double s = 0.0;
    const int t0 = r0 >> 5;
#pragma unroll
    for (int t = 0; t < EC_NMAX / 32; ++t) {
      if (t < t0) continue;
      const int r = lane + 32 * t;
      if (r >= r0 && r < n) s += v[r] * z[t];
    }
    s = ec_warp_sum(s) * ti;
#pragma unroll
    for (int t = 0; t < EC_NMAX / 32; ++t) {
      if (t < t0) continue;
      const int r = lane + 32 * t;
      if (r >= r0 && r < n) z[t] -= s * v[r];
    }
    buf ^= 1;
  }
  if (!live) return;
  double* out = W + ((size_t)mat * k + (k - 1 - j)) * n;
#pragma unroll
  for (int t = 0; t < EC_NMAX / 32; ++t) {
    const int r = lane + 32 * t;
    if (r < n) out[r] = z[t];
  }
  if (lane == 0) lam_asc[(size_t)mat * k + (k - 1 - j)] = lamdesc[(size_t)mat * k + j];
}

// ---- E: self-check -----------------------------------------------------------------------------------------------------
// status[mat] |= 1 when two neighbouring eigenvectors of non-negligible eigenvalues are not orthogonal (a cluster the
// independent inverse iterations could not separate), |= 2 when something is not finite or not of unit length
__global__ void __launch_bounds__(256) check_kernel(const double* __restrict__ W, const double* __restrict__ lam_asc,
                                                    int n, int k, int* __restrict__ status) {
  const int mat = blockIdx.x;
  const double* Wm = W + (size_t)mat * k * n;
  const double* lam = lam_asc + (size_t)mat * k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const double top = fabs(lam[k - 1]);
  int bad = 0;
  for (int r = warp; r < k; r += nw) {
    const double* a = Wm + (size_t)r * n;
    const double* b = r + 1 < k ? a + n : a;
    double dot = 0.0, nrm = 0.0;
    for (int i = lane; i < n; i += 32) {
      dot += a[i] * b[i];
      nrm += a[i] * a[i];
    }
    dot = ec_warp_sum(dot);
    nrm = ec_warp_sum(nrm);
    if (!isfinite(nrm) || fabs(nrm - 1.0) > 1e-6 || !isfinite(lam[r])) bad |= 2;
    if (r + 1 < k && fabs(lam[r]) > 1e-10 * top && fabs(dot) > 1e-6) bad |= 1;
  }
  if (bad && lane == 0) atomicOr(status + mat, bad);
}

}  // namespace fm

extern "C" size_t fm_eigh_topk_scratch_bytes(int n, int batch, int k) {
  const size_t kpad = (size_t)((k + 63) / 64) * 64;
  size_t doubles = (size_t)batch * ((size_t)3 * n + k + 1 + (size_t)k * n + (size_t)5 * n * kpad);
  return doubles * sizeof(double) + (size_t)batch * n * kpad + 256;
}

// The k largest eigenpairs of `batch` symmetric n x n float64 matrices stored back to back (n <= 1024).  G is
// destroyed (it holds the Householder reflectors afterwards).  lam [batch, k] and W [batch, k, n] come back in
// ASCENDING order of eigenvalue (row k-1 = the largest), the layout fm_topk_components reads with nev = k.
// status [batch] (int32, zeroed by this call) receives the self-check flags; the caller reads it when it synchronises
// anyway.  Everything is asynchronous on `stream`.
extern "C" int fm_eigh_topk_batched(double* G, int n, int batch, int k, double* lam, double* W, void* scratch,
                                    size_t scratch_bytes, int* status, void* stream) {
  using namespace fm;
  FM_REQUIRE(G && lam && W && scratch && status && batch >= 1, "fm_eigh_topk_batched: bad args");
  FM_REQUIRE(n >= 2 && n <= EC_NMAX && k >= 1 && k <= n, "fm_eigh_topk_batched: needs 2 <= n <= %d and 1 <= k <= n (n = %d, k = %d)",
             EC_NMAX, n, k);
  FM_REQUIRE(scratch_bytes >= fm_eigh_topk_scratch_bytes(n, batch, k), "fm_eigh_topk_batched: scratch too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t kpad = (size_t)((k + 63) / 64) * 64;
  double* d = static_cast<double*>(scratch);
  double* e = d + (size_t)batch * n;
  double* tau = e + (size_t)batch * n;
  double* lamdesc = tau + (size_t)batch * n;
  double* tnorm = lamdesc + (size_t)batch * k;
  double* Z = tnorm + batch;
  double* fac = Z + (size_t)batch * k * n;
  unsigned char* swp = reinterpret_cast<unsigned char*>(fac + (size_t)batch * 5 * n * kpad);
  FM_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)batch, st));
  {
    const size_t smem = sizeof(double) * ((size_t)(3 + EC_WARPS) * EC_NMAX + 32);
    static bool configured = false;
    if (!configured) {
      FM_CUDA_TRY(cudaFuncSetAttribute(sytrd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    sytrd_cluster_kernel<<<dim3((unsigned)(batch * EC_CLUSTER)), EC_THREADS, smem, st>>>(G, n, d, e, tau);
    FM_LAUNCH_CHECK();
  }
  {
    const size_t smem = sizeof(double) * 2 * (size_t)n;
    bisect_kernel<<<dim3((unsigned)cdiv(k, 128), (unsigned)batch), 128, smem, st>>>(d, e, n, k, lamdesc, tnorm);
    FM_LAUNCH_CHECK();
  }
  invit_kernel<<<dim3((unsigned)cdiv(k, 64), (unsigned)batch), 64, 0, st>>>(d, e, n, k, (int)kpad, lamdesc, tnorm, fac, swp, Z, 3);
  FM_LAUNCH_CHECK();
  backtransform_kernel<<<dim3((unsigned)cdiv(k, BT_WARPS), (unsigned)batch), BT_WARPS * 32, 0, st>>>(G, n, k, tau, Z, lamdesc, W, lam);
  FM_LAUNCH_CHECK();
  check_kernel<<<(unsigned)batch, 256, 0, st>>>(W, lam, n, k, status);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
