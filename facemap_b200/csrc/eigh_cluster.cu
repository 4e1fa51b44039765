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
#include <stdlib.h>
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

// ---- A': the same with ONE sweep over the trailing matrix per column ------------------------------------------------
// The rank-2 update of column i-1 is not applied in a sweep of its own: it stays pending (vp, wp) and is applied while
// the next column's product p = tau A v is formed — every trailing element is read once, updated, stored and used.
// The pivot column gets the pending update on the fly (every CTA, redundantly); two cluster barriers per column.
__global__ void __cluster_dims__(EC_CLUSTER, 1, 1) __launch_bounds__(EC_THREADS, 1)
    sytrd_fused_kernel(double* __restrict__ Gall, int n, double* __restrict__ dall, double* __restrict__ eall,
                       double* __restrict__ tauall) {
  extern __shared__ double ec_smem[];
  double* vp = ec_smem;                      // pending reflector (of the previous column), indexed by matrix row
  double* wp = vp + EC_NMAX;                 // pending w
  double* vn = wp + EC_NMAX;                 // reflector of this column
  double* pw = vn + EC_NMAX;                 // p, then w of this column
  double* ppart = pw + EC_NMAX;              // this CTA's partial p
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
  const int colbase = rank + EC_CLUSTER * warp;
  const int colstep = EC_CLUSTER * EC_WARPS;
  double* myacc = accw + warp * EC_NMAX;
  bool pending = false;

  for (int i = 0; i + 1 < n; ++i) {
    const int r0 = i + 1;
    // 1. column i below the diagonal, pending update applied on the fly -> vn, its norm -> reflector
    const double vpi = pending ? vp[i] : 0.0, wpi = pending ? wp[i] : 0.0;
    double ss = 0.0;
    for (int r = r0 + tid; r < n; r += EC_THREADS) {
      double x = __ldcg(A + (size_t)i * n + r);
      if (pending) x -= vp[r] * wpi + wp[r] * vpi;
      vn[r] = x;
      if (r > r0) ss += x * x;
    }
    ss = ec_block_sum(ss, red);
    const double alpha = vn[r0];
    double beta = alpha, tau = 0.0;
    if (ss > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      tau = (beta - alpha) / beta;
      const double scale = 1.0 / (alpha - beta);
      __syncthreads();
      for (int r = r0 + 1 + tid; r < n; r += EC_THREADS) vn[r] *= scale;
    }
    __syncthreads();
    if (tid == 0) vn[r0] = 1.0;
    if (rank == 0 && tid == 0) {
      ee[i] = beta;
      tt[i] = tau;
      double di = __ldcg(A + (size_t)i * n + i);
      if (pending) di -= 2.0 * vpi * wpi;
      dd[i] = di;
    }
    __syncthreads();
    // 2. ONE sweep over this CTA's trailing columns: apply the pending update, store, accumulate p = A v
    const bool sym = tau != 0.0;                         // identical on every CTA
    if (sym) {
      for (int r = r0 + lane; r < n; r += 32) myacc[r] = 0.0;
      __syncwarp();
    }
    int q0 = (r0 - colbase + colstep - 1) / colstep;
    if (q0 < 0) q0 = 0;
    if (sym || pending) {
      for (int j = colbase + colstep * q0; j < n; j += colstep) {
        const double vnj = vn[j];
        const double vpj = pending ? vp[j] : 0.0, wpj = pending ? wp[j] : 0.0;
        double* col = A + (size_t)j * n;
        double dot = 0.0;
        for (int r = j + lane; r < n; r += 256) {
          double a[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = (r + 32 * u < n) ? __ldcg(col + r + 32 * u) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int rr = r + 32 * u;
            if (rr < n) {
              double x = a[u];
              if (pending) {
                x -= vp[rr] * wpj + wp[rr] * vpj;
                __stcg(col + rr, x);
              }
              if (sym) {
                dot += x * vn[rr];
                if (rr > j) myacc[rr] += x * vnj;
              }
            }
          }
        }
        if (sym) {
          dot = ec_warp_sum(dot);
          __syncwarp();
          if (lane == 0) myacc[j] += dot;
          __syncwarp();
        }
      }
    }
    if (!sym) {                                          // nothing to multiply: the trailing matrix is up to date
      pending = false;
      cluster.sync();
      continue;
    }
    __syncthreads();
    for (int r = r0 + tid; r < n; r += EC_THREADS) {
      double s = 0.0;
#pragma unroll
      for (int wv = 0; wv < EC_WARPS; ++wv) s += accw[wv * EC_NMAX + r];
      ppart[r] = s;
    }
    cluster.sync();                                      // #1: every CTA's partial p is in its shared memory
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
        for (int c = 0; c < EC_CLUSTER; ++c) cluster.map_shared_rank(pw, c)[r] = s;
      }
      if (rank == 0)                                     // everybody has read column i: keep the reflector there
        for (int r = r0 + 1 + tid; r < n; r += EC_THREADS) __stcg(A + (size_t)i * n + r, vn[r]);
    }
    cluster.sync();                                      // #2: p complete everywhere
    double pv = 0.0;
    for (int r = r0 + tid; r < n; r += EC_THREADS) pv += pw[r] * vn[r];
    pv = ec_block_sum(pv, red);
    const double half = 0.5 * tau * pv;
    for (int r = r0 + tid; r < n; r += EC_THREADS) pw[r] -= half * vn[r];
    __syncthreads();
    // this column's pair becomes the pending update of the next one (the same swap on every CTA of the cluster)
    double* t1 = vp; vp = vn; vn = t1;
    double* t2 = wp; wp = pw; pw = t2;
    pending = true;
  }
  if (rank == 0 && tid == 0) {
    double dl = __ldcg(A + (size_t)(n - 1) * n + (n - 1));
    if (pending) dl -= 2.0 * vp[n - 1] * wp[n - 1];
    dd[n - 1] = dl;
  }
}

// ---- B: bisection ------------------------------------------------------------------------------------------------
// lam_desc[mat][j] = j-th largest eigenvalue of the tridiagonal (d, e); also tnorm[mat] = max(|d|, |e|).
// Eight lanes share one eigenvalue: each evaluates the Sturm count at one of 8 points inside the bracket, which then
// shrinks to a ninth per round (21 rounds: 9^-21 of the Gershgorin interval, below one ulp of the norm).
constexpr int BS_LANES = 8;
constexpr int BS_THREADS = 128;
__global__ void __launch_bounds__(BS_THREADS) bisect_kernel(const double* __restrict__ dall, const double* __restrict__ eall,
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
  const int sub = threadIdx.x & (BS_LANES - 1);
  int j = blockIdx.x * (BS_THREADS / BS_LANES) + (threadIdx.x / BS_LANES);
  const bool live = j < k;
  if (!live) j = k - 1;                // idle groups shadow the last eigenvalue (the shuffles below need whole warps)
  const int want = n - 1 - j;          // count(x) <= want  <=>  x <= lambda_j
  const unsigned gshift = (threadIdx.x & 31) & ~(BS_LANES - 1);
  for (int it = 0; it < 21; ++it) {
    const double x = lo + (hi - lo) * ((double)(sub + 1) / (double)(BS_LANES + 1));
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    int c = q < 0.0;
    for (int i = 1; i < n; ++i) {
      q = d[i] - x - e2[i - 1] / q;
      if (fabs(q) < pivmin) q = -pivmin;
      c += q < 0.0;
    }
    const unsigned below = (__ballot_sync(0xffffffffu, c <= want) >> gshift) & ((1u << BS_LANES) - 1u);
    const int nb = __popc(below);      // the counts rise with x: the first nb points are at or below the eigenvalue
    const double xlo = __shfl_sync(0xffffffffu, x, (int)gshift + max(nb - 1, 0));
    const double xhi = __shfl_sync(0xffffffffu, x, (int)gshift + min(nb, BS_LANES - 1));
    if (nb > 0) lo = xlo;
    if (nb < BS_LANES) hi = xhi;
  }
  if (live && sub == 0) lamdesc[(size_t)mat * k + j] = 0.5 * (lo + hi);
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

constexpr int IV_TILE = 8;      // rows whose factors are loaded ahead of the dependent recurrences
__global__ void __launch_bounds__(64) invit_kernel(const double* __restrict__ dall, const double* __restrict__ eall,
                                                   int n, int k, int kpad, const double* __restrict__ lamdesc,
                                                   const double* __restrict__ tnorm_in, double* __restrict__ fac,
                                                   unsigned char* __restrict__ swp, double* __restrict__ Z, int iters) {
  extern __shared__ double iv_smem[];
  double* d = iv_smem;            // [n]
  double* e = d + n;              // [n]
  const int mat = blockIdx.y;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    d[i] = dall[(size_t)mat * n + i];
    e[i] = i + 1 < n ? eall[(size_t)mat * n + i] : 0.0;
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
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
  // factor T - lam I = P L U (dlagtf); the start vector is written alongside
  double cur_a = d[0] - lam, cur_b = n > 1 ? e[0] : 0.0;
  for (int i = 0; i + 1 < n; ++i) {
    const double ci = e[i];
    const double next_a = d[i + 1] - lam;
    const double next_b = i + 2 < n ? e[i + 1] : 0.0;
    const size_t o = (size_t)i * kpad;
    X[o] = ec_start_value(i, j);
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
    Aa[o] = cur_a; Bb[o] = 0.0; D2[o] = 0.0; Cc[o] = 0.0; SW[o] = 0;
    X[o] = ec_start_value(n - 1, j);
  }
  double scale = 1.0, ss = 0.0;
  for (int it = 0; it < iters; ++it) {
    // forward: the row operations applied to the right-hand side (scaled by the previous maximum); the factors of
    // IV_TILE rows are loaded before the dependent chain runs over them
    double carry = X[0] * scale;
    for (int i0 = 0; i0 + 1 < n; i0 += IV_TILE) {
      double xn[IV_TILE], m[IV_TILE];
      unsigned char sw[IV_TILE];
#pragma unroll
      for (int u = 0; u < IV_TILE; ++u) {
        const int i = i0 + u;
        const size_t o = (size_t)i * kpad;
        const bool in = i + 1 < n;
        xn[u] = in ? X[o + kpad] : 0.0;
        m[u] = in ? Cc[o] : 0.0;
        sw[u] = in ? SW[o] : 0;
      }
#pragma unroll
      for (int u = 0; u < IV_TILE; ++u) {
        const int i = i0 + u;
        if (i + 1 < n) {
          const double x1 = xn[u] * scale;
          if (sw[u]) {
            X[(size_t)i * kpad] = x1;
            carry = carry - m[u] * x1;
          } else {
            X[(size_t)i * kpad] = carry;
            carry = x1 - m[u] * carry;
          }
        }
      }
    }
    X[(size_t)(n - 1) * kpad] = carry;
    // backward substitution with U = (A, B, D2), tiny pivots replaced
    double x1 = 0.0, x2 = 0.0, xmax = 0.0;
    ss = 0.0;
    for (int i0 = n - 1; i0 >= 0; i0 -= IV_TILE) {
      double xr[IV_TILE], bb[IV_TILE], dd2[IV_TILE], aa[IV_TILE];
#pragma unroll
      for (int u = 0; u < IV_TILE; ++u) {
        const int i = i0 - u;
        const size_t o = (size_t)(i >= 0 ? i : 0) * kpad;
        xr[u] = X[o]; bb[u] = Bb[o]; dd2[u] = D2[o]; aa[u] = Aa[o];
      }
#pragma unroll
      for (int u = 0; u < IV_TILE; ++u) {
        const int i = i0 - u;
        if (i >= 0) {
          const double t = xr[u] - bb[u] * x1 - dd2[u] * x2;
          double piv = aa[u];
          if (fabs(piv) < tol) piv = piv < 0.0 ? -tol : tol;
          const double xi = t / piv;
          X[(size_t)i * kpad] = xi;
          xmax = fmax(xmax, fabs(xi));
          ss += xi * xi;
          x2 = x1;
          x1 = xi;
        }
      }
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
  // Z in the scratch layout [mat][row][kpad]: the back-transformation reads it transposed
  for (int i = 0; i < n; ++i) X[(size_t)i * kpad] *= rn;
}

// ---- D: back-transformation ------------------------------------------------------------------------------------------
// W[mat][k-1-j][:] = H_0 H_1 ... H_{n-2} z_j  (rows in ASCENDING order of eigenvalue, like syevd's output).  One warp
// per eigenvector, the vector in registers; the block stages reflector i-1 in shared memory while it applies reflector i.
constexpr int BT_WARPS = 8;
constexpr int BT_THREADS = BT_WARPS * 32;
__global__ void __launch_bounds__(BT_THREADS, 2) backtransform_kernel(const double* __restrict__ Gall, int n, int k, int kpad,
                                                                   const double* __restrict__ tauall,
                                                                   const double* __restrict__ fac,
                                                                   const double* __restrict__ lamdesc,
                                                                   double* __restrict__ W, double* __restrict__ lam_asc) {
  __shared__ double vbuf[2][EC_NMAX];
  const int mat = blockIdx.y;
  const double* A = Gall + (size_t)mat * n * n;
  const double* tau = tauall + (size_t)mat * n;
  const double* Zt = fac + ((size_t)mat * 5 + 4) * (size_t)n * kpad;     // unit eigenvectors of T, [row][kpad]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * BT_WARPS + warp;
  const bool live = j < k;
  double z[EC_NMAX / 32];
#pragma unroll
  for (int t = 0; t < EC_NMAX / 32; ++t) {
    const int r = lane + 32 * t;
    z[t] = (live && r < n) ? Zt[(size_t)r * kpad + j] : 0.0;
  }
  constexpr int PRE = EC_NMAX / BT_THREADS;
  // both staging buffers start as zeros: reflector i only ever writes rows >= i + 1 (< n), and the reflectors are
  // staged in DEcreasing i, so every row a buffer holds below the current reflector's first row is still zero, and
  // rows >= n stay zero.  Reflector n-2 (a single 1 at row n-1) needs no data: it starts staged.
  for (int r = threadIdx.x; r < EC_NMAX; r += BT_THREADS) {
    vbuf[0][r] = 0.0;
    vbuf[1][r] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) vbuf[0][n - 1] = 1.0;
  __syncthreads();
  int buf = 0;
  for (int i = n - 2; i >= 0; --i) {
    const int r0 = i + 1;
    // stage reflector i-1 (rows >= i+1 from column i-1, a 1 at row i) while reflector i is applied
    double pre[PRE];
#pragma unroll
    for (int u = 0; u < PRE; ++u) {
      const int r = r0 + threadIdx.x + u * BT_THREADS;
      pre[u] = (i > 0 && r < n) ? __ldcg(A + (size_t)(i - 1) * n + r) : 0.0;
    }
    const double ti = tau[i];
    if (ti != 0.0) {                                     // uniform over the block
      const double* v = vbuf[buf];
      // rows below r0 hold zeros in the staged reflector (vbuf is zero-filled once and only rows >= r0 are ever
      // written for reflector i), so whole 32-row groups need no row test; four independent partial sums
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const int t0 = r0 >> 5;
#pragma unroll
      for (int t = 0; t < EC_NMAX / 32; t += 4) {
        if (t + 3 < t0) continue;
        s0 += v[lane + 32 * t] * z[t];
        s1 += v[lane + 32 * (t + 1)] * z[t + 1];
        s2 += v[lane + 32 * (t + 2)] * z[t + 2];
        s3 += v[lane + 32 * (t + 3)] * z[t + 3];
      }
      const double s = ec_warp_sum((s0 + s1) + (s2 + s3)) * ti;
#pragma unroll
      for (int t = 0; t < EC_NMAX / 32; t += 4) {
        if (t + 3 < t0) continue;
        z[t] -= s * v[lane + 32 * t];
        z[t + 1] -= s * v[lane + 32 * (t + 1)];
        z[t + 2] -= s * v[lane + 32 * (t + 2)];
        z[t + 3] -= s * v[lane + 32 * (t + 3)];
      }
    }
    if (i > 0) {
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int r = r0 + threadIdx.x + u * BT_THREADS;
        if (r < n) vbuf[buf ^ 1][r] = pre[u];
      }
      if (threadIdx.x == 0) vbuf[buf ^ 1][i] = 1.0;
    }
    __syncthreads();
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
constexpr int CK_SPLIT = 8;
__global__ void __launch_bounds__(256) check_kernel(const double* __restrict__ W, const double* __restrict__ lam_asc,
                                                    int n, int k, int* __restrict__ status) {
  const int mat = blockIdx.x;
  const double* Wm = W + (size_t)mat * k * n;
  const double* lam = lam_asc + (size_t)mat * k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x >> 5) * CK_SPLIT;
  const double top = fabs(lam[k - 1]);
  int bad = 0;
  for (int r = warp + (blockDim.x >> 5) * blockIdx.y; r < k; r += nw) {
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
  size_t doubles = (size_t)batch * ((size_t)3 * n + k + 1 + (size_t)5 * n * kpad);
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
  double* fac = tnorm + batch;
  unsigned char* swp = reinterpret_cast<unsigned char*>(fac + (size_t)batch * 5 * n * kpad);
  FM_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)batch, st));
  {
    // FM_SYTRD_IMPL=0 selects the two-sweep kernel (kept for comparison); the default applies the rank-2 update and
    // forms the next product in one sweep.  (A third variant — lanes owning absolute rows with their share of the
    // product in registers, 16 loads in flight, serpentine column dealing, 8 warps of 254 registers — measured 20.7 ms
    // against 18.3 ms for 15 x 999^2 and was dropped: the sweep is bound by the latency of its dependent chain at low
    // occupancy, not by shared-memory traffic; profiles/r02_eig.md.)
    static const int impl = getenv("FM_SYTRD_IMPL") ? atoi(getenv("FM_SYTRD_IMPL")) : 1;
    const size_t smem = sizeof(double) * ((size_t)(5 + EC_WARPS) * EC_NMAX + 32);
    static bool configured = false;
    if (!configured) {
      FM_CUDA_TRY(cudaFuncSetAttribute(sytrd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FM_CUDA_TRY(cudaFuncSetAttribute(sytrd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    if (impl == 0)
      sytrd_cluster_kernel<<<dim3((unsigned)(batch * EC_CLUSTER)), EC_THREADS, smem, st>>>(G, n, d, e, tau);
    else
      sytrd_fused_kernel<<<dim3((unsigned)(batch * EC_CLUSTER)), EC_THREADS, smem, st>>>(G, n, d, e, tau);
    FM_LAUNCH_CHECK();
  }
  {
    const size_t smem = sizeof(double) * 2 * (size_t)n;
    bisect_kernel<<<dim3((unsigned)cdiv(k, BS_THREADS / BS_LANES), (unsigned)batch), BS_THREADS, smem, st>>>(d, e, n, k, lamdesc, tnorm);
    FM_LAUNCH_CHECK();
    invit_kernel<<<dim3((unsigned)cdiv(k, 64), (unsigned)batch), 64, smem, st>>>(d, e, n, k, (int)kpad, lamdesc, tnorm, fac, swp, nullptr, 2);
    FM_LAUNCH_CHECK();
  }
  backtransform_kernel<<<dim3((unsigned)cdiv(k, BT_WARPS), (unsigned)batch), BT_THREADS, 0, st>>>(G, n, k, (int)kpad, tau, fac, lamdesc, W, lam);
  FM_LAUNCH_CHECK();
  check_kernel<<<dim3((unsigned)batch, CK_SPLIT), 256, 0, st>>>(W, lam, n, k, status);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
