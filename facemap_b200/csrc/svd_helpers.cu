// Small kernels around the GEMMs: PCA centring terms, split-K reduction into the float64 Gram,
// top-k extraction with the reference's sign rule, transpose, ROI gather.
// Reference seams: facemap/utils.py:751-776 (svdecon) -> sklearn/decomposition/_pca.py:733-758,
// sklearn/utils/extmath.py:924-982 (svd_flip); facemap/process.py:218-222, 479-483 (ROI slices).
#include "common.cuh"
#include "slice_math.cuh"

namespace fm {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mean[r] = sum_c X[r,c] / ncols, float64 accumulation, one block per row.
__global__ void __launch_bounds__(256) row_means_kernel(const float* __restrict__ X, int64_t ldx, int64_t ncols,
                                                        double* __restrict__ mean) {
  __shared__ double sh[8];
  const float* row = X + (size_t)blockIdx.x * ldx;
  double acc = 0.0;
  const bool v4 = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (v4) {
    const int64_t n4 = ncols / 4;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int64_t i = threadIdx.x; i < n4; i += 256) {
      const float4 v = r4[i];
      acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    for (int64_t i = n4 * 4 + threadIdx.x; i < ncols; i += 256) acc += (double)row[i];
  } else {
    for (int64_t i = threadIdx.x; i < ncols; i += 256) acc += (double)row[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) mean[blockIdx.x] = v / (double)ncols;
  }
}

// G[i,j] = sum_s part[s,i,j] - ncols*mean[i]*mean[j]; with upper_only the (i>j) entries are read
// from (j,i).
// A thread reduces 4 adjacent columns of the UPPER triangle (float4 loads over the split planes) and
// writes both G[i,j] and its mirror G[j,i], so the lower triangle is never read.
__global__ void gram_assemble_kernel(const float* __restrict__ part, int ksplit, int n, int64_t ldw,
                                     const double* __restrict__ mean, double ncols, int upper_only,
                                     double* __restrict__ G) {
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j0 >= n) return;
  if (upper_only && j0 + 3 < i) return;                     // entirely below the diagonal: mirrored from above
  const size_t plane = (size_t)n * ldw;
  const float* src = part + (size_t)i * ldw + j0;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (j0 + 4 <= n && (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(part) & 15) == 0)) {
#pragma unroll 4
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * plane);
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
  } else {
    for (int s = 0; s < ksplit; ++s) {
      const float* q = src + (size_t)s * plane;
      a0 += q[0];
      if (j0 + 1 < n) a1 += q[1];
      if (j0 + 2 < n) a2 += q[2];
      if (j0 + 3 < n) a3 += q[3];
    }
  }
  const double acc[4] = {a0, a1, a2, a3};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = j0 + t;
    if (j >= n) break;
    if (upper_only && j < i) continue;
    double v = acc[t];
    if (mean) v -= ncols * (mean[i] * mean[j]);  // commutative product: G stays exactly symmetric
    G[(size_t)i * n + j] = v;
    if (upper_only && j != i) G[(size_t)j * n + i] = v;
  }
}

// one block per component c
__global__ void __launch_bounds__(256) topk_kernel(const double* __restrict__ evec, const double* __restrict__ lam,
                                                   int n, int nev, int k, const double* __restrict__ mean,
                                                   float* __restrict__ Wt, int64_t ldwt, float* __restrict__ sval,
                                                   float* __restrict__ rbias, float* __restrict__ rscale) {
  __shared__ double s_abs[256];
  __shared__ int s_idx[256];
  __shared__ double s_dot[8];
  __shared__ double s_sign;
  const int c = blockIdx.x;
  const double* v = evec + (size_t)(nev - 1 - c) * n;
  // argmax |v| with the smallest index on ties (np.argmax)
  double best = -1.0;
  int bidx = 0;
  for (int j = threadIdx.x; j < n; j += 256) {
    const double a = fabs(v[j]);
    if (a > best) { best = a; bidx = j; }
  }
  s_abs[threadIdx.x] = best;
  s_idx[threadIdx.x] = bidx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double a = s_abs[threadIdx.x + o];
      const int ia = s_idx[threadIdx.x + o];
      if (a > s_abs[threadIdx.x] || (a == s_abs[threadIdx.x] && ia < s_idx[threadIdx.x])) {
        s_abs[threadIdx.x] = a;
        s_idx[threadIdx.x] = ia;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) s_sign = (v[s_idx[0]] < 0.0) ? -1.0 : 1.0;
  __syncthreads();
  const double sg = s_sign;
  double dot = 0.0;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float w = (float)(sg * v[j]);
    Wt[(size_t)c * ldwt + j] = w;
    if (mean) dot += (double)w * mean[j];
  }
  for (int64_t j = n + threadIdx.x; j < ldwt; j += 256) Wt[(size_t)c * ldwt + j] = 0.f;  // pad
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) s_dot[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < 8; ++w) d += s_dot[w];
    const double l = lam[nev - 1 - c];
    const float s = (float)sqrt(l > 0.0 ? l : 0.0);
    const float inv = s > 0.f ? 1.0f / s : 0.f;
    if (sval) sval[c] = s;
    if (rscale) rscale[c] = inv;
    // epilogue form rscale*acc + rbias: with a scale the bias is pre-multiplied
    if (rbias) rbias[c] = rscale ? (float)(-d * (double)inv) : (float)(-d);
  }
}

template <bool SWAP_GRID>
__global__ void transpose_kernel(const float* __restrict__ in, int64_t ldin, int rows, int cols,
                                 float* __restrict__ out, int64_t ldout) {
  __shared__ float tile[32][33];
  const int c0 = (SWAP_GRID ? blockIdx.y : blockIdx.x) * 32, r0 = (SWAP_GRID ? blockIdx.x : blockIdx.y) * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ldin + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * ldout + r] = tile[threadIdx.x][i];
  }
}

__global__ void gather_roi_kernel(const float* __restrict__ X, int64_t ldx, int rows, int64_t col0, int Lxb,
                                  int y0, int x0, int ny, int nx, float* __restrict__ out, int64_t ldout) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (q >= ny * nx) return;
  const int y = q / nx, x = q - y * nx;
  out[(size_t)r * ldout + q] = X[(size_t)r * ldx + col0 + (size_t)(y0 + y) * Lxb + (x0 + x)];
}

// the same gather on a byte plane (the integer operands of fm_project_i8)
__global__ void gather_roi_u8_kernel(const uint8_t* __restrict__ X, int64_t ldx, int rows, int64_t col0, int Lxb,
                                     int y0, int x0, int ny, int nx, uint8_t* __restrict__ out, int64_t ldout) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (q >= ny * nx) return;
  const int y = q / nx, x = q - y * nx;
  out[(size_t)r * ldout + q] = X[(size_t)r * ldx + col0 + (size_t)(y0 + y) * Lxb + (x0 + x)];
}

}  // namespace fm

extern "C" int fm_row_means(const float* X, int64_t ldx, int rows, int64_t ncols, double* mean, void* stream) {
  using namespace fm;
  FM_REQUIRE(X && mean && rows >= 0 && ncols >= 1 && ldx >= ncols, "fm_row_means: bad args");
  if (rows == 0) return FM_OK;
  row_means_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, ncols, mean);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gram_assemble(const float* part, int ksplit, int n, int64_t ldw, const double* mean,
                                int64_t ncols, int upper_only, double* G, void* stream) {
  using namespace fm;
  FM_REQUIRE(part && G && ksplit >= 1 && n >= 1 && ldw >= n, "fm_gram_assemble: bad args");
  dim3 block(32, 8), grid((unsigned)cdiv(cdiv(n, 4), 32), (unsigned)cdiv(n, 8));
  gram_assemble_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      part, ksplit, n, ldw, mean, (double)ncols, upper_only, G);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_topk_components(const double* evec, const double* lam, int n, int nev, int k,
                                  const double* mean, float* Wt, int64_t ldwt, float* sval, float* rbias,
                                  float* rscale, void* stream) {
  using namespace fm;
  FM_REQUIRE(evec && lam && Wt && n >= 1 && nev >= 1 && nev <= n && k >= 1 && k <= nev && ldwt >= n,
             "fm_topk_components: bad args");
  topk_kernel<<<k, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(evec, lam, n, nev, k, mean, Wt, ldwt, sval,
                                                                    rbias, rscale);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_transpose(const float* in, int64_t ldin, int rows, int cols, float* out, int64_t ldout,
                            void* stream) {
  using namespace fm;
  FM_REQUIRE(in && out && rows >= 0 && cols >= 0 && ldin >= cols && ldout >= rows, "fm_transpose: bad args");
  if (rows == 0 || cols == 0) return FM_OK;
  // the longer dimension goes on grid.x (2^31 blocks), the shorter on grid.y (65535 blocks = 2M elements)
  dim3 block(32, 8), grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32));
  const bool swap = grid.y > 65535;
  if (swap) grid = dim3((unsigned)cdiv(rows, 32), (unsigned)cdiv(cols, 32));
  FM_REQUIRE(grid.y <= 65535, "fm_transpose: matrix too large (%d x %d)", rows, cols);
  if (swap)
    transpose_kernel<true><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, ldin, rows, cols, out, ldout);
  else
    transpose_kernel<false><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, ldin, rows, cols, out, ldout);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gather_roi(const float* X, int64_t ldx, int rows, int64_t col0, int Lxb, int y0, int y1,
                             int x0, int x1, float* out, int64_t ldout, void* stream) {
  using namespace fm;
  const int ny = y1 - y0, nx = x1 - x0;
  FM_REQUIRE(X && out && rows >= 0 && ny >= 1 && nx >= 1 && y0 >= 0 && x0 >= 0 && col0 >= 0 && x1 <= Lxb &&
                 col0 + (int64_t)y1 * Lxb <= ldx && ldout >= (int64_t)ny * nx,
             "fm_gather_roi: rectangle rows [%d, %d) x columns [%d, %d) does not fit a view of %d columns at column %lld of "
             "a %lld-wide operand", y0, y1, x0, x1, Lxb, (long long)col0, (long long)ldx);
  if (rows == 0) return FM_OK;
  FM_REQUIRE(rows <= 65535, "fm_gather_roi: too many rows per call (%d)", rows);
  dim3 grid((unsigned)cdiv((int64_t)ny * nx, 256), (unsigned)rows);
  gather_roi_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, rows, col0, Lxb, y0, x0, ny,
                                                                             nx, out, ldout);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gather_roi_u8(const uint8_t* X, int64_t ldx, int rows, int64_t col0, int Lxb, int y0, int y1,
                                int x0, int x1, uint8_t* out, int64_t ldout, void* stream) {
  using namespace fm;
  const int ny = y1 - y0, nx = x1 - x0;
  FM_REQUIRE(X && out && rows >= 0 && ny >= 1 && nx >= 1 && y0 >= 0 && x0 >= 0 && col0 >= 0 && x1 <= Lxb &&
                 col0 + (int64_t)y1 * Lxb <= ldx && ldout >= (int64_t)ny * nx,
             "fm_gather_roi_u8: rectangle rows [%d, %d) x columns [%d, %d) does not fit a view of %d columns at column %lld "
             "of a %lld-wide operand", y0, y1, x0, x1, Lxb, (long long)col0, (long long)ldx);
  if (rows == 0) return FM_OK;
  FM_REQUIRE(rows <= 65535, "fm_gather_roi_u8: too many rows per call (%d)", rows);
  dim3 grid((unsigned)cdiv((int64_t)ny * nx, 256), (unsigned)rows);
  gather_roi_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, rows, col0, Lxb, y0, x0, ny,
                                                                                nx, out, ldout);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// C[m,n] = rscale[m] * sum_s part[s,m,n] + rbias[m]: reduction of split-K partial products in
// float64 (the tensor cores accumulate with truncation, so short K chains + an exact reduction
// keep the GEMMs at float32 grade).
namespace fm {
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int ksplit, int M, int N, int64_t ldw,
                                     const float* __restrict__ rscale, const float* __restrict__ rbias,
                                     float* __restrict__ C, int64_t ldc) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int m = blockIdx.y;
  if (n >= N) return;
  const size_t plane = (size_t)M * ldw;
  const float* src = part + (size_t)m * ldw + n;
  const double sc = rscale ? (double)rscale[m] : 1.0;
  const double bi = rbias ? (double)rbias[m] : 0.0;
  if (n + 4 <= N && (ldw % 4 == 0) && (ldc % 4 == 0) &&
      (((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(C)) & 15) == 0)) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t)s * plane);
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) =
        make_float4((float)(a0 * sc + bi), (float)(a1 * sc + bi), (float)(a2 * sc + bi), (float)(a3 * sc + bi));
  } else {
    for (int j = 0; j < 4 && n + j < N; ++j) {
      double a = 0;
      for (int s = 0; s < ksplit; ++s) a += (double)src[(size_t)s * plane + j];
      C[(size_t)m * ldc + n + j] = (float)(a * sc + bi);
    }
  }
}
}  // namespace fm

extern "C" int fm_splitk_reduce(const float* part, int ksplit, int M, int N, int64_t ldw, const float* rscale,
                                const float* rbias, float* C, int64_t ldc, void* stream) {
  using namespace fm;
  FM_REQUIRE(part && C && ksplit >= 1 && M >= 1 && N >= 1 && ldw >= N && ldc >= N, "fm_splitk_reduce: bad args");
  FM_REQUIRE(M <= 65535 * 16, "fm_splitk_reduce: too many rows");
  dim3 grid((unsigned)cdiv(cdiv(N, 4), 128), (unsigned)M);
  if (M > 65535) {  // fold rows into z for very tall outputs
    set_error("fm_splitk_reduce: M > 65535 unsupported");
    return FM_ERR_INVALID;
  }
  splitk_reduce_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, ksplit, M, N, ldw, rscale,
                                                                                rbias, C, ldc);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// lo[r,c] = x - tf32_trunc(x): the second operand plane of the 3xTF32 split for a matrix that is
// reused by many GEMMs (the projection masks).
namespace fm {
__global__ void split_lo_kernel(const float* __restrict__ X, int64_t ldx, int rows, int cols, float* __restrict__ lo,
                                int64_t ldlo) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const bool v4 = (c + 4 <= cols) && (ldx % 4 == 0) && (ldlo % 4 == 0) &&
                  (((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0);
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    if (v4) {
      const float4 x = *reinterpret_cast<const float4*>(X + (size_t)r * ldx + c);
      float4 l;
      l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
      l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
      l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      *reinterpret_cast<float4*>(lo + (size_t)r * ldlo + c) = l;
    } else {
      for (int t = 0; t < 4 && c + t < cols; ++t) {
        const float x = X[(size_t)r * ldx + c + t];
        lo[(size_t)r * ldlo + c + t] = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      }
    }
  }
}
}  // namespace fm

extern "C" int fm_split_lo(const float* X, int64_t ldx, int rows, int cols, float* lo, int64_t ldlo, void* stream) {
  using namespace fm;
  FM_REQUIRE(X && lo && rows >= 1 && cols >= 1 && ldx >= cols && ldlo >= cols, "fm_split_lo: bad args");
  dim3 grid((unsigned)cdiv(cdiv(cols, 4), 256), (unsigned)(rows < 65535 ? rows : 65535));
  split_lo_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, rows, cols, lo, ldlo);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// ---- 8-bit row slices of a Gram operand (svd.sliced_gram) -----------------------------------------
// One block per row: |max| of the row -> its slice grid -> three planes (slice_math.cuh).  The row is read twice
// (the second time from L2); pad columns [ncols, lds) are zeroed so the planes are clean TMA operands.
namespace fm {
__global__ void __launch_bounds__(256) slice_rows_kernel(const float* __restrict__ X, int64_t ldx, int64_t ncols,
                                                         float* __restrict__ S0, float* __restrict__ S1,
                                                         float* __restrict__ S2, int64_t lds) {
  __shared__ float sh[8];
  __shared__ int s_ex;
  const float* row = X + (size_t)blockIdx.x * ldx;
  float amax = 0.0f;
  for (int64_t i = threadIdx.x; i < ncols; i += 256) amax = fmaxf(amax, fabsf(row[i]));
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = sh[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, sh[i]);
    s_ex = slice_exponent(v);
  }
  __syncthreads();
  const SliceGrid g = slice_grid(s_ex);
  float* o0 = S0 + (size_t)blockIdx.x * lds;
  float* o1 = S1 + (size_t)blockIdx.x * lds;
  float* o2 = S2 + (size_t)blockIdx.x * lds;
  for (int64_t i = threadIdx.x; i < lds; i += 256) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    if (i < ncols) slice3(row[i], g, a0, a1, a2);
    o0[i] = a0;
    o1[i] = a1;
    o2[i] = a2;
  }
}
}  // namespace fm

extern "C" int fm_slice_rows(const float* X, int64_t ldx, int rows, int64_t ncols, float* S0, float* S1, float* S2,
                             int64_t lds, void* stream) {
  using namespace fm;
  FM_REQUIRE(X && S0 && S1 && S2 && rows >= 0 && ncols >= 1 && ldx >= ncols && lds >= ncols, "fm_slice_rows: bad args");
  if (rows == 0) return FM_OK;
  slice_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, ncols, S0, S1, S2, lds);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// ---- integer digit planes and their Gram assembly (EXPERIMENTAL, svd.i8_gram; see gemm_i8.cu) -------------
namespace fm {
// one block per row: |max| -> exponent -> s8 / u8 / u8 digit planes (slice_math.cuh:slice3_i8); pad bytes zeroed
__global__ void __launch_bounds__(256) slice_rows_i8_kernel(const float* __restrict__ X, int64_t ldx, int64_t ncols,
                                                            int8_t* __restrict__ Q0, uint8_t* __restrict__ Q1,
                                                            uint8_t* __restrict__ Q2, int64_t ldq,
                                                            int32_t* __restrict__ exps) {
  __shared__ float sh[8];
  __shared__ int s_ex;
  const float* row = X + (size_t)blockIdx.x * ldx;
  float amax = 0.0f;
  for (int64_t i = threadIdx.x; i < ncols; i += 256) amax = fmaxf(amax, fabsf(row[i]));
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = sh[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, sh[i]);
    s_ex = slice_exponent(v);
    exps[blockIdx.x] = s_ex;
  }
  __syncthreads();
  const int ex = s_ex;
  int8_t* o0 = Q0 + (size_t)blockIdx.x * ldq;
  uint8_t* o1 = Q1 + (size_t)blockIdx.x * ldq;
  uint8_t* o2 = Q2 + (size_t)blockIdx.x * ldq;
  // four elements per thread: one packed 32-bit store per plane (ldq is a multiple of 4 and the planes are 4-byte
  // aligned, checked by the caller), float4 loads where the source row allows
  const bool v4 = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (int64_t i = (int64_t)threadIdx.x * 4; i < ldq; i += 1024) {
    float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (v4 && i + 4 <= ncols) {
      const float4 v = *reinterpret_cast<const float4*>(row + i);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (i + t < ncols) x[t] = row[i + t];
    }
    uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      signed char a0 = 0;
      unsigned char a1 = 0, a2 = 0;
      if (i + t < ncols) slice3_i8(x[t], ex, a0, a1, a2);
      w0 |= (uint32_t)(unsigned char)a0 << (8 * t);
      w1 |= (uint32_t)a1 << (8 * t);
      w2 |= (uint32_t)a2 << (8 * t);
    }
    *reinterpret_cast<uint32_t*>(o0 + i) = w0;
    *reinterpret_cast<uint32_t*>(o1 + i) = w1;
    *reinterpret_cast<uint32_t*>(o2 + i) = w2;
  }
}

// P: int32 [6][ksplit][n][ldp] products of the digit planes in the order 00, 11, 22 (upper triangles), 01, 02, 12
// (full), P_de[i,j] = sum_k Q_d[i,k] Q_e[j,k].  With y ~ 2^(e-7) (q0 + q1/2^8 + q2/2^16):
//   G[i,j] = 2^(e_i+e_j-14) [ P00 + (P01[i,j] + P01[j,i]) / 2^8 + (P11 + P02[i,j] + P02[j,i]) / 2^16
//                             + (P12[i,j] + P12[j,i]) / 2^24 + P22 / 2^32 ] - ncols m_i m_j        (j >= i, mirrored)
// Everything inside the bracket is an exact integer combination evaluated in float64.
// ROWSIDE: the operand rows are the pixels of a wide matrix Y [p, c] and the centring is over its columns
// (gram_assemble_rowside_kernel): the correction is  - q_i - q_j + mu.mu  instead of  - ncols m_i m_j.
template <bool ROWSIDE>
__global__ void gram_assemble_i8_kernel(const int32_t* __restrict__ P, int ksplit, int n, int64_t ldp,
                                        const int32_t* __restrict__ exps, const double* __restrict__ mean,
                                        double ncols, const double* __restrict__ mu, int nmu,
                                        double* __restrict__ G, int accumulate) {
  __shared__ double s_mumu;
  if (ROWSIDE) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      double a = 0.0;
      for (int t = 0; t < nmu; ++t) a += mu[t] * mu[t];
      s_mumu = a;
    }
    __syncthreads();
  }
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j >= n || j < i) return;
  const size_t plane = (size_t)n * ldp;
  const size_t prod = (size_t)ksplit * plane;
  const size_t ij = (size_t)i * ldp + j, ji = (size_t)j * ldp + i;
  long long s00 = 0, s11 = 0, s22 = 0, s01 = 0, s02 = 0, s12 = 0;
  for (int s = 0; s < ksplit; ++s) {
    const int32_t* base = P + (size_t)s * plane;
    s00 += base[ij];
    s11 += base[prod + ij];
    s22 += base[2 * prod + ij];
    s01 += (long long)base[3 * prod + ij] + base[3 * prod + ji];
    s02 += (long long)base[4 * prod + ij] + base[4 * prod + ji];
    s12 += (long long)base[5 * prod + ij] + base[5 * prod + ji];
  }
  double v = (double)s00 + (double)s01 * (1.0 / 256.0) + ((double)s11 + (double)s02) * (1.0 / 65536.0) +
             (double)s12 * (1.0 / 16777216.0) + (double)s22 * (1.0 / 4294967296.0);
  v = ldexp(v, exps[i] + exps[j] - 14);
  if (ROWSIDE)
    v = v - (mean[i] + mean[j]) + s_mumu;          // `mean` carries q = Y mu here
  else if (mean)
    v -= ncols * (mean[i] * mean[j]);
  if (accumulate) v += G[(size_t)i * n + j];       // a further block of columns of the same operand
  G[(size_t)i * n + j] = v;
  if (j != i) G[(size_t)j * n + i] = v;
}
}  // namespace fm

extern "C" int fm_slice_rows_i8(const float* X, int64_t ldx, int rows, int64_t ncols, int8_t* Q0, uint8_t* Q1,
                                uint8_t* Q2, int64_t ldq, int32_t* exps, void* stream) {
  using namespace fm;
  FM_REQUIRE(X && Q0 && Q1 && Q2 && exps && rows >= 0 && ncols >= 1 && ldx >= ncols && ldq >= ncols,
             "fm_slice_rows_i8: bad args");
  FM_REQUIRE(ldq % 4 == 0 && ((reinterpret_cast<uintptr_t>(Q0) | reinterpret_cast<uintptr_t>(Q1) |
                               reinterpret_cast<uintptr_t>(Q2)) & 3) == 0,
             "fm_slice_rows_i8: planes need 4-byte aligned bases and a row pitch that is a multiple of 4");
  if (rows == 0) return FM_OK;
  slice_rows_i8_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, ldx, ncols, Q0, Q1, Q2, ldq, exps);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gram_assemble_i8(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps,
                                   const double* mean, int64_t ncols, double* G, void* stream) {
  using namespace fm;
  FM_REQUIRE(P && exps && G && ksplit >= 1 && n >= 1 && ldp >= n, "fm_gram_assemble_i8: bad args");
  dim3 block(32, 8), grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 8));
  gram_assemble_i8_kernel<false><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      P, ksplit, n, ldp, exps, mean, (double)ncols, nullptr, 0, G, 0);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// G += the weighted plane products of a further block of columns (operands too wide to slice at once are processed
// in column blocks, each with its own row exponents; the mean term went in with the first block)
extern "C" int fm_gram_accumulate_i8(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps, double* G,
                                     void* stream) {
  using namespace fm;
  FM_REQUIRE(P && exps && G && ksplit >= 1 && n >= 1 && ldp >= n, "fm_gram_accumulate_i8: bad args");
  dim3 block(32, 8), grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 8));
  gram_assemble_i8_kernel<false><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      P, ksplit, n, ldp, exps, nullptr, 0.0, nullptr, 0, G, 1);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gram_assemble_i8_rowside(const int32_t* P, int ksplit, int n, int64_t ldp, const int32_t* exps,
                                           const double* q, const double* mu, int nmu, double* G, void* stream) {
  using namespace fm;
  FM_REQUIRE(P && exps && q && mu && G && ksplit >= 1 && n >= 1 && nmu >= 1 && ldp >= n,
             "fm_gram_assemble_i8_rowside: bad args");
  dim3 block(32, 8), grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 8));
  gram_assemble_i8_kernel<true><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      P, ksplit, n, ldp, exps, q, 0.0, mu, nmu, G, 0);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// ---- row-side Gram for wide matrices (merge of an ROI: p pixels < c stacked components) -------
// svdecon(Y [p, c]) with PCA column centring Yc = Y - 1 mu^T:
//   Yc Yc^T = Y Y^T - q 1^T - 1 q^T + (mu.mu) 1 1^T,   q = Y mu
// Its eigenvectors ARE the left singular vectors; the sign rule needs the right ones,
//   V S = Yc^T U = Y^T U - mu (1^T U).
namespace fm {
// q[r] = sum_c Y[r,c] * w[c]   (float64), one block per row
__global__ void __launch_bounds__(256) row_dot_kernel(const float* __restrict__ Y, int64_t ldy, int64_t ncols,
                                                      const double* __restrict__ w, double* __restrict__ q) {
  __shared__ double sh[8];
  const float* row = Y + (size_t)blockIdx.x * ldy;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < ncols; i += 256) acc += (double)row[i] * w[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += sh[i];
    q[blockIdx.x] = v;
  }
}

__global__ void gram_assemble_rowside_kernel(const float* __restrict__ part, int ksplit, int n, int64_t ldw,
                                             const double* __restrict__ q, const double* __restrict__ mu, int nmu,
                                             int upper_only, double* __restrict__ G) {
  __shared__ double s_mumu;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    double a = 0.0;
    for (int i = 0; i < nmu; ++i) a += mu[i] * mu[i];
    s_mumu = a;
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j >= n) return;
  int si = i, sj = j;
  if (upper_only && i > j) { si = j; sj = i; }
  double acc = 0.0;
  const size_t plane = (size_t)n * ldw;
  for (int s = 0; s < ksplit; ++s) acc += (double)part[(size_t)s * plane + (size_t)si * ldw + sj];
  G[(size_t)i * n + j] = acc - (q[i] + q[j]) + s_mumu;
}

// one block per component: sign of the largest-|.| entry of its RIGHT singular vector
//   v[j] = T[c,j] - mu[j] * sum_a Ut[c,a];  negative -> flip row c of Ut
__global__ void __launch_bounds__(256) sign_from_right_kernel(const float* __restrict__ T, int64_t ldt, int ncols,
                                                              const double* __restrict__ mu, float* __restrict__ Ut,
                                                              int64_t ldu, int p) {
  __shared__ double s_val[256];
  __shared__ double s_abs[256];
  __shared__ int s_idx[256];
  __shared__ double s_su;
  const int c = blockIdx.x;
  float* u = Ut + (size_t)c * ldu;
  double su = 0.0;
  for (int a = threadIdx.x; a < p; a += 256) su += (double)u[a];
  s_val[threadIdx.x] = su;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_val[threadIdx.x] += s_val[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_su = s_val[0];
  __syncthreads();
  su = s_su;
  double best = -1.0, bval = 0.0;
  int bidx = 0;
  for (int j = threadIdx.x; j < ncols; j += 256) {
    const double v = (double)T[(size_t)c * ldt + j] - mu[j] * su;
    if (fabs(v) > best) { best = fabs(v); bval = v; bidx = j; }
  }
  __syncthreads();
  s_abs[threadIdx.x] = best; s_val[threadIdx.x] = bval; s_idx[threadIdx.x] = bidx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double a = s_abs[threadIdx.x + o];
      if (a > s_abs[threadIdx.x] || (a == s_abs[threadIdx.x] && s_idx[threadIdx.x + o] < s_idx[threadIdx.x])) {
        s_abs[threadIdx.x] = a; s_val[threadIdx.x] = s_val[threadIdx.x + o]; s_idx[threadIdx.x] = s_idx[threadIdx.x + o];
      }
    }
    __syncthreads();
  }
  const bool flip = s_val[0] < 0.0;
  if (flip)
    for (int a = threadIdx.x; a < p; a += 256) u[a] = -u[a];
}
}  // namespace fm

extern "C" int fm_row_dot(const float* Y, int64_t ldy, int rows, int64_t ncols, const double* w, double* q,
                          void* stream) {
  using namespace fm;
  FM_REQUIRE(Y && w && q && rows >= 1 && ncols >= 1 && ldy >= ncols, "fm_row_dot: bad args");
  row_dot_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(Y, ldy, ncols, w, q);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_gram_assemble_rowside(const float* part, int ksplit, int n, int64_t ldw, const double* q,
                                        const double* mu, int nmu, int upper_only, double* G, void* stream) {
  using namespace fm;
  FM_REQUIRE(part && q && mu && G && ksplit >= 1 && n >= 1 && nmu >= 1 && ldw >= n, "fm_gram_assemble_rowside: bad args");
  dim3 block(32, 8), grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 8));
  gram_assemble_rowside_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, ksplit, n, ldw, q, mu,
                                                                                          nmu, upper_only, G);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_sign_from_right(const float* T, int64_t ldt, int k, int ncols, const double* mu, float* Ut,
                                  int64_t ldu, int p, void* stream) {
  using namespace fm;
  FM_REQUIRE(T && mu && Ut && k >= 1 && ncols >= 1 && p >= 1 && ldt >= ncols && ldu >= p, "fm_sign_from_right: bad args");
  sign_from_right_kernel<<<k, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(T, ldt, ncols, mu, Ut, ldu, p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}


// ---- K-blocked copy of a byte plane (operand layout of fm_project_i8_blocked) ------------------------------------------
namespace fm {
// dst[(c / kb) * block_stride + r * kb + c % kb] = src[r * ld + c] for c < cols, 0 for cols <= c < nblk * kb.
// One thread moves 16 bytes (kb, ld, cols offsets: multiples of 16 where the fast path applies).
__global__ void __launch_bounds__(256) retile_u8_kernel(const uint8_t* __restrict__ src, int64_t ld, int rows, int64_t cols,
                                                        uint8_t* __restrict__ dst, int64_t kb, int64_t block_stride,
                                                        int64_t cols_pad, int vec_ok) {
  const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int r = blockIdx.y;
  if (c0 >= cols_pad) return;
  const int64_t blk = c0 / kb, kin = c0 - blk * kb;
  uint8_t* out = dst + blk * block_stride + (int64_t)r * kb + kin;
  if (c0 + 16 <= cols && vec_ok) {
    *reinterpret_cast<uint4*>(out) = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c0);
    return;
  }
  for (int b = 0; b < 16; ++b) out[b] = (c0 + b < cols) ? src[(int64_t)r * ld + c0 + b] : (uint8_t)0;
}
}  // namespace fm

extern "C" int fm_retile_u8(const uint8_t* src, int64_t ld, int rows, int64_t cols, uint8_t* dst, int64_t kb,
                            int64_t block_stride, void* stream) {
  using namespace fm;
  FM_REQUIRE(src && dst && rows >= 0 && cols >= 1 && ld >= cols && kb >= 16 && kb % 16 == 0 && block_stride >= (int64_t)rows * kb,
             "fm_retile_u8: bad args (rows %d cols %lld ld %lld kb %lld block stride %lld)", rows, (long long)cols,
             (long long)ld, (long long)kb, (long long)block_stride);
  if (rows == 0) return FM_OK;
  FM_REQUIRE(rows <= 65535, "fm_retile_u8: too many rows per call (%d)", rows);
  const int64_t cols_pad = cdiv(cols, kb) * kb;
  const int vec_ok = (ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0) &&
                     (block_stride % 16 == 0);
  dim3 grid((unsigned)cdiv(cols_pad / 16, 256), (unsigned)rows);
  retile_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, ld, rows, cols, dst, kb, block_stride,
                                                                           cols_pad, vec_ok);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
