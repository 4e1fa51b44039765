// Symmetric eigensolve through cuSOLVER (syevd) — the single library call on the path.
// It stands in for the LAPACK work inside the reference's svdecon (facemap/utils.py:751-776:
// sklearn randomized_svd -> scipy lu/qr/gesdd; matlab/utils/svdecon.m:24 `eig`).  Its time is
// reported separately by the host and never claimed as a kernel of this library.
#include <cusolverDn.h>
#include <vector>
#include "common.cuh"

#define FM_SOLVER_INFO_SLOTS 4096

struct fm_solver {
  cusolverDnHandle_t handle = nullptr;
  cusolverDnParams_t params = nullptr;
  void* d_work = nullptr;
  size_t d_work_bytes = 0;
  void* h_work = nullptr;
  size_t h_work_bytes = 0;
  // devInfo of every solve since the last fm_solver_check: a ring on the device, read back (one copy, one
  // synchronisation) only when the caller asks — the solves themselves never block the host
  int* d_info = nullptr;
  int info_used = 0;
  int info_n[FM_SOLVER_INFO_SLOTS];     // matrix order of the solve that owns each slot (for the error text)
  float* d_f32 = nullptr;      // staging for the float32 solve
  size_t d_f32_elems = 0;
};

namespace fm {

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

#define FM_SOLVER_TRY(expr)                                                          \
  do {                                                                               \
    cusolverStatus_t _s = (expr);                                                    \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                             \
      fm::set_error("%s failed with cusolverStatus %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
      return FM_ERR_SOLVER;                                                          \
    }                                                                                \
  } while (0)

static int ensure_workspace(fm_solver* s, size_t dbytes, size_t hbytes) {
  if (dbytes > s->d_work_bytes) {
    if (s->d_work) cudaFree(s->d_work);
    s->d_work = nullptr;
    s->d_work_bytes = 0;
    FM_CUDA_TRY(cudaMalloc(&s->d_work, dbytes));
    s->d_work_bytes = dbytes;
  }
  if (hbytes > s->h_work_bytes) {
    free(s->h_work);
    s->h_work = malloc(hbytes);
    FM_REQUIRE(s->h_work != nullptr, "fm_syevd: host workspace allocation of %zu bytes failed", hbytes);
    s->h_work_bytes = hbytes;
  }
  return FM_OK;
}

// `count` consecutive devInfo slots for one call; a full ring is drained (this does synchronise, once per
// FM_SOLVER_INFO_SLOTS solves that nobody checked)
static int read_infos(fm_solver* s, cudaStream_t st);
static int take_info_slots(fm_solver* s, int count, int n, cudaStream_t st, int** out) {
  FM_REQUIRE(count <= FM_SOLVER_INFO_SLOTS, "batch of %d eigensolves exceeds the %d devInfo slots", count, FM_SOLVER_INFO_SLOTS);
  if (s->info_used + count > FM_SOLVER_INFO_SLOTS) {
    int rc = read_infos(s, st);
    if (rc != FM_OK) return rc;
  }
  *out = s->d_info + s->info_used;
  for (int i = 0; i < count; ++i) s->info_n[s->info_used + i] = n;
  s->info_used += count;
  return FM_OK;
}

static int read_infos(fm_solver* s, cudaStream_t st) {
  const int used = s->info_used;
  s->info_used = 0;
  if (used == 0) return FM_OK;
  std::vector<int> h((size_t)used);
  FM_CUDA_TRY(cudaMemcpyAsync(h.data(), s->d_info, sizeof(int) * (size_t)used, cudaMemcpyDeviceToHost, st));
  FM_CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < used; ++i)
    if (h[(size_t)i] != 0) {
      set_error("cuSOLVER eigensolve %d of %d since the last check: devInfo = %d (n = %d)", i, used, h[(size_t)i], s->info_n[i]);
      return FM_ERR_SOLVER;
    }
  return FM_OK;
}

}  // namespace fm

// Reads back the devInfo of every solve issued through `s` since the previous check (one copy + one stream
// synchronisation) and fails on the first non-zero entry.  The solves themselves are asynchronous.
extern "C" int fm_solver_check(fm_solver* s, void* stream) {
  using namespace fm;
  FM_REQUIRE(s != nullptr, "fm_solver_check: null solver");
  return read_infos(s, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fm_solver_create(fm_solver** out) {
  using namespace fm;
  FM_REQUIRE(out != nullptr, "fm_solver_create: null output");
  fm_solver* s = new fm_solver();
  cusolverStatus_t st = cusolverDnCreate(&s->handle);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnCreate failed with status %d", (int)st);
    delete s;
    return FM_ERR_SOLVER;
  }
  st = cusolverDnCreateParams(&s->params);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnCreateParams failed with status %d", (int)st);
    cusolverDnDestroy(s->handle);
    delete s;
    return FM_ERR_SOLVER;
  }
  if (cudaMalloc(&s->d_info, sizeof(int) * FM_SOLVER_INFO_SLOTS) != cudaSuccess ||
      cudaMemset(s->d_info, 0, sizeof(int) * FM_SOLVER_INFO_SLOTS) != cudaSuccess) {
    set_error("fm_solver_create: cudaMalloc(devInfo) failed");
    cusolverDnDestroyParams(s->params);
    cusolverDnDestroy(s->handle);
    delete s;
    return FM_ERR_CUDA;
  }
  *out = s;
  return FM_OK;
}

extern "C" int fm_solver_destroy(fm_solver* s) {
  if (!s) return FM_OK;
  if (s->d_work) cudaFree(s->d_work);
  if (s->d_f32) cudaFree(s->d_f32);
  if (s->d_info) cudaFree(s->d_info);
  free(s->h_work);
  if (s->params) cusolverDnDestroyParams(s->params);
  if (s->handle) cusolverDnDestroy(s->handle);
  delete s;
  return FM_OK;
}

extern "C" int fm_syevd(fm_solver* s, double* G, int n, double* lam, int use_fp32, void* stream) {
  using namespace fm;
  FM_REQUIRE(s && G && lam && n >= 1, "fm_syevd: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FM_SOLVER_TRY(cusolverDnSetStream(s->handle, st));
  const size_t nn = (size_t)n * n;
  void* A = G;
  void* W = lam;
  cudaDataType dt = CUDA_R_64F;
  if (use_fp32) {
    const size_t need = nn + (size_t)n;
    if (need > s->d_f32_elems) {
      if (s->d_f32) cudaFree(s->d_f32);
      s->d_f32 = nullptr;
      s->d_f32_elems = 0;
      FM_CUDA_TRY(cudaMalloc(&s->d_f32, need * sizeof(float)));
      s->d_f32_elems = need;
    }
    f64_to_f32_kernel<<<(unsigned)cdiv((int64_t)nn, 256), 256, 0, st>>>(G, s->d_f32, nn);
    FM_LAUNCH_CHECK();
    A = s->d_f32;
    W = s->d_f32 + nn;
    dt = CUDA_R_32F;
  }
  size_t dbytes = 0, hbytes = 0;
  // G is symmetric and fully populated, so row-major == column-major; eigenvector j comes back as
  // column j of the column-major result, i.e. ROW j of the row-major array.
  FM_SOLVER_TRY(cusolverDnXsyevd_bufferSize(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                            n, dt, A, n, dt, W, dt, &dbytes, &hbytes));
  int rc = ensure_workspace(s, dbytes, hbytes);
  if (rc != FM_OK) return rc;
  int* d_info = nullptr;
  rc = take_info_slots(s, 1, n, st, &d_info);
  if (rc != FM_OK) return rc;
  FM_SOLVER_TRY(cusolverDnXsyevd(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dt, A, n,
                                 dt, W, dt, s->d_work, dbytes, s->h_work, hbytes, d_info));
  if (use_fp32) {
    f32_to_f64_kernel<<<(unsigned)cdiv((int64_t)nn, 256), 256, 0, st>>>(s->d_f32, G, nn);
    FM_LAUNCH_CHECK();
    f32_to_f64_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(s->d_f32 + nn, lam, (size_t)n);
    FM_LAUNCH_CHECK();
  }
  return FM_OK;
}

// Batched variant (cusolverDnXsyevBatched): `batch` symmetric n x n float64 matrices stored back to
// back; each is overwritten by its eigenvectors (rows), lam is [batch, n].
extern "C" int fm_syevd_batched(fm_solver* s, double* G, int n, int batch, double* lam, void* stream) {
  using namespace fm;
  FM_REQUIRE(s && G && lam && n >= 1 && batch >= 1, "fm_syevd_batched: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FM_SOLVER_TRY(cusolverDnSetStream(s->handle, st));
  size_t dbytes = 0, hbytes = 0;
  FM_SOLVER_TRY(cusolverDnXsyevBatched_bufferSize(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                                  n, CUDA_R_64F, G, n, CUDA_R_64F, lam, CUDA_R_64F, &dbytes, &hbytes,
                                                  batch));
  int rc = ensure_workspace(s, dbytes, hbytes);
  if (rc != FM_OK) return rc;
  int* info = nullptr;
  rc = take_info_slots(s, batch, n, st, &info);
  if (rc != FM_OK) return rc;
  FM_SOLVER_TRY(cusolverDnXsyevBatched(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                       CUDA_R_64F, G, n, CUDA_R_64F, lam, CUDA_R_64F, s->d_work, dbytes, s->h_work,
                                       hbytes, info, batch));
  return FM_OK;
}

// Top-`k` eigenpairs only (cusolverDnXsyevdx, index range [n-k+1, n]): the first k rows of G
// receive the eigenvectors of the k largest eigenvalues in ascending order, lam[0..k).
extern "C" int fm_syevdx_top(fm_solver* s, double* G, int n, int k, double* lam, void* stream) {
  using namespace fm;
  FM_REQUIRE(s && G && lam && n >= 1 && k >= 1 && k <= n, "fm_syevdx_top: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FM_SOLVER_TRY(cusolverDnSetStream(s->handle, st));
  size_t dbytes = 0, hbytes = 0;
  int64_t meig = 0;
  double vl = 0.0, vu = 0.0;
  FM_SOLVER_TRY(cusolverDnXsyevdx_bufferSize(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                             CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, G, n, &vl, &vu, n - k + 1, n, &meig,
                                             CUDA_R_64F, lam, CUDA_R_64F, &dbytes, &hbytes));
  int rc = ensure_workspace(s, dbytes, hbytes);
  if (rc != FM_OK) return rc;
  int* d_info = nullptr;
  rc = take_info_slots(s, 1, n, st, &d_info);
  if (rc != FM_OK) return rc;
  FM_SOLVER_TRY(cusolverDnXsyevdx(s->handle, s->params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                  CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, G, n, &vl, &vu, n - k + 1, n, &meig, CUDA_R_64F,
                                  lam, CUDA_R_64F, s->d_work, dbytes, s->h_work, hbytes, d_info));
  if (meig != k) {       // an index range: the count is known to the host when the call returns
    set_error("cusolverDnXsyevdx: found %lld of %d eigenpairs (n = %d)", (long long)meig, k, n);
    return FM_ERR_SOLVER;
  }
  return FM_OK;
}
