// Full-resolution ROI processors that ride on stage 3's frame sweep:
//   fm_paint          frames | ormask            (process.py:543, 567 paint the chunk buffer in place)
//   fm_blink          per-frame sum of a 256-entry term table over a painted ROI   (process.py:557-572)
//   fm_running        first argmax of the sign-agreement map on the four corner blocks (running.py:91-133
//                     as executed: its fft2/ifft2 results are dropped, so no transform takes place)
//   fm_pupil          iterative Gaussian fit of the dark pixels of a painted ROI   (pupil.py:8-132)
// All of them read the raw uint8 frames once; the work per frame is a few ROI-sized passes that stay in
// L1/L2, so they are bound by the frame read (HBM) or, for the pupil fit, by float64 latency chains.
//
// Arithmetic that decides results bit-for-bit follows NumPy exactly: float32 sums use NumPy's
// pairwise order (8 strided accumulators per <=128-element leaf, halves rounded down to multiples of
// 8), float32 divisions / square roots are the correctly rounded ones, comparisons treat NaN like
// `~(x <= thr)`.
#include <math.h>
#include "common.cuh"
#include "numpy_sum.cuh"
#include "pupil_math.cuh"

namespace fm {

// block-wide float64 sums of NV values, result broadcast to every thread (fixed order: deterministic)
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /* [NV * NT/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[k * (NT / 32) + warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) x += red[k * (NT / 32) + w];
    v[k] = x;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// paint
// ------------------------------------------------------------------------------------------------
__global__ void paint_vec_kernel(const uint4* __restrict__ in, const uint4* __restrict__ mask, uint4* __restrict__ out,
                                 int T, int words) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const uint4 m = mask[w];
  for (int t = blockIdx.y; t < T; t += gridDim.y) {
    uint4 v = in[(size_t)t * words + w];
    v.x |= m.x;
    v.y |= m.y;
    v.z |= m.z;
    v.w |= m.w;
    out[(size_t)t * words + w] = v;
  }
}

__global__ void paint_byte_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ mask,
                                  uint8_t* __restrict__ out, int T, int hw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const uint8_t m = mask[i];
  for (int t = blockIdx.y; t < T; t += gridDim.y) out[(size_t)t * hw + i] = in[(size_t)t * hw + i] | m;
}

// ------------------------------------------------------------------------------------------------
// blink: histogram of the painted ROI, then sum_h count[h] * term[h] in a fixed order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blink_kernel(const uint8_t* __restrict__ frames, int T, int H, int W, int y0,
                                                    int x0, int ny, int nx, const uint8_t* __restrict__ mask,
                                                    const double* __restrict__ term, double* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ double part[256];
  const int t = blockIdx.x;
  const uint8_t* fr = frames + (size_t)t * H * W;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int npx = ny * nx;
  for (int i = threadIdx.x; i < npx; i += 256) {
    const int y = i / nx, x = i - y * nx;
    const unsigned pix = mask[i] ? 255u : (unsigned)fr[(size_t)(y0 + y) * W + x0 + x];
    atomicAdd(&hist[pix], 1u);
  }
  __syncthreads();
  part[threadIdx.x] = (double)hist[threadIdx.x] * term[threadIdx.x];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[t] = part[0];
}

// ------------------------------------------------------------------------------------------------
// running
// ------------------------------------------------------------------------------------------------
// mean over the painted ROI the way NumPy's X.mean(-1).mean(-1) rounds it in float32: exact integer
// row sums / nx, pairwise sum of the row means / ny.
__global__ void __launch_bounds__(256) running_mean_kernel(const uint8_t* __restrict__ frames, int T, int H, int W,
                                                           int y0, int x0, int ny, int nx,
                                                           const uint8_t* __restrict__ mask, float* __restrict__ means) {
  extern __shared__ float rowmean[];
  const int t = blockIdx.x;
  const uint8_t* fr = frames + (size_t)t * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int y = warp; y < ny; y += 8) {
    unsigned s = 0;
    const uint8_t* row = fr + (size_t)(y0 + y) * W + x0;
    const uint8_t* mrow = mask + (size_t)y * nx;
    for (int x = lane; x < nx; x += 32) s += (unsigned)(row[x] | mrow[x]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rowmean[y] = __fdiv_rn((float)s, (float)nx);
  }
  __syncthreads();
  if (threadIdx.x == 0) means[t] = __fdiv_rn(pairwise_sum_f32(rowmean, ny), (float)ny);
}

__device__ __forceinline__ float sign_of_centered(unsigned pix, float mean) {
  const float d = __fsub_rn((float)pix, mean);
  return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
}

__global__ void __launch_bounds__(256) running_argmax_kernel(
    const uint8_t* __restrict__ frames, int T, int H, int W, int y0, int x0, int ny, int nx,
    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ prev_crop, const float* __restrict__ prev_mean,
    const float* __restrict__ means, const float* __restrict__ g2, int lcorr, int* __restrict__ out) {
  __shared__ float bval[256];
  __shared__ int bidx[256];
  const int t = blockIdx.x;
  if (t == 0 && prev_crop == nullptr) return;
  const uint8_t* cur = frames + (size_t)t * H * W;
  const uint8_t* prv = t > 0 ? frames + (size_t)(t - 1) * H * W : nullptr;
  const float mc = means[t];
  const float mp = t > 0 ? means[t - 1] : prev_mean[0];
  const int w2 = 2 * lcorr + 1;
  float best = 0.f;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < w2 * w2; i += 256) {
    const int r = i / w2, c = i - r * w2;
    const int y = r < lcorr ? ny - lcorr + r : r - lcorr;
    const int x = c < lcorr ? nx - lcorr + c : c - lcorr;
    const unsigned m = mask[(size_t)y * nx + x];
    const unsigned pc = (unsigned)cur[(size_t)(y0 + y) * W + x0 + x] | m;
    const unsigned pp = prv ? ((unsigned)prv[(size_t)(y0 + y) * W + x0 + x] | m) : (unsigned)prev_crop[(size_t)y * nx + x];
    const float v = sign_of_centered(pc, mc) * sign_of_centered(pp, mp) * g2[(size_t)y * nx + x];
    if (besti == 0x7fffffff || v > best) {
      best = v;
      besti = i;
    }
  }
  bval[threadIdx.x] = best;
  bidx[threadIdx.x] = besti;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const float ov = bval[threadIdx.x + s];
      const int oi = bidx[threadIdx.x + s];
      const int mi = bidx[threadIdx.x];
      if (oi != 0x7fffffff && (mi == 0x7fffffff || ov > bval[threadIdx.x] || (ov == bval[threadIdx.x] && oi < mi))) {
        bval[threadIdx.x] = ov;
        bidx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int i = bidx[0];
    out[2 * t + 0] = i / w2 - lcorr;
    out[2 * t + 1] = i % w2 - lcorr;
  }
}

// ------------------------------------------------------------------------------------------------
// pupil
// ------------------------------------------------------------------------------------------------
constexpr int PUPIL_NT = 256;

struct PupilMoments {
  double mu0, mu1, s00, s01, s11;
};

// weights = val / S (float32), mean = sum w * (y, x), scatter = sum (d sqrt(w)) (d sqrt(w))^T + 0.01 I
__device__ __forceinline__ PupilMoments pupil_moments(const float* val, const unsigned* pos, int n, float S,
                                                      double* red) {
  double a[2] = {0.0, 0.0};
  for (int r = threadIdx.x; r < n; r += PUPIL_NT) {
    const double lam = (double)__fdiv_rn(val[r], S);
    const unsigned p = pos[r];
    a[0] += lam * (double)(p & 0xfffu);
    a[1] += lam * (double)((p >> 12) & 0xfffu);
  }
  block_sum<2, PUPIL_NT>(a, red);
  double b[3] = {0.0, 0.0, 0.0};
  for (int r = threadIdx.x; r < n; r += PUPIL_NT) {
    const float lam = __fdiv_rn(val[r], S);
    const double s = (double)__fsqrt_rn(lam);
    const unsigned p = pos[r];
    const double wy = ((double)(p & 0xfffu) - a[0]) * s;
    const double wx = ((double)((p >> 12) & 0xfffu) - a[1]) * s;
    b[0] += wy * wy;
    b[1] += wy * wx;
    b[2] += wx * wx;
  }
  block_sum<3, PUPIL_NT>(b, red);
  PupilMoments m;
  m.mu0 = a[0];
  m.mu1 = a[1];
  m.s00 = b[0] + 0.01;
  m.s01 = b[1];
  m.s11 = b[2] + 0.01;
  return m;
}

__global__ void __launch_bounds__(PUPIL_NT) pupil_kernel(
    const uint8_t* __restrict__ frames, int T, int H, int W, int y0, int x0, int ny, int nx,
    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ miss, float c32, double sigma,
    float* scratch_val, unsigned* scratch_pos, double* __restrict__ out) {
  __shared__ double red[3 * (PUPIL_NT / 32)];
  __shared__ int warp_tot[PUPIL_NT / 32];
  __shared__ unsigned hist[257];
  __shared__ float sh_f[2];
  const int npx = ny * nx;
  float* val = scratch_val + (size_t)blockIdx.x * npx;
  unsigned* pos = scratch_pos + (size_t)blockIdx.x * npx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double thr2 = 2.0 * sigma * sigma, thr1 = sigma * sigma, thrm = 1.15 * sigma * sigma;
  const double NaN = __longlong_as_double(0x7ff8000000000000LL);

  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const uint8_t* fr = frames + (size_t)t * H * W;
    // ---- ordered compaction: dark non-reflector pixels first (row-major), then the reflector pixels ----
    int n0 = 0, nm = 0;
    for (int pass = 0; pass < (miss ? 2 : 1); ++pass) {
      int base = 0;
      for (int i0 = 0; i0 < npx; i0 += PUPIL_NT) {
        const int i = i0 + threadIdx.x;
        bool take = false;
        float v = 0.f;
        unsigned pk = 0;
        if (i < npx) {
          const int y = i / nx, x = i - y * nx;
          const unsigned pix = mask[i] ? 255u : (unsigned)fr[(size_t)(y0 + y) * W + x0 + x];
          v = fmaxf(0.f, __fsub_rn((float)(255 - (int)pix), c32));
          const bool missing = miss && miss[i];
          take = pass == 0 ? (v > 0.f && !missing) : missing;
          pk = (unsigned)y | ((unsigned)x << 12) | (pix << 24);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < PUPIL_NT / 32; ++w) {
          const int c = warp_tot[w];
          before += w < warp ? c : 0;
          total += c;
        }
        if (take) {
          const int r = (pass == 0 ? 0 : n0) + base + before + __popc(bal & ((1u << lane) - 1u));
          val[r] = pass == 0 ? v : 0.f;
          pos[r] = pk;
        }
        base += total;
        __syncthreads();
      }
      if (pass == 0) n0 = base; else nm = base;
    }
    // ---- initial weights: the dark pixels only (pupil.py:21-32) ----
    if (threadIdx.x == 0) sh_f[0] = pairwise_sum_f32(val, n0);
    __syncthreads();
    PupilMoments m = pupil_moments(val, pos, n0, sh_f[0], red);
    const int ntot = n0 + nm;
    // ---- five refinement rounds (pupil.py:33-66) ----
    for (int it = 0; it < 5; ++it) {
      double i00, i01, i10, i11;
      inv2x2(m.s00, m.s01, m.s11, i00, i01, i10, i11);
      if (miss) {
        for (int h = threadIdx.x; h < 257; h += PUPIL_NT) hist[h] = 0;
        __syncthreads();
      }
      for (int r = threadIdx.x; r < n0; r += PUPIL_NT) {
        const unsigned p = pos[r];
        const double dy = (double)(p & 0xfffu) - m.mu0, dx = (double)((p >> 12) & 0xfffu) - m.mu1;
        const double dd = (dy * i00 + dx * i10) * dy + (dy * i01 + dx * i11) * dx;
        float v = val[r];
        if (!(dd <= thr2)) {
          v = 0.f;
          val[r] = 0.f;
        }
        if (miss && dd <= thr1) atomicAdd(&hist[v == 0.f ? 256u : (p >> 24)], 1u);
      }
      __syncthreads();
      if (miss) {
        if (threadIdx.x == 0) {
          const float med = median_from_hist(hist, c32);
          sh_f[1] = med;
        }
        __syncthreads();
        const float fill = __fmul_rn(sh_f[1], 1.1f);
        for (int q = threadIdx.x; q < nm; q += PUPIL_NT) {
          const unsigned p = pos[n0 + q];
          const double dy = (double)(p & 0xfffu) - m.mu0, dx = (double)((p >> 12) & 0xfffu) - m.mu1;
          const double dd = (dy * i00 + dx * i10) * dy + (dy * i01 + dx * i11) * dx;
          val[n0 + q] = dd <= thrm ? fill : 0.f;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) sh_f[0] = pairwise_sum_f32(val, ntot);
      __syncthreads();
      m = pupil_moments(val, pos, ntot, sh_f[0], red);
    }
    // ---- axes: eigen-decomposition of sigma^2 * scatter in LAPACK's dlanv2 convention (pupil.py:70-78) ----
    if (threadIdx.x == 0) {
      double* o = out + (size_t)t * 9;
      const double a = sigma * sigma * m.s00, b = sigma * sigma * m.s01, c = sigma * sigma * m.s11;
      if (!(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(m.mu0) && isfinite(m.mu1))) {
        for (int k = 0; k < 9; ++k) o[k] = NaN;      // numpy.linalg.eig raises -> the frame stays NaN
      } else {
        double w0, w1, u00, u01, u10, u11;
        sym2x2_eig_lapack(a, b, c, w0, w1, u00, u01, u10, u11);
        const double lo = fmin(w0, w1);
        w0 = lo * fmin(4.0, w0 / lo);
        w1 = lo * fmin(4.0, w1 / lo);
        o[0] = m.mu0;
        o[1] = m.mu1;
        o[2] = 3.141592653589793 * sqrt(w1 * w0);
        o[3] = u01;      // axdir = eigenvectors with the columns reversed (pupil.py:77)
        o[4] = u00;
        o[5] = u11;
        o[6] = u10;
        o[7] = w1;       // axlen reversed as well
        o[8] = w0;
      }
    }
    __syncthreads();
  }
}

}  // namespace fm

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" int fm_paint(const uint8_t* frames, int T, int H, int W, const uint8_t* ormask, uint8_t* out,
                        void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && ormask && out && T >= 0 && H >= 1 && W >= 1, "fm_paint: bad args");
  if (T == 0) return FM_OK;
  const int64_t hw = (int64_t)H * W;
  FM_REQUIRE(hw < (int64_t)1 << 31, "fm_paint: frame too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = hw % 16 == 0 && ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(ormask) |
                                     reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const unsigned gy = (unsigned)(T < 4096 ? T : 4096);
  if (vec) {
    const int words = (int)(hw / 16);
    dim3 grid((unsigned)cdiv(words, 256), gy);
    paint_vec_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(frames), reinterpret_cast<const uint4*>(ormask),
                                           reinterpret_cast<uint4*>(out), T, words);
  } else {
    dim3 grid((unsigned)cdiv(hw, 256), gy);
    paint_byte_kernel<<<grid, 256, 0, st>>>(frames, ormask, out, T, (int)hw);
  }
  FM_LAUNCH_CHECK();
  return FM_OK;
}

static int check_rect(const char* who, int H, int W, int y0, int y1, int x0, int x1) {
  using namespace fm;
  FM_REQUIRE(y0 >= 0 && x0 >= 0 && y1 > y0 && x1 > x0 && y1 <= H && x1 <= W && y1 - y0 <= 4095 && x1 - x0 <= 4095,
             "%s: ROI rectangle [%d,%d)x[%d,%d) does not fit the %dx%d frame", who, y0, y1, x0, x1, H, W);
  return FM_OK;
}

extern "C" int fm_blink(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
                        const uint8_t* mask, const double* term, double* out, void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && mask && term && out && T >= 0, "fm_blink: bad args");
  if (int rc = check_rect("fm_blink", H, W, y0, y1, x0, x1)) return rc;
  if (T == 0) return FM_OK;
  blink_kernel<<<(unsigned)T, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(frames, T, H, W, y0, x0, y1 - y0,
                                                                                  x1 - x0, mask, term, out);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_running(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
                          const uint8_t* ormask, const uint8_t* prev_crop, const float* prev_mean, const float* g2,
                          float* means, int32_t* out, void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && ormask && g2 && means && out && T >= 0, "fm_running: bad args");
  FM_REQUIRE((prev_crop == nullptr) == (prev_mean == nullptr), "fm_running: prev_crop and prev_mean go together");
  if (int rc = check_rect("fm_running", H, W, y0, y1, x0, x1)) return rc;
  const int ny = y1 - y0, nx = x1 - x0;
  const int ly = (int)floor(ny * 0.4), lx = (int)floor(nx * 0.4);
  const int lcorr = ly < lx ? ly : lx;
  FM_REQUIRE(lcorr >= 1, "fm_running: ROI %dx%d is too small (the reference fails on it too)", ny, nx);
  if (T == 0) return FM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  running_mean_kernel<<<(unsigned)T, 256, ny * sizeof(float), st>>>(frames, T, H, W, y0, x0, ny, nx, ormask, means);
  FM_LAUNCH_CHECK();
  running_argmax_kernel<<<(unsigned)T, 256, 0, st>>>(frames, T, H, W, y0, x0, ny, nx, ormask, prev_crop, prev_mean,
                                                     means, g2, lcorr, out);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_pupil_scratch_bytes(int ny, int nx, int blocks, size_t* bytes) {
  using namespace fm;
  FM_REQUIRE(ny >= 1 && nx >= 1 && blocks >= 1 && bytes, "fm_pupil_scratch_bytes: bad args");
  *bytes = (size_t)blocks * ny * nx * 8;
  return FM_OK;
}

extern "C" int fm_pupil(const uint8_t* frames, int T, int H, int W, int y0, int y1, int x0, int x1,
                        const uint8_t* mask, const uint8_t* reflector, float dark_offset, double sigma,
                        void* scratch, int blocks, double* out, void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && mask && scratch && out && T >= 0 && blocks >= 1 && sigma > 0, "fm_pupil: bad args");
  if (int rc = check_rect("fm_pupil", H, W, y0, y1, x0, x1)) return rc;
  if (T == 0) return FM_OK;
  const int ny = y1 - y0, nx = x1 - x0;
  const size_t npx = (size_t)ny * nx;
  float* sval = reinterpret_cast<float*>(scratch);
  unsigned* spos = reinterpret_cast<unsigned*>(sval + npx * blocks);
  const int grid = T < blocks ? T : blocks;
  pupil_kernel<<<grid, PUPIL_NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      frames, T, H, W, y0, x0, ny, nx, mask, reflector, dark_offset, sigma, sval, spos, out);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
