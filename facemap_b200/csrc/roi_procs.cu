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
// One thread block per frame, persistent over the frames.  The fit is a chain of short dependent passes over
// the dark pixels (threshold, float32 weight sum, float64 moments, x6), so what matters is the instruction
// count and latency of each pass: the compacted pixel list lives in shared memory as ONE 32-bit word per
// entry (y, x, pixel value, "weight is zero" bit; entries beyond POS_CAP spill to a global scratch and take
// a slower instantiation), weights are recomputed from the packed pixel value, the first and second moments
// come from one fused pass, and the frame is read once with 8 pixels per thread in flight.
constexpr int PUPIL_NT = 256;
constexpr int POS_CAP = 8192;       // dark pixels kept in shared memory (a 100 px wide pupil); more spill to global
constexpr int LEAF_CAP = 512;       // pairwise leaves hold 64..128 elements: 32768 weights, thread-0 serial beyond
constexpr int PX_PER_THREAD = 8;
constexpr int PUPIL_MAX_DIM = 2047; // 11 bits per coordinate in the packed entry
constexpr unsigned DEAD_BIT = 0x80000000u;

__device__ __forceinline__ unsigned pack_entry(int y, int x, unsigned pix) {
  return (unsigned)y | ((unsigned)x << 11) | (pix << 22);
}
__device__ __forceinline__ int entry_y(unsigned e) { return (int)(e & 0x7ffu); }
__device__ __forceinline__ int entry_x(unsigned e) { return (int)((e >> 11) & 0x7ffu); }
__device__ __forceinline__ unsigned entry_pix(unsigned e) { return (e >> 22) & 0xffu; }

struct PupilShared {
  unsigned ent[POS_CAP];
  int leaf_lo[LEAF_CAP];
  int leaf_len[LEAF_CAP];
  float leaf_sum[LEAF_CAP];
  unsigned char leaf_comb[LEAF_CAP];
  int plan_stack[96];
  float fold_stack[32];
  int plan_n;                       // the weight count the current leaf plan was made for (-1: none)
  unsigned hist[257];
  double red[8 * (PUPIL_NT / 32)];
  int warp_tot[PUPIL_NT / 32];
  int count;
  float total;
  float med;
};

// the compacted list: entry r < n0 is a dark pixel, n0 <= r < n0 + nm a reflector pixel
template <bool SPILL>
struct PupilList {
  PupilShared* sh;
  unsigned* gent;
  float c32;      // float32(255 - saturation)
  float fill;     // weight of a reflector pixel inside the fit this round
  int n0;
  __device__ __forceinline__ unsigned get(int r) const {
    if (SPILL && r >= POS_CAP) return gent[r];
    return sh->ent[r];
  }
  __device__ __forceinline__ void set(int r, unsigned e) const {
    if (SPILL && r >= POS_CAP) gent[r] = e; else sh->ent[r] = e;
  }
  __device__ __forceinline__ float darkness(unsigned pix) const {
    return fmaxf(0.f, __fsub_rn((float)(255 - (int)pix), c32));
  }
  // the float32 weight before normalisation (pupil.py's lam0 / lamm)
  __device__ __forceinline__ float weight_of(unsigned e, int r) const {
    if (e & DEAD_BIT) return 0.f;
    return r < n0 ? darkness(entry_pix(e)) : fill;
  }
  __device__ __forceinline__ float weight(int r) const { return weight_of(get(r), r); }
};

// NumPy's float32 pairwise sum of the weights [0, n) by the whole block: thread 0 lists the leaves of the
// split tree as a flat plan (once per distinct n), groups of 8 lanes sum one leaf each (lane j = NumPy's
// accumulator r[j], combined in NumPy's order), thread 0 folds the leaf sums by the plan.  Same bits as
// numpy.sum (tests/test_numpy_sum.py pins plan + fold on the CPU).  Every thread returns the sum.
template <bool SPILL>
struct WeightLeaf {
  const PupilList<SPILL>* L;
  __device__ float operator()(int lo, int len) const {      // serial form (more than LEAF_CAP leaves)
    if (len < 8) {
      float r = -0.0f;
      for (int i = 0; i < len; ++i) r = __fadd_rn(r, L->weight(lo + i));
      return r;
    }
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = L->weight(lo + j);
    const int m = len - (len % 8);
    for (int i = 8; i < m; i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], L->weight(lo + i + j));
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (int i = m; i < len; ++i) res = __fadd_rn(res, L->weight(lo + i));
    return res;
  }
};

template <bool SPILL>
__device__ float block_weight_sum(const PupilList<SPILL>& L, int n) {
  PupilShared& sh = *L.sh;
  if (threadIdx.x == 0 && sh.plan_n != n) {       // the plan depends on n only: at most two per frame
    sh.count = pairwise_plan(n, sh.leaf_lo, sh.leaf_len, sh.leaf_comb, LEAF_CAP, sh.plan_stack);
    sh.plan_n = n;
  }
  __syncthreads();
  const int count = sh.count;
  if (count > LEAF_CAP) {
    if (threadIdx.x == 0) sh.total = __fadd_rn(0.0f, pairwise_walk(n, WeightLeaf<SPILL>{&L}));
  } else {
    const int group = threadIdx.x >> 3, j = threadIdx.x & 7;
    for (int l0 = 0; l0 < count; l0 += PUPIL_NT / 8) {       // uniform trip count: the shuffles stay convergent
      const int l = l0 + group;
      const int len = l < count ? sh.leaf_len[l] : 0;
      const int lo = l < count ? sh.leaf_lo[l] : 0;
      const int m = len >= 8 ? len - (len % 8) : 0;
      float r = len >= 8 ? L.weight(lo + j) : 0.f;
      for (int i = 8; i < m; i += 8) r = __fadd_rn(r, L.weight(lo + i + j));
      __syncwarp();
      float s1 = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));       // r0+r1, r2+r3, r4+r5, r6+r7
      s1 = __fadd_rn(s1, __shfl_down_sync(0xffffffffu, s1, 2, 8));           // (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
      s1 = __fadd_rn(s1, __shfl_down_sync(0xffffffffu, s1, 4, 8));
      if (j == 0 && l < count) {
        if (len < 8) s1 = -0.0f;
        for (int i = m; i < len; ++i) s1 = __fadd_rn(s1, L.weight(lo + i));
        sh.leaf_sum[l] = s1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
      sh.total = __fadd_rn(0.0f, pairwise_fold_plan(count, sh.leaf_sum, sh.leaf_comb, sh.fold_stack));
  }
  __syncthreads();
  const float total = sh.total;
  __syncthreads();
  return total;
}

struct PupilMoments {
  double mu0, mu1, s00, s01, s11;
};

// weights lam = weight / S (float32); mean = sum lam * (y, x); scatter = sum s^2 (d d^T) + 0.01 I with
// s = float32 sqrt(lam) (pupil.py:26-32).  One pass: the second moments are accumulated about the ROI centre
// (cy, cx) and moved to the mean afterwards, which costs ~1e-14 relative to the reference's two passes.
template <bool SPILL>
__device__ __forceinline__ PupilMoments pupil_moments(const PupilList<SPILL>& L, int n, float S, double cy, double cx) {
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int r = threadIdx.x; r < n; r += PUPIL_NT) {
    const unsigned e = L.get(r);
    const float lam = __fdiv_rn(L.weight_of(e, r), S);
    const double s = (double)__fsqrt_rn(lam), s2 = s * s;
    const double y = (double)entry_y(e), x = (double)entry_x(e);
    const double yc = y - cy, xc = x - cx;
    a[0] += (double)lam * y;
    a[1] += (double)lam * x;
    a[2] += s2;
    a[3] += s2 * yc;
    a[4] += s2 * xc;
    a[5] += s2 * yc * yc;
    a[6] += s2 * yc * xc;
    a[7] += s2 * xc * xc;
  }
  block_sum<8, PUPIL_NT>(a, L.sh->red);
  PupilMoments m;
  m.mu0 = a[0];
  m.mu1 = a[1];
  const double e0 = m.mu0 - cy, e1 = m.mu1 - cx;
  m.s00 = a[5] - 2.0 * e0 * a[3] + e0 * e0 * a[2] + 0.01;
  m.s01 = a[6] - e0 * a[4] - e1 * a[3] + e0 * e1 * a[2];
  m.s11 = a[7] - 2.0 * e1 * a[4] + e1 * e1 * a[2] + 0.01;
  return m;
}

// the fit proper on a compacted list (pupil.py:21-92); writes the 9 outputs of frame t
template <bool SPILL>
__device__ void pupil_fit(PupilList<SPILL>& L, int n0, int nm, bool have_miss, double sigma, double cy, double cx,
                          double* o) {
  PupilShared& sh = *L.sh;
  const double thr2 = 2.0 * sigma * sigma, thr1 = sigma * sigma, thrm = 1.15 * sigma * sigma;
  L.n0 = n0;
  L.fill = 0.f;
  // ---- initial weights: the dark pixels only (pupil.py:21-32) ----
  PupilMoments m = pupil_moments(L, n0, block_weight_sum(L, n0), cy, cx);
  const int ntot = n0 + nm;
  // ---- five refinement rounds (pupil.py:33-66) ----
  for (int it = 0; it < 5; ++it) {
    double i00, i01, i10, i11;
    inv2x2(m.s00, m.s01, m.s11, i00, i01, i10, i11);
    if (have_miss) {
      for (int h = threadIdx.x; h < 257; h += PUPIL_NT) sh.hist[h] = 0;
      __syncthreads();
    }
    for (int r = threadIdx.x; r < n0; r += PUPIL_NT) {
      unsigned e = L.get(r);
      const double dy = (double)entry_y(e) - m.mu0, dx = (double)entry_x(e) - m.mu1;
      const double dd = (dy * i00 + dx * i10) * dy + (dy * i01 + dx * i11) * dx;
      if (!(e & DEAD_BIT) && !(dd <= thr2)) {
        e |= DEAD_BIT;
        L.set(r, e);
      }
      if (have_miss && dd <= thr1) atomicAdd(&sh.hist[(e & DEAD_BIT) ? 256u : entry_pix(e)], 1u);
    }
    __syncthreads();
    if (have_miss) {
      if (threadIdx.x == 0) sh.med = median_from_hist(sh.hist, L.c32);
      __syncthreads();
      L.fill = __fmul_rn(sh.med, 1.1f);
      for (int q = threadIdx.x; q < nm; q += PUPIL_NT) {
        const unsigned e = L.get(n0 + q) & ~DEAD_BIT;
        const double dy = (double)entry_y(e) - m.mu0, dx = (double)entry_x(e) - m.mu1;
        const double dd = (dy * i00 + dx * i10) * dy + (dy * i01 + dx * i11) * dx;
        L.set(n0 + q, dd <= thrm ? e : (e | DEAD_BIT));
      }
      __syncthreads();
    }
    m = pupil_moments(L, ntot, block_weight_sum(L, ntot), cy, cx);
  }
  // ---- axes: eigen-decomposition of sigma^2 * scatter in LAPACK's convention (pupil.py:70-78) ----
  if (threadIdx.x == 0) {
    const double a = sigma * sigma * m.s00, b = sigma * sigma * m.s01, c = sigma * sigma * m.s11;
    if (!(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(m.mu0) && isfinite(m.mu1))) {
      const double NaN = __longlong_as_double(0x7ff8000000000000LL);
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
}

__global__ void __launch_bounds__(PUPIL_NT) pupil_kernel(
    const uint8_t* __restrict__ frames, int T, int H, int W, int y0, int x0, int ny, int nx,
    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ miss, float c32, double sigma,
    unsigned* scratch_ent, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char pupil_smem[];
  PupilShared& sh = *reinterpret_cast<PupilShared*>(pupil_smem);
  const int npx = ny * nx;
  PupilList<true> L;                // compaction does not know the count yet: spill-aware stores
  L.sh = &sh;
  L.gent = scratch_ent + (size_t)blockIdx.x * npx;
  L.c32 = c32;
  L.fill = 0.f;
  L.n0 = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double cy = 0.5 * ny, cx = 0.5 * nx;

  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const uint8_t* fr = frames + (size_t)t * H * W;
    // ---- ordered compaction: dark non-reflector pixels first (row-major), then the reflector pixels ----
    int n0 = 0, nm = 0;
    for (int pass = 0; pass < (miss ? 2 : 1); ++pass) {
      int base = pass == 0 ? 0 : n0;
      for (int tile = 0; tile < npx; tile += PUPIL_NT * PX_PER_THREAD) {
        const int i0 = tile + threadIdx.x * PX_PER_THREAD;
        unsigned pk[PX_PER_THREAD];
        unsigned takebits = 0;
        int y = i0 / nx, x = i0 - y * nx;
#pragma unroll
        for (int k = 0; k < PX_PER_THREAD; ++k) {
          pk[k] = 0;
          if (i0 + k < npx) {
            const unsigned pix = mask[i0 + k] ? 255u : (unsigned)fr[(size_t)(y0 + y) * W + x0 + x];
            const bool missing = miss && miss[i0 + k];
            const bool take = pass == 0 ? (L.darkness(pix) > 0.f && !missing) : missing;
            takebits |= (take ? 1u : 0u) << k;
            pk[k] = pack_entry(y, x, pix);
          }
          if (++x == nx) {
            x = 0;
            ++y;
          }
        }
        const int cnt = __popc(takebits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        if (lane == 31) sh.warp_tot[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < PUPIL_NT / 32; ++w) {
          const int c = sh.warp_tot[w];
          before += w < warp ? c : 0;
          total += c;
        }
        int r = base + before + incl - cnt;
#pragma unroll
        for (int k = 0; k < PX_PER_THREAD; ++k) {
          if (takebits & (1u << k)) L.set(r++, pk[k]);
        }
        base += total;
        __syncthreads();
      }
      if (pass == 0) n0 = base; else nm = base - n0;
    }
    if (threadIdx.x == 0) sh.plan_n = -1;
    __syncthreads();
    if (n0 + nm <= POS_CAP) {           // block-uniform: the whole list sits in shared memory
      PupilList<false> F;
      F.sh = &sh;
      F.gent = nullptr;
      F.c32 = c32;
      pupil_fit<false>(F, n0, nm, miss != nullptr, sigma, cy, cx, out + (size_t)t * 9);
    } else {
      pupil_fit<true>(L, n0, nm, miss != nullptr, sigma, cy, cx, out + (size_t)t * 9);
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
  *bytes = (size_t)blocks * ny * nx * 4;
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
  FM_REQUIRE(ny <= PUPIL_MAX_DIM && nx <= PUPIL_MAX_DIM, "fm_pupil: ROI %dx%d exceeds %d pixels per side", ny, nx, PUPIL_MAX_DIM);
  // per device and cheap: set on every launch rather than caching a flag that a second GPU would not see
  FM_CUDA_TRY(cudaFuncSetAttribute(pupil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PupilShared)));
  const int grid = T < blocks ? T : blocks;
  pupil_kernel<<<grid, PUPIL_NT, sizeof(PupilShared), reinterpret_cast<cudaStream_t>(stream)>>>(
      frames, T, H, W, y0, x0, ny, nx, mask, reflector, dark_offset, sigma, reinterpret_cast<unsigned*>(scratch), out);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
