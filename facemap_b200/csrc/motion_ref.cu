// fm_motion_ref — the motion energy `imbin_mot.sum(-1)` (facemap/process.py:460) evaluated with the
// reference's own floating-point arithmetic, for the cases where the exact integer energy of K1 cannot decide
// the reference's float32 result:
//   * sbin not a power of two: the binned frames are float64 means of means (process.py:38-39: sum of a raw
//     row segment / sbin, those summed row by row / sbin), their |differences| are summed over the pixels in
//     NumPy's pairwise order, and the per-view sums are accumulated into a float32 array.  With two or more
//     views ~3 % of the frames land exactly on a float32 rounding tie of the exact value, which the reference
//     breaks by the sign of its 1e-13 float64 rounding error.
//   * sbin == 1 with more than 65 793 pixels per view: the reference sums float32 values whose total exceeds
//     2^24, so its pairwise float32 rounding differs from the correctly rounded integer sum.
// One block per frame pair: |b_t - b_{t-1}| per binned pixel into a scratch row, leaf sums by 8-lane groups in
// NumPy's accumulator order, fold by the (host-made, per view) pairwise plan.  Costs one more read of the frames.
#include "common.cuh"
#include "numpy_sum.cuh"

namespace fm {

constexpr int MR_THREADS = 256;

template <class F>
struct MotionRefParams {
  const uint8_t* frames;
  const uint8_t* prev_frame;   // frame preceding frames[0] or NULL
  const uint8_t* ormask;       // [H, W] 0x00 / 0xFF or NULL
  int T, H, W, sbin, Lyb, Lxb, rows, frame_off;
  const int* leaf_lo;
  const int* leaf_len;
  const unsigned char* leaf_comb;
  int nleaves;
  F* scratch;                  // [grid, 2 * npix + nleaves]: |differences|, previous binned frame, leaf sums
  double* out;                 // [rows]
  // sbin == 1 only (motion_ref_u8_kernel<true>): the |difference| bytes as the stage-3 operand plane and the exact
  // integer energy — what K1 would produce in a pass of its own
  uint8_t* plane;              // [rows, ldq], written at columns [col0, col0 + H * W)
  int64_t ldq, col0;
  unsigned long long* energy;  // [rows], += sum of |differences|
  // fold plan over the exact subtrees (groups of leaves whose float32 sum cannot round), or ngroups == 0
  const int* grp_first;
  const int* grp_n;
  const unsigned char* grp_comb;
  int ngroups;
};

// the reference's binned value of one pixel block (process.py:34-43)
template <class F>
__device__ __forceinline__ F binned_value(const uint8_t* fr, const uint8_t* mask, int W, int sbin, int by, int bx) {
  const uint8_t* p = fr + (size_t)by * sbin * W + (size_t)bx * sbin;
  const uint8_t* m = mask ? mask + (size_t)by * sbin * W + (size_t)bx * sbin : nullptr;
  if constexpr (sizeof(F) == 4) {        // sbin == 1: imbin = im.astype(float32)
    return (F)(unsigned)(m ? (p[0] | m[0]) : p[0]);
  } else {
    const double s = (double)sbin;
    double acc = 0.0;
    for (int r = 0; r < sbin; ++r) {
      int rowsum = 0;
      for (int c = 0; c < sbin; ++c) rowsum += (int)(m ? (p[(size_t)r * W + c] | m[(size_t)r * W + c]) : p[(size_t)r * W + c]);
      const double mean_r = __ddiv_rn((double)rowsum, s);      // .mean(axis=-1)
      acc = r == 0 ? mean_r : __dadd_rn(acc, mean_r);          // .mean(axis=-2) adds the rows in order
    }
    return __ddiv_rn(acc, s);
  }
}

// The same for 4 horizontally adjacent binned pixels (bx0 a multiple of 4) of a 4-byte aligned frame: their
// 4*sbin raw bytes per row are `sbin` aligned words, each word feeds at most two pixels (sbin >= 3) through
// dp4a byte masks.  Words at or beyond the row end are skipped (pixels past Lxb come out as garbage: unused).
__device__ __forceinline__ void binned_group4(const uint8_t* fr, const uint8_t* mask, int W, int sbin, int by, int bx0,
                                              double (&out)[4]) {
  const size_t off = (size_t)by * sbin * W + (size_t)bx0 * sbin;
  const int wleft = (W - bx0 * sbin) / 4;
  const int nw = sbin < wleft ? sbin : wleft;
  const double s = (double)sbin;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < sbin; ++r) {
    const unsigned* w = reinterpret_cast<const unsigned*>(fr + off + (size_t)r * W);
    const unsigned* mw = mask ? reinterpret_cast<const unsigned*>(mask + off + (size_t)r * W) : nullptr;
    int rs[4] = {0, 0, 0, 0};
    for (int q = 0; q < nw; ++q) {
      unsigned word = __ldg(w + q);
      if (mw) word |= __ldg(mw + q);
      const int j0 = (4 * q) / sbin;                       // pixel of the word's first byte
      const int c0 = min(4, (j0 + 1) * sbin - 4 * q);      // how many of its 4 bytes belong to that pixel
      const unsigned lo_ones = 0x01010101u >> (8 * (4 - c0));
      const int v0 = (int)__dp4a(word, lo_ones, 0u);
      const int v1 = (int)__dp4a(word, 0x01010101u & ~lo_ones, 0u);     // the rest goes to the next pixel
#pragma unroll
      for (int j = 0; j < 4; ++j) rs[j] += (j == j0 ? v0 : 0) + (j == j0 + 1 ? v1 : 0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double mean_r = __ddiv_rn((double)rs[j], s);
      acc[j] = r == 0 ? mean_r : __dadd_rn(acc[j], mean_r);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = __ddiv_rn(acc[j], s);
}

// One block per run of consecutive rows: the binned values of a frame are kept (scratch `bprev`) for the next
// row, so every frame is binned once per block run instead of twice per row.
template <class F, bool VEC4>
__global__ void __launch_bounds__(MR_THREADS) motion_ref_kernel(const MotionRefParams<F> p) {
  const int npix = p.Lyb * p.Lxb;
  F* d = p.scratch + (size_t)blockIdx.x * (2 * (size_t)npix + p.nleaves);
  F* bprev = d + npix;
  F* leaf_sum = bprev + npix;
  const size_t hw = (size_t)p.H * p.W;
  const int per_block = (p.rows + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * per_block;
  const int r_end = min(p.rows, r_begin + per_block);
  for (int row = r_begin; row < r_end; ++row) {
    const int f = row + p.frame_off;
    const uint8_t* cur = p.frames + (size_t)f * hw;
    const uint8_t* prv = f > 0 ? p.frames + (size_t)(f - 1) * hw : p.prev_frame;
    const bool first = row == r_begin;
    // every thread owns the same pixels in every row of the run: bprev needs no synchronisation
    if constexpr (VEC4) {
      const int GX = (p.Lxb + 3) / 4;
      for (int g = threadIdx.x; g < p.Lyb * GX; g += MR_THREADS) {
        const int by = g / GX, bx0 = 4 * (g - by * GX);
        const int nvalid = min(4, p.Lxb - bx0);
        double a[4], b[4];
        binned_group4(cur, p.ormask, p.W, p.sbin, by, bx0, a);
        if (first) binned_group4(prv, p.ormask, p.W, p.sbin, by, bx0, b);
        const int i = by * p.Lxb + bx0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nvalid) {
            const double bj = first ? b[j] : (double)bprev[i + j];
            d[i + j] = (F)fabs(fadd_rn(a[j], -bj));
            bprev[i + j] = (F)a[j];
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < npix; i += MR_THREADS) {
        const int by = i / p.Lxb, bx = i - by * p.Lxb;
        const F a = binned_value<F>(cur, p.ormask, p.W, p.sbin, by, bx);
        const F b = first ? binned_value<F>(prv, p.ormask, p.W, p.sbin, by, bx) : bprev[i];
        d[i] = fabs(fadd_rn(a, -b));
        bprev[i] = a;
      }
    }
    __syncthreads();
    const int group = threadIdx.x >> 3, j = threadIdx.x & 7;
    for (int l0 = 0; l0 < p.nleaves; l0 += MR_THREADS / 8) {       // uniform trip count: the shuffles stay convergent
      const int l = l0 + group;
      const int len = l < p.nleaves ? p.leaf_len[l] : 0;
      const F* a = d + (l < p.nleaves ? p.leaf_lo[l] : 0);
      const int m = len >= 8 ? len - (len % 8) : 0;
      F r = len >= 8 ? a[j] : (F)0;
      for (int i = 8; i < m; i += 8) r = fadd_rn(r, a[i + j]);
      __syncwarp();
      F s1 = fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));       // r0+r1, r2+r3, r4+r5, r6+r7
      s1 = fadd_rn(s1, __shfl_down_sync(0xffffffffu, s1, 2, 8));       // (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
      s1 = fadd_rn(s1, __shfl_down_sync(0xffffffffu, s1, 4, 8));
      if (j == 0 && l < p.nleaves) {
        if (len < 8) s1 = (F)-0.0;
        for (int i = m; i < len; ++i) s1 = fadd_rn(s1, a[i]);
        leaf_sum[l] = s1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      F vstack[32];
      p.out[row] = (double)fadd_rn((F)0, pairwise_fold_plan<F>(p.nleaves, leaf_sum, p.leaf_comb, vstack));
    }
    __syncthreads();
  }
}

// sbin == 1: the binned frame IS the frame (float32 of a byte), |difference| is an integer <= 255 and every partial sum
// inside a leaf of numpy's pairwise tree (<= 128 elements: <= 32 640) is an integer below 2^24 — exact in float32 in
// ANY order.  So the leaves are summed as integers straight from the bytes (8-byte loads, SAD) and only the fold of
// the leaf sums, where the float32 roundings of the reference happen, follows the plan.  One block walks a run of
// consecutive rows; a half-warp owns a leaf; scratch = the leaf sums only.
template <bool PLANE>
__global__ void __launch_bounds__(MR_THREADS) motion_ref_u8_kernel(const MotionRefParams<float> p) {
  __shared__ unsigned long long s_energy;
  __shared__ float s_group[256];
  float* leaf_sum = p.scratch + (size_t)blockIdx.x * p.nleaves;
  const size_t hw = (size_t)p.H * p.W;
  const int per_block = (p.rows + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * per_block;
  const int r_end = min(p.rows, r_begin + per_block);
  const int group = threadIdx.x >> 4, j = threadIdx.x & 15;
  for (int row = r_begin; row < r_end; ++row) {
    const int f = row + p.frame_off;
    const uint8_t* cur = p.frames + (size_t)f * hw;
    const uint8_t* prv = f > 0 ? p.frames + (size_t)(f - 1) * hw : p.prev_frame;
    const bool al = ((reinterpret_cast<uintptr_t>(cur) | reinterpret_cast<uintptr_t>(prv) |
                      reinterpret_cast<uintptr_t>(p.ormask)) & 7) == 0;
    uint8_t* prow = PLANE ? p.plane + (size_t)row * p.ldq + p.col0 : nullptr;
    const bool pal = PLANE && ((reinterpret_cast<uintptr_t>(prow) & 7) == 0);
    unsigned long long esum = 0;
    if (PLANE && threadIdx.x == 0) s_energy = 0;
    constexpr int GROUPS = MR_THREADS / 16, UNR = 4;                // four leaves per half-warp in flight
    for (int l0 = 0; l0 < p.nleaves; l0 += GROUPS * UNR) {         // uniform trip count: the shuffles stay convergent
      unsigned sum[UNR];
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int l = l0 + q * GROUPS + group;
        const int len = l < p.nleaves ? p.leaf_len[l] : 0;
        const int lo = l < p.nleaves ? p.leaf_lo[l] : 0;
        sum[q] = 0;
        for (int w0 = 8 * j; w0 < len; w0 += 128) {                // one pass: leaves hold at most 128 elements
          const int nb = min(8, len - w0);
          if (nb == 8 && al && ((lo + w0) & 7) == 0) {
            uint2 c = __ldg(reinterpret_cast<const uint2*>(cur + lo + w0));
            uint2 qv = __ldg(reinterpret_cast<const uint2*>(prv + lo + w0));
            if (p.ormask) {
              const uint2 m = __ldg(reinterpret_cast<const uint2*>(p.ormask + lo + w0));
              c.x |= m.x; c.y |= m.y; qv.x |= m.x; qv.y |= m.y;
            }
            sum[q] = __vsadu4(c.x, qv.x) + __vsadu4(c.y, qv.y) + sum[q];
            if (PLANE) {
              const uint2 dv = make_uint2(__vabsdiffu4(c.x, qv.x), __vabsdiffu4(c.y, qv.y));
              if (pal) {
                *reinterpret_cast<uint2*>(prow + lo + w0) = dv;
              } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                  prow[lo + w0 + b] = (uint8_t)(dv.x >> (8 * b));
                  prow[lo + w0 + 4 + b] = (uint8_t)(dv.y >> (8 * b));
                }
              }
            }
          } else {
            for (int b = 0; b < nb; ++b) {
              int x = cur[lo + w0 + b], y = prv[lo + w0 + b];
              if (p.ormask) { x |= p.ormask[lo + w0 + b]; y |= p.ormask[lo + w0 + b]; }
              sum[q] += (unsigned)abs(x - y);
              if (PLANE) prow[lo + w0 + b] = (uint8_t)abs(x - y);
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        unsigned v = sum[q];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, 16);
        const int l = l0 + q * GROUPS + group;
        if (j == 0 && l < p.nleaves) {
          leaf_sum[l] = (float)v;
          esum += v;
        }
      }
    }
    if (PLANE && p.energy) {                                       // exact integer energy of the row (K1's)
      __syncthreads();
      if (esum) atomicAdd(&s_energy, esum);
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(p.energy + row, s_energy);
    }
    __syncthreads();
    if (p.ngroups > 0) {
      // the exact subtrees: integer sums of their leaves by whole warps, in any order
      for (int g = threadIdx.x >> 5; g < p.ngroups; g += MR_THREADS / 32) {
        const float* a = leaf_sum + p.grp_first[g];
        unsigned v = 0;
        for (int i = threadIdx.x & 31; i < p.grp_n[g]; i += 32) v += __float2uint_rn(a[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_group[g] = (float)v;              // < 2^24: exact
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      float vstack[32];
      const float r = p.ngroups > 0 ? pairwise_fold_plan<float>(p.ngroups, s_group, p.grp_comb, vstack)
                                    : pairwise_fold_plan<float>(p.nleaves, leaf_sum, p.leaf_comb, vstack);
      p.out[row] = (double)fadd_rn(0.0f, r);
    }
    __syncthreads();
  }
}

// Stage-1 segment means in the reference's arithmetic (process.py:77-87) for a non-power-of-two sbin: per binned
// pixel the float64 binned values of the segment's frames are added in frame order and divided by T
// (`imbin.mean(axis=0)` reduces the leading axis sequentially), likewise the T-1 absolute differences.
__global__ void __launch_bounds__(MR_THREADS) mean_ref_kernel(const uint8_t* frames, int T, int H, int W, int sbin,
                                                              int Lyb, int Lxb, double* mean_bin, double* mean_adiff) {
  const int i = blockIdx.x * MR_THREADS + threadIdx.x;
  if (i >= Lyb * Lxb) return;
  const int by = i / Lxb, bx = i - by * Lxb;
  const size_t hw = (size_t)H * W;
  double prev = binned_value<double>(frames, nullptr, W, sbin, by, bx);
  double accb = prev, accd = 0.0;
  for (int t = 1; t < T; ++t) {
    const double b = binned_value<double>(frames + (size_t)t * hw, nullptr, W, sbin, by, bx);
    accb = __dadd_rn(accb, b);
    const double d = fabs(__dadd_rn(b, -prev));
    accd = t == 1 ? d : __dadd_rn(accd, d);
    prev = b;
  }
  mean_bin[i] = __ddiv_rn(accb, (double)T);
  mean_adiff[i] = __ddiv_rn(accd, (double)(T - 1));       // T == 1: 0/0 = NaN, like the mean of an empty array
}

}  // namespace fm

extern "C" int fm_mean_ref(const uint8_t* frames, int T, int H, int W, int sbin, double* mean_bin, double* mean_adiff,
                           void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && mean_bin && mean_adiff && T >= 1 && H >= 1 && W >= 1 && sbin >= 2, "fm_mean_ref: bad args");
  const int Lyb = H / sbin, Lxb = W / sbin;
  FM_REQUIRE(Lyb >= 1 && Lxb >= 1, "fm_mean_ref: sbin %d larger than the %dx%d frame", sbin, H, W);
  mean_ref_kernel<<<(unsigned)cdiv((int64_t)Lyb * Lxb, MR_THREADS), MR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      frames, T, H, W, sbin, Lyb, Lxb, mean_bin, mean_adiff);
  FM_LAUNCH_CHECK();
  return FM_OK;
}

// the leaf plan of numpy's pairwise sum over n elements (host arrays of capacity `cap`); returns the leaf count
extern "C" int fm_pairwise_plan(int n, int cap, int32_t* lo, int32_t* len, uint8_t* comb) {
  int st[96];
  if (n < 0 || cap < 1 || !lo || !len || !comb) return -1;
  return fm::pairwise_plan(n, lo, len, comb, cap, st);
}

extern "C" int fm_motion_ref_scratch_bytes(int H, int W, int sbin, int nleaves, int blocks, size_t* bytes) {
  using namespace fm;
  FM_REQUIRE(H >= 1 && W >= 1 && sbin >= 1 && nleaves >= 1 && blocks >= 1 && bytes, "fm_motion_ref_scratch_bytes: bad args");
  // sbin 1 keeps only the leaf sums (motion_ref_u8_kernel); other bins also two rows of binned values
  *bytes = sbin == 1 ? (size_t)blocks * nleaves * 4 : (size_t)blocks * (2 * (size_t)(H / sbin) * (W / sbin) + nleaves) * 8;
  return FM_OK;
}

extern "C" int fm_motion_ref(const uint8_t* frames, int T, int H, int W, int sbin, const uint8_t* prev_frame,
                             const uint8_t* ormask, const int32_t* leaf_lo, const int32_t* leaf_len,
                             const uint8_t* leaf_comb, int nleaves, void* scratch, int blocks, double* out,
                             void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && leaf_lo && leaf_len && leaf_comb && scratch && out && T >= 1 && H >= 1 && W >= 1 && sbin >= 1 &&
                 nleaves >= 1 && blocks >= 1, "fm_motion_ref: bad args");
  const int Lyb = H / sbin, Lxb = W / sbin;
  FM_REQUIRE(Lyb >= 1 && Lxb >= 1, "fm_motion_ref: sbin %d larger than the %dx%d frame", sbin, H, W);
  const int rows = prev_frame ? T : T - 1;
  if (rows == 0) return FM_OK;
  const int grid = rows < blocks ? rows : blocks;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (sbin == 1) {
    MotionRefParams<float> p{frames, prev_frame, ormask, T, H, W, sbin, Lyb, Lxb, rows, prev_frame ? 0 : 1,
                             leaf_lo, leaf_len, leaf_comb, nleaves, reinterpret_cast<float*>(scratch), out,
                             nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 0};
    motion_ref_u8_kernel<false><<<grid, MR_THREADS, 0, st>>>(p);
  } else {
    MotionRefParams<double> p{frames, prev_frame, ormask, T, H, W, sbin, Lyb, Lxb, rows, prev_frame ? 0 : 1,
                              leaf_lo, leaf_len, leaf_comb, nleaves, reinterpret_cast<double*>(scratch), out,
                              nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, 0};
    const bool aligned = sbin >= 3 && W % 4 == 0 &&
                         ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(prev_frame) |
                           reinterpret_cast<uintptr_t>(ormask)) % 4 == 0);
    if (aligned) motion_ref_kernel<double, true><<<grid, MR_THREADS, 0, st>>>(p);
    else motion_ref_kernel<double, false><<<grid, MR_THREADS, 0, st>>>(p);
  }
  FM_LAUNCH_CHECK();
  return FM_OK;
}


// sbin == 1, one pass for everything stage 3 needs of a view: the reference-arithmetic motion sums (as fm_motion_ref), the
// |difference| bytes as the projection's operand plane (what fm_bin_motion_planes writes at sbin 1) and the exact integer
// energy.  Replaces K1 + fm_motion_ref (two passes over the frames) on large full-resolution views (BASELINE configs[3]).
extern "C" int fm_motion_ref_planes(const uint8_t* frames, int T, int H, int W, const uint8_t* prev_frame,
                                    const uint8_t* ormask, const int32_t* leaf_lo, const int32_t* leaf_len,
                                    const uint8_t* leaf_comb, int nleaves, const int32_t* grp_first, const int32_t* grp_n,
                                    const uint8_t* grp_comb, int ngroups, void* scratch, int blocks, double* out,
                                    uint8_t* plane, int64_t ldq, int64_t col0, uint64_t* energy, void* stream) {
  using namespace fm;
  FM_REQUIRE(frames && leaf_lo && leaf_len && leaf_comb && scratch && out && plane && T >= 1 && H >= 1 && W >= 1 &&
                 nleaves >= 1 && blocks >= 1 && col0 >= 0 && ldq >= col0 + (int64_t)H * W,
             "fm_motion_ref_planes: bad args");
  const int rows = prev_frame ? T : T - 1;
  if (rows == 0) return FM_OK;
  const int grid = rows < blocks ? rows : blocks;
  MotionRefParams<float> p{frames, prev_frame, ormask, T, H, W, 1, H, W, rows, prev_frame ? 0 : 1,
                           leaf_lo, leaf_len, leaf_comb, nleaves, reinterpret_cast<float*>(scratch), out,
                           plane, ldq, col0, reinterpret_cast<unsigned long long*>(energy), nullptr, nullptr, nullptr, 0};
  if (ngroups > 0 && ngroups <= 256 && grp_first && grp_n && grp_comb) {
    p.grp_first = grp_first;
    p.grp_n = grp_n;
    p.grp_comb = grp_comb;
    p.ngroups = ngroups;
  }
  motion_ref_u8_kernel<true><<<grid, MR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
