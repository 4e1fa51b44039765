// K1 — fused spatial bin + frame difference + |.| + average subtract + motion energy + running
// sums over raw uint8 frames (HBM-bound; every frame byte is read once per slab).
//
// Restates, for one camera view (reference = /root/reference):
//   spatial_bin                      facemap/process.py:34-43
//   np.abs(np.diff(imbin)) - avg     facemap/process.py:201-212 (stage 2), :448-463 (stage 3)
//   imbin_mot.sum(-1)                facemap/process.py:460, :480
//   per-segment sums of stage 1      facemap/process.py:77-87
//
// Work decomposition: a thread owns PX horizontally adjacent binned pixels (PX*sbin = 16 raw
// bytes, one 128-bit load per raw row) and walks a slab of consecutive frames keeping the
// previous frame's integer block sums in registers, so the difference costs no second read.
// Blocks tile (pixel groups) x (time slabs); a slab re-reads one halo frame.
// All arithmetic up to the final float conversion is integer: block sums, |dS|, per-frame energy
// (warp redux -> shared atomics -> one 64-bit global atomic per row per block) and the stage-1
// time sums.  For power-of-two sbin the float results are bit-identical to the reference's
// float64 evaluation (single rounding via FMA); other sbin use float64 division (<= 1 ulp).
#include "common.cuh"

namespace fm {

constexpr int K1_THREADS = 256;
constexpr int K1_MAX_SLAB = 128;

struct K1Params {
  const uint8_t* frames;
  int T, H, W, sbin, Lyb, Lxb;
  const int32_t* prev_sum;
  const float* avgmotion;
  const float* avgframe;
  float* x_mot;
  float* x_mov;
  int64_t ldx, col0;
  unsigned long long* energy;
  unsigned long long* roi_energy;
  int32_t* last_sum;
  unsigned long long* tsum_bin;
  unsigned long long* tsum_adiff;
  int rows;       // T or T-1
  int frame_off;  // frame index of row 0
  int slab;       // rows per slab
  fm_rois rois;
  const uint8_t* ormask;  // optional [H, W] bytes OR-ed into every frame (0x00 / 0xFF: painted pixels read 255)  // optional integer operands for fm_project_i8: the motion energy |S_t - S_(t-1)| (mot) and the block sum S_t (mov)
  // as u8 planes value = 256 * hi + lo, rows like x_mot / x_mov, pitch ldq bytes (hi may be NULL when sbin == 1)
  uint8_t* q_mot_hi;
  uint8_t* q_mot_lo;
  uint8_t* q_mov_hi;
  uint8_t* q_mov_lo;
  int64_t ldq;
};

// Per-thread geometry of the vector kernel.  Power-of-two bins: 16 raw bytes per row (4 * sbin below sbin 4),
// i.e. 4, 4, 4, 2, 1 binned pixels for sbin 1, 2, 4, 8, 16.  Other bins: PX binned pixels with PX * sbin a
// multiple of 4, so a thread's row segment is PX * sbin / 4 aligned words (PX = 4, or 2 for large even bins).
constexpr bool k1_pow2(int sbin) { return (sbin & (sbin - 1)) == 0; }
constexpr int k1_px(int sbin) {
  return k1_pow2(sbin) ? ((sbin >= 4 ? 16 : 4 * sbin) / sbin) : ((sbin >= 8 && sbin % 2 == 0) ? 2 : 4);
}
constexpr int k1_nw(int sbin) { return k1_px(sbin) * sbin / 4; }
// bins the vector kernel is instantiated for
inline bool k1_vector_bin(int sbin) {
  return sbin == 1 || sbin == 2 || sbin == 4 || sbin == 8 || sbin == 16 || sbin == 3 || sbin == 5 || sbin == 6 ||
         sbin == 7 || sbin == 10 || sbin == 12;
}

// streaming (read-once) loads of VB bytes as 32-bit words; word q is skipped (0) when it starts at or beyond
// `nwords_left` (only the last pixel group of a row of a non-power-of-two bin can run past the row end)
template <int NW>
__device__ __forceinline__ void ldg_stream(const uint8_t* p, unsigned (&w)[NW], int nwords_left = NW) {
  if constexpr (NW != 1 && NW != 2 && NW != 4) {
#pragma unroll
    for (int q = 0; q < NW; ++q) {
      w[q] = 0u;
      if (q < nwords_left) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[q]) : "l"(p + 4 * q));
    }
  } else if constexpr (NW == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                 : "l"(p));
  } else if constexpr (NW == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
  }
}

// mask words (read by every slab and frame: ordinary cached loads)
template <int NW>
__device__ __forceinline__ void ldg_mask(const uint8_t* p, unsigned (&w)[NW], int nwords_left = NW) {
  if constexpr (NW != 1 && NW != 2 && NW != 4) {
#pragma unroll
    for (int q = 0; q < NW; ++q) w[q] = q < nwords_left ? __ldg(reinterpret_cast<const unsigned*>(p) + q) : 0u;
  } else if constexpr (NW == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else if constexpr (NW == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = v.x; w[1] = v.y;
  } else {
    w[0] = __ldg(reinterpret_cast<const unsigned*>(p));
  }
}

// Add the bytes of one raw row segment (NW words) into the PX block sums.
template <int SBIN, int PX, int NW>
__device__ __forceinline__ void accumulate_row(const unsigned (&w)[NW], int (&s)[PX]) {
  if constexpr (!k1_pow2(SBIN)) {
    // binned pixel j owns bytes [j*SBIN, (j+1)*SBIN) of the segment: one dp4a per (pixel, overlapping word)
#pragma unroll
    for (int j = 0; j < PX; ++j) {
#pragma unroll
      for (int q = (j * SBIN) / 4; q <= ((j + 1) * SBIN - 1) / 4; ++q) {
        const int lo = (j * SBIN > 4 * q ? j * SBIN : 4 * q) - 4 * q;
        const int hi = ((j + 1) * SBIN < 4 * q + 4 ? (j + 1) * SBIN : 4 * q + 4) - 4 * q;
        const unsigned ones = (0x01010101u >> (8 * (4 - hi))) & (0x01010101u << (8 * lo));
        s[j] = (int)__dp4a(w[q], ones, (unsigned)s[j]);
      }
    }
  } else if constexpr (SBIN >= 4) {
    constexpr int WPP = SBIN / 4;  // words per binned pixel
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
      for (int q = 0; q < WPP; ++q) s[j] = (int)__dp4a(w[j * WPP + q], 0x01010101u, (unsigned)s[j]);
  } else if constexpr (SBIN == 2) {
#pragma unroll
    for (int j = 0; j < PX; ++j)
      s[j] = (int)__dp4a(w[j >> 1], (j & 1) ? 0x01010000u : 0x00000101u, (unsigned)s[j]);
  } else {
#pragma unroll
    for (int j = 0; j < PX; ++j) s[j] += (int)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
  }
}

template <int PX>
__device__ __forceinline__ void store_row(float* dst, const float (&x)[PX], int nvalid, bool vec_ok) {
  if (vec_ok && nvalid == PX) {
    if constexpr (PX == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
      return;
    } else if constexpr (PX == 2) {
      *reinterpret_cast<float2*>(dst) = make_float2(x[0], x[1]);
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < PX; ++j)
    if (j < nvalid) dst[j] = x[j];
}

// PX integer values (< 65 536) as two byte planes; one packed store per plane when the destination allows
template <int PX>
__device__ __forceinline__ void store_planes(uint8_t* hi, uint8_t* lo, size_t off, const unsigned (&v)[PX], int nvalid,
                                             bool pack_ok) {
  if (pack_ok && nvalid == PX && PX >= 2) {
    unsigned wl = 0, wh = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      wl |= (v[j] & 0xffu) << (8 * j);
      wh |= ((v[j] >> 8) & 0xffu) << (8 * j);
    }
    if constexpr (PX == 4) {
      *reinterpret_cast<unsigned*>(lo + off) = wl;
      if (hi) *reinterpret_cast<unsigned*>(hi + off) = wh;
    } else {
      *reinterpret_cast<unsigned short*>(lo + off) = (unsigned short)wl;
      if (hi) *reinterpret_cast<unsigned short*>(hi + off) = (unsigned short)wh;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < PX; ++j)
    if (j < nvalid) {
      lo[off + j] = (uint8_t)(v[j] & 0xffu);
      if (hi) hi[off + j] = (uint8_t)(v[j] >> 8);
    }
}

// ---------------------------------------------------------------------------------------------
// Vector path: sbin in {1,2,4,8,16}; a thread reads VB = min(16, 4*sbin) bytes per raw row
// (PX = VB/sbin <= 4 binned pixels).  Needs W % VB == 0 and VB-aligned frames.
// ROI energies and the movie operand are compile-time options so the common motion-only
// instantiation stays light on registers.
// ---------------------------------------------------------------------------------------------
// PAINT ORs a per-pixel byte mask into the frames as they are loaded (facemap/process.py:543, 567: pupil and
// blink ROIs paint the chunk buffer that is binned afterwards) — no painted copy of the frames is made.
template <int SBIN, bool ROI, bool MOV, bool PAINT, bool PLANES = false>
__global__ void __launch_bounds__(K1_THREADS) bin_motion_vec_kernel(const K1Params p) {
  constexpr int PX = k1_px(SBIN);
  constexpr int NW = k1_nw(SBIN);
  constexpr bool POW2 = k1_pow2(SBIN);
  // frames fetched per batch (>= 8 load instructions in flight: one vector load per raw row for power-of-two
  // bins, NW scalar loads per raw row otherwise)
  constexpr int FB = POW2 ? ((SBIN >= 8) ? 1 : 8 / SBIN) : 1;
  constexpr float INV_S2 = 1.0f / (float)(SBIN * SBIN);
  constexpr double S2 = (double)(SBIN * SBIN);
  constexpr int NROI_SLOTS = ROI ? FM_MAX_ROIS : 0;

  __shared__ unsigned s_energy[(1 + NROI_SLOTS) * K1_MAX_SLAB];

  const int GX = (p.Lxb + PX - 1) / PX;
  const int ngroups = p.Lyb * GX;
  const int g = blockIdx.x * K1_THREADS + threadIdx.x;
  const bool active = g < ngroups;
  const int r0 = blockIdx.y * p.slab;
  const int r1 = min(p.rows, r0 + p.slab);
  const int nroi = ROI ? p.rois.n : 0;

  for (int i = threadIdx.x; i < (1 + nroi) * K1_MAX_SLAB; i += K1_THREADS) s_energy[i] = 0u;
  __syncthreads();

  const int by = active ? g / GX : 0;
  const int gx = active ? g - by * GX : 0;
  const int bx0 = gx * PX;
  const int nvalid = active ? min(PX, p.Lxb - bx0) : 0;
  const int64_t pix0 = (int64_t)by * p.Lxb + bx0;
  const size_t frame_bytes = (size_t)p.H * p.W;
  const uint8_t* base = p.frames + (size_t)by * SBIN * p.W + (size_t)bx0 * SBIN;

  float am[PX], af[MOV ? PX : 1];
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    am[j] = (p.avgmotion && j < nvalid) ? p.avgmotion[pix0 + j] : 0.f;
    if constexpr (MOV) af[j] = (p.avgframe && j < nvalid) ? p.avgframe[pix0 + j] : 0.f;
  }
  // bit (4*q + j): binned pixel j of this thread lies inside ROI q
  unsigned roimask = 0u;
  if constexpr (ROI) {
    for (int q = 0; q < nroi; ++q)
      if (by >= p.rois.y0[q] && by < p.rois.y1[q])
#pragma unroll
        for (int j = 0; j < PX; ++j)
          if (j < nvalid && bx0 + j >= p.rois.x0[q] && bx0 + j < p.rois.x1[q]) roimask |= 1u << (4 * q + j);
  }
  constexpr int VALIGN = (PX >= 4) ? 4 : 2;  // float4 / float2 stores
  const bool vec_ok = (PX >= 2) && ((p.col0 + pix0) % VALIGN == 0) && (p.ldx % VALIGN == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.x_mot) | reinterpret_cast<uintptr_t>(p.x_mov)) % 16 == 0);

  // words of the raw row from this thread's first byte on (only non-power-of-two bins can run past the end)
  const int wleft = POW2 ? NW : (p.W - bx0 * SBIN) / 4;

  // paint words of this thread's raw rows: registers for small sbin, re-read (L1) for sbin >= 8
  constexpr bool PAINT_REGS = PAINT && SBIN * NW <= 16;
  unsigned pm[PAINT_REGS ? SBIN : 1][NW];
  const uint8_t* mbase = PAINT ? p.ormask + (size_t)by * SBIN * p.W + (size_t)bx0 * SBIN : nullptr;
  if constexpr (PAINT_REGS) {
#pragma unroll
    for (int r = 0; r < SBIN; ++r) {
      if (active) ldg_mask<NW>(mbase + (size_t)r * p.W, pm[r], wleft);
      else
#pragma unroll
        for (int q = 0; q < NW; ++q) pm[r][q] = 0u;
    }
  }
  auto painted = [&](const unsigned (&raw_row)[NW], int r, unsigned (&out)[NW]) {
    if constexpr (PAINT_REGS) {
#pragma unroll
      for (int q = 0; q < NW; ++q) out[q] = raw_row[q] | pm[r][q];
    } else if constexpr (PAINT) {
      unsigned m[NW];
      ldg_mask<NW>(mbase + (size_t)r * p.W, m, wleft);
#pragma unroll
      for (int q = 0; q < NW; ++q) out[q] = raw_row[q] | m[q];
    } else {
#pragma unroll
      for (int q = 0; q < NW; ++q) out[q] = raw_row[q];
    }
  };

  // previous frame's block sums
  int prev[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) prev[j] = 0;
  const int f_prev = r0 + p.frame_off - 1;  // frame preceding the slab's first row
  int tsb[PX], tsa[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) tsb[j] = tsa[j] = 0;
  if (active) {
    if (f_prev >= 0) {
      const uint8_t* fp = base + (size_t)f_prev * frame_bytes;
#pragma unroll
      for (int r = 0; r < SBIN; ++r) {
        unsigned w[NW], wp[NW];
        ldg_stream<NW>(fp + (size_t)r * p.W, w, wleft);
        painted(w, r, wp);
        accumulate_row<SBIN, PX, NW>(wp, prev);
      }
      if (r0 == 0) {  // halo frame of slab 0 is frame 0 of a carry-less call: part of the time sum
#pragma unroll
        for (int j = 0; j < PX; ++j) tsb[j] += prev[j];
      }
    } else if (p.prev_sum) {
#pragma unroll
      for (int j = 0; j < PX; ++j) prev[j] = (j < nvalid) ? p.prev_sum[pix0 + j] : 0;
    }
  }

  for (int rb = r0; rb < r1; rb += FB) {
    unsigned raw[FB][SBIN][NW];
    if (active) {
#pragma unroll
      for (int f = 0; f < FB; ++f) {
        const uint8_t* fp = base + (size_t)(min(rb + f, r1 - 1) + p.frame_off) * frame_bytes;
#pragma unroll
        for (int r = 0; r < SBIN; ++r) ldg_stream<NW>(fp + (size_t)r * p.W, raw[f][r], wleft);
      }
    }
#pragma unroll
    for (int f = 0; f < FB; ++f) {
      const int row = rb + f;
      if (row >= r1) break;  // uniform across the block
      unsigned esum = 0;
      unsigned dvals[PX];
      if (active) {
        int cur[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) cur[j] = 0;
#pragma unroll
        for (int r = 0; r < SBIN; ++r) {
          if constexpr (PAINT) {
            unsigned wp[NW];
            painted(raw[f][r], r, wp);
            accumulate_row<SBIN, PX, NW>(wp, cur);
          } else {
            accumulate_row<SBIN, PX, NW>(raw[f][r], cur);
          }
        }
        float xm[PX], xf[MOV ? PX : 1];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const int d = abs(cur[j] - prev[j]);
          dvals[j] = (j < nvalid) ? (unsigned)d : 0u;
          if constexpr (POW2) {       // exact: d / s^2 is a float, one rounding in the FMA
            xm[j] = __fmaf_rn((float)d, INV_S2, -am[j]);
            if constexpr (MOV) xf[j] = __fmaf_rn((float)cur[j], INV_S2, -af[j]);
          } else {                    // the reference's float64 evaluation, as in the generic kernel
            xm[j] = (float)((double)d / S2 - (double)am[j]);
            if constexpr (MOV) xf[j] = (float)((double)cur[j] / S2 - (double)af[j]);
          }
          if (j < nvalid) {
            esum += (unsigned)d;
            tsa[j] += d;
            tsb[j] += cur[j];
          }
          prev[j] = cur[j];
        }
        if (p.x_mot) store_row<PX>(p.x_mot + (size_t)row * p.ldx + p.col0 + pix0, xm, nvalid, vec_ok);
        if constexpr (MOV)
          if (p.x_mov) store_row<PX>(p.x_mov + (size_t)row * p.ldx + p.col0 + pix0, xf, nvalid, vec_ok);
        if constexpr (PLANES) {                 // integer operands of the kind::i8 projection (own instantiations:
                                                // the default kernels' code is untouched)
          const size_t qoff = (size_t)row * p.ldq + p.col0 + pix0;
          const bool pack_ok = ((p.col0 + pix0) % PX == 0) && (p.ldq % 4 == 0) &&
                               ((reinterpret_cast<uintptr_t>(p.q_mot_hi) | reinterpret_cast<uintptr_t>(p.q_mot_lo) |
                                 reinterpret_cast<uintptr_t>(p.q_mov_hi) | reinterpret_cast<uintptr_t>(p.q_mov_lo)) % 4 == 0);
          if (p.q_mot_lo) store_planes<PX>(p.q_mot_hi, p.q_mot_lo, qoff, dvals, nvalid, pack_ok);
          if (p.q_mov_lo) {
            unsigned cv[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) cv[j] = (unsigned)prev[j];     // prev == cur after the loop above
            store_planes<PX>(p.q_mov_hi, p.q_mov_lo, qoff, cv, nvalid, pack_ok);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < PX; ++j) dvals[j] = 0u;
      }
      if (p.energy) {
        const unsigned wsum = __reduce_add_sync(0xffffffffu, esum);
        if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_energy[row - r0], wsum);
      }
      if constexpr (ROI) {
        for (int q = 0; q < nroi; ++q) {
          unsigned e = 0;
#pragma unroll
          for (int j = 0; j < PX; ++j)
            if ((roimask >> (4 * q + j)) & 1u) e += dvals[j];
          const unsigned wsum = __reduce_add_sync(0xffffffffu, e);
          if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&s_energy[(1 + q) * K1_MAX_SLAB + row - r0], wsum);
        }
      }
    }
  }

  if (active) {
    if (p.last_sum && r1 == p.rows) {
#pragma unroll
      for (int j = 0; j < PX; ++j)
        if (j < nvalid) p.last_sum[pix0 + j] = prev[j];
    }
    if (p.tsum_bin) {
#pragma unroll
      for (int j = 0; j < PX; ++j)
        if (j < nvalid) atomicAdd(&p.tsum_bin[pix0 + j], (unsigned long long)tsb[j]);
    }
    if (p.tsum_adiff) {
#pragma unroll
      for (int j = 0; j < PX; ++j)
        if (j < nvalid) atomicAdd(&p.tsum_adiff[pix0 + j], (unsigned long long)tsa[j]);
    }
  }
  __syncthreads();
  const int nr = r1 - r0;
  if (p.energy)
    for (int i = threadIdx.x; i < nr; i += K1_THREADS)
      if (s_energy[i]) atomicAdd(&p.energy[r0 + i], (unsigned long long)s_energy[i]);
  if constexpr (ROI) {
    for (int i = threadIdx.x; i < nr * nroi; i += K1_THREADS) {
      const int q = i / nr, rr = i - q * nr;
      const unsigned v = s_energy[(1 + q) * K1_MAX_SLAB + rr];
      if (v) atomicAdd(&p.roi_energy[(size_t)q * p.rows + r0 + rr], (unsigned long long)v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic path: any sbin / width / alignment.  One thread per binned pixel, byte loads,
// float64 division (the reference's float64 means are reproduced to <= 1 ulp of float32).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum_bytes(const uint8_t* p, int sbin, int W, const uint8_t* m) {
  int s = 0;
  for (int r = 0; r < sbin; ++r) {
    const uint8_t* q = p + (size_t)r * W;
    if (m) {
      const uint8_t* mq = m + (size_t)r * W;
      for (int c = 0; c < sbin; ++c) s += (int)(__ldg(q + c) | __ldg(mq + c));
    } else {
      for (int c = 0; c < sbin; ++c) s += (int)__ldg(q + c);
    }
  }
  return s;
}

__global__ void __launch_bounds__(K1_THREADS) bin_motion_generic_kernel(const K1Params p) {
  __shared__ unsigned long long s_energy[(1 + FM_MAX_ROIS) * K1_MAX_SLAB];
  const int npix = p.Lyb * p.Lxb;
  const int g = blockIdx.x * K1_THREADS + threadIdx.x;
  const bool active = g < npix;
  const int r0 = blockIdx.y * p.slab;
  const int r1 = min(p.rows, r0 + p.slab);
  const int nroi = p.rois.n;
  for (int i = threadIdx.x; i < (1 + nroi) * K1_MAX_SLAB; i += K1_THREADS) s_energy[i] = 0ull;
  __syncthreads();

  const int by = active ? g / p.Lxb : 0;
  const int bx = active ? g - by * p.Lxb : 0;
  const size_t frame_bytes = (size_t)p.H * p.W;
  const uint8_t* base = p.frames + (size_t)by * p.sbin * p.W + (size_t)bx * p.sbin;
  const uint8_t* mbase = p.ormask ? p.ormask + (size_t)by * p.sbin * p.W + (size_t)bx * p.sbin : nullptr;
  const double s2 = (double)p.sbin * (double)p.sbin;
  const double am = (active && p.avgmotion) ? (double)p.avgmotion[g] : 0.0;
  const double af = (active && p.avgframe) ? (double)p.avgframe[g] : 0.0;
  unsigned inroi = 0;
  for (int q = 0; q < nroi; ++q)
    if (active && by >= p.rois.y0[q] && by < p.rois.y1[q] && bx >= p.rois.x0[q] && bx < p.rois.x1[q])
      inroi |= 1u << q;

  int prev = 0;
  long long tsb = 0, tsa = 0;
  const int f_prev = r0 + p.frame_off - 1;
  if (active) {
    if (f_prev >= 0) {
      prev = block_sum_bytes(base + (size_t)f_prev * frame_bytes, p.sbin, p.W, mbase);
      if (r0 == 0) tsb += prev;
    } else if (p.prev_sum) {
      prev = p.prev_sum[g];
    }
  }
  for (int row = r0; row < r1; ++row) {
    unsigned long long e = 0;
    if (active) {
      const int cur = block_sum_bytes(base + (size_t)(row + p.frame_off) * frame_bytes, p.sbin, p.W, mbase);
      const int d = abs(cur - prev);
      if (p.x_mot) p.x_mot[(size_t)row * p.ldx + p.col0 + g] = (float)((double)d / s2 - am);
      if (p.x_mov) p.x_mov[(size_t)row * p.ldx + p.col0 + g] = (float)((double)cur / s2 - af);
      if (p.q_mot_lo) {
        p.q_mot_lo[(size_t)row * p.ldq + p.col0 + g] = (uint8_t)(d & 0xff);
        if (p.q_mot_hi) p.q_mot_hi[(size_t)row * p.ldq + p.col0 + g] = (uint8_t)(d >> 8);
      }
      if (p.q_mov_lo) {
        p.q_mov_lo[(size_t)row * p.ldq + p.col0 + g] = (uint8_t)(cur & 0xff);
        if (p.q_mov_hi) p.q_mov_hi[(size_t)row * p.ldq + p.col0 + g] = (uint8_t)(cur >> 8);
      }
      e = (unsigned long long)d;
      tsa += d;
      tsb += cur;
      prev = cur;
    }
    if (p.energy && e) atomicAdd(&s_energy[row - r0], e);
    if (p.roi_energy && e)
      for (int q = 0; q < nroi; ++q)
        if ((inroi >> q) & 1u) atomicAdd(&s_energy[(1 + q) * K1_MAX_SLAB + row - r0], e);
  }
  if (active) {
    if (p.last_sum && r1 == p.rows) p.last_sum[g] = prev;
    if (p.tsum_bin) atomicAdd(&p.tsum_bin[g], (unsigned long long)tsb);
    if (p.tsum_adiff) atomicAdd(&p.tsum_adiff[g], (unsigned long long)tsa);
  }
  __syncthreads();
  const int nr = r1 - r0;
  if (p.energy)
    for (int i = threadIdx.x; i < nr; i += K1_THREADS)
      if (s_energy[i]) atomicAdd(&p.energy[r0 + i], s_energy[i]);
  if (p.roi_energy)
    for (int i = threadIdx.x; i < nr * nroi; i += K1_THREADS) {
      const int q = i / nr, rr = i - q * nr;
      const unsigned long long v = s_energy[(1 + q) * K1_MAX_SLAB + rr];
      if (v) atomicAdd(&p.roi_energy[(size_t)q * p.rows + r0 + rr], v);
    }
}

// stage-1 accumulator update (facemap/process.py:82-86, 95-96)
__global__ void mean_update_kernel(float* acc, const long long* tsum, int64_t n, double s2, double count,
                                   float final_div) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = acc[i];
  if (tsum) {
    // sbin == 1 stays float32 end to end in the reference (process.py:40-41): float32 mean, float32 add
    if (s2 == 1.0) a = a + __fdiv_rn((float)tsum[i], (float)count);
    else a = (float)((double)a + ((double)tsum[i] / s2) / count);
  }
  if (final_div > 0.f) a = a / final_div;
  acc[i] = a;
}

}  // namespace fm

static int bin_motion_impl(const uint8_t* frames, int T, int H, int W, int sbin,
                           const int32_t* prev_sum, const float* avgmotion, const float* avgframe,
                           float* x_mot, float* x_mov, int64_t ldx, int64_t col0,
                           unsigned long long* energy, const fm_rois* rois,
                           unsigned long long* roi_energy, int32_t* last_sum, long long* tsum_bin,
                           long long* tsum_adiff, const uint8_t* ormask, void* stream,
                           uint8_t* q_mot_hi = nullptr, uint8_t* q_mot_lo = nullptr, uint8_t* q_mov_hi = nullptr,
                           uint8_t* q_mov_lo = nullptr, int64_t ldq = 0) {
  using namespace fm;
  FM_REQUIRE(frames != nullptr && T >= 1 && H >= 1 && W >= 1 && sbin >= 1, "fm_bin_motion: bad shape");
  if (q_mot_lo || q_mov_lo) {
    FM_REQUIRE(sbin <= 16, "fm_bin_motion_planes: two byte planes hold block sums up to sbin 16, got %d", sbin);
    FM_REQUIRE(sbin == 1 || ((q_mot_lo == nullptr || q_mot_hi) && (q_mov_lo == nullptr || q_mov_hi)),
               "fm_bin_motion_planes: sbin > 1 needs the high planes");
    FM_REQUIRE(ldq >= col0 + (int64_t)(H / sbin) * (W / sbin), "fm_bin_motion_planes: ldq too small");
  }
  const int Lyb = H / sbin, Lxb = W / sbin;
  FM_REQUIRE(Lyb >= 1 && Lxb >= 1, "fm_bin_motion: sbin %d larger than the %dx%d frame", sbin, H, W);
  FM_REQUIRE((int64_t)sbin * sbin * 255 * K1_MAX_SLAB < (1ll << 31), "fm_bin_motion: sbin too large");
  FM_REQUIRE(roi_energy == nullptr || (rois != nullptr && rois->n >= 0 && rois->n <= FM_MAX_ROIS),
             "fm_bin_motion: roi_energy needs 0..%d rois", FM_MAX_ROIS);
  K1Params p{};
  p.frames = frames; p.T = T; p.H = H; p.W = W; p.sbin = sbin; p.Lyb = Lyb; p.Lxb = Lxb;
  p.prev_sum = prev_sum; p.avgmotion = avgmotion; p.avgframe = avgframe;
  p.x_mot = x_mot; p.x_mov = x_mov; p.ldx = ldx; p.col0 = col0;
  p.energy = energy; p.roi_energy = roi_energy; p.last_sum = last_sum;
  p.tsum_bin = reinterpret_cast<unsigned long long*>(tsum_bin);
  p.tsum_adiff = reinterpret_cast<unsigned long long*>(tsum_adiff);
  p.rows = prev_sum ? T : T - 1;
  p.frame_off = prev_sum ? 0 : 1;
  if (rois && roi_energy) p.rois = *rois; else p.rois.n = 0;
  p.ormask = ormask;
  p.q_mot_hi = q_mot_hi; p.q_mot_lo = q_mot_lo; p.q_mov_hi = q_mov_hi; p.q_mov_lo = q_mov_lo; p.ldq = ldq;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  if (p.rows == 0) {
    // single frame without carry: no difference rows, but the frame still feeds the running sums
    if (last_sum == nullptr && tsum_bin == nullptr) return FM_OK;
    p.rows = 0;
  }
  // power-of-two bins: whole 16-byte (4*sbin below sbin 4) groups per row; other bins: 4-byte aligned rows
  const int vb = k1_pow2(sbin) ? (sbin >= 4 ? 16 : 4 * sbin) : 4;
  const bool vec = k1_vector_bin(sbin) && (W % vb == 0) && ((reinterpret_cast<uintptr_t>(frames) % vb) == 0) &&
                   ((reinterpret_cast<uintptr_t>(ormask) % vb) == 0);
  const int px = vec ? k1_px(sbin) : 1;
  const int64_t ngroups = vec ? (int64_t)Lyb * ((Lxb + px - 1) / px) : (int64_t)Lyb * Lxb;
  const int gx = (int)cdiv(ngroups, K1_THREADS);
  // slab length: long enough to amortise the halo frame, short enough to fill the machine
  int slab = K1_MAX_SLAB;
  while (slab > 16 && (int64_t)gx * cdiv(std::max(p.rows, 1), slab) < 148 * 8) slab >>= 1;
  p.slab = slab;
  const int gy = (int)std::max<int64_t>(1, cdiv(p.rows, slab));
  FM_REQUIRE(gy <= 65535, "fm_bin_motion: too many rows per call (%d)", p.rows);
  dim3 grid(gx, gy);
  const bool roi = p.rois.n > 0;
  const bool mov = x_mov != nullptr;
  const bool paint = ormask != nullptr;
  const bool planes = q_mot_lo != nullptr || q_mov_lo != nullptr;     // float operands are not written then
#define FM_K1_LAUNCH_P(S, PT)                                                                        \
  do {                                                                                               \
    if (planes && roi) bin_motion_vec_kernel<S, true, false, PT, true><<<grid, K1_THREADS, 0, st>>>(p); \
    else if (planes) bin_motion_vec_kernel<S, false, false, PT, true><<<grid, K1_THREADS, 0, st>>>(p); \
    else if (roi && mov) bin_motion_vec_kernel<S, true, true, PT><<<grid, K1_THREADS, 0, st>>>(p);   \
    else if (roi) bin_motion_vec_kernel<S, true, false, PT><<<grid, K1_THREADS, 0, st>>>(p);         \
    else if (mov) bin_motion_vec_kernel<S, false, true, PT><<<grid, K1_THREADS, 0, st>>>(p);         \
    else bin_motion_vec_kernel<S, false, false, PT><<<grid, K1_THREADS, 0, st>>>(p);                 \
  } while (0)
#define FM_K1_LAUNCH(S)                                                                              \
  do {                                                                                               \
    if (paint) FM_K1_LAUNCH_P(S, true);                                                              \
    else FM_K1_LAUNCH_P(S, false);                                                                   \
  } while (0)
  if (vec) {
    switch (sbin) {
      case 1: FM_K1_LAUNCH(1); break;
      case 2: FM_K1_LAUNCH(2); break;
      case 4: FM_K1_LAUNCH(4); break;
      case 8: FM_K1_LAUNCH(8); break;
      case 16: FM_K1_LAUNCH(16); break;
      case 3: FM_K1_LAUNCH(3); break;
      case 5: FM_K1_LAUNCH(5); break;
      case 6: FM_K1_LAUNCH(6); break;
      case 7: FM_K1_LAUNCH(7); break;
      case 10: FM_K1_LAUNCH(10); break;
      default: FM_K1_LAUNCH(12); break;
    }
  } else {
    bin_motion_generic_kernel<<<grid, K1_THREADS, 0, st>>>(p);
  }
#undef FM_K1_LAUNCH
#undef FM_K1_LAUNCH_P
  FM_LAUNCH_CHECK();
  return FM_OK;
}

extern "C" int fm_bin_motion(const uint8_t* frames, int T, int H, int W, int sbin,
                             const int32_t* prev_sum, const float* avgmotion, const float* avgframe,
                             float* x_mot, float* x_mov, int64_t ldx, int64_t col0,
                             unsigned long long* energy, const fm_rois* rois,
                             unsigned long long* roi_energy, int32_t* last_sum, long long* tsum_bin,
                             long long* tsum_adiff, void* stream) {
  return bin_motion_impl(frames, T, H, W, sbin, prev_sum, avgmotion, avgframe, x_mot, x_mov, ldx, col0, energy, rois,
                         roi_energy, last_sum, tsum_bin, tsum_adiff, nullptr, stream);
}

extern "C" int fm_bin_motion_painted(const uint8_t* frames, int T, int H, int W, int sbin,
                                     const int32_t* prev_sum, const float* avgmotion, const float* avgframe,
                                     float* x_mot, float* x_mov, int64_t ldx, int64_t col0,
                                     unsigned long long* energy, const fm_rois* rois,
                                     unsigned long long* roi_energy, int32_t* last_sum, long long* tsum_bin,
                                     long long* tsum_adiff, const uint8_t* ormask, void* stream) {
  FM_REQUIRE(ormask != nullptr, "fm_bin_motion_painted: ormask is NULL (use fm_bin_motion)");
  return bin_motion_impl(frames, T, H, W, sbin, prev_sum, avgmotion, avgframe, x_mot, x_mov, ldx, col0, energy, rois,
                         roi_energy, last_sum, tsum_bin, tsum_adiff, ormask, stream);
}

extern "C" int fm_bin_motion_planes(const uint8_t* frames, int T, int H, int W, int sbin, const int32_t* prev_sum,
                                    uint8_t* q_mot_hi, uint8_t* q_mot_lo, uint8_t* q_mov_hi, uint8_t* q_mov_lo,
                                    int64_t ldq, int64_t col0, unsigned long long* energy, const fm_rois* rois,
                                    unsigned long long* roi_energy, int32_t* last_sum, const uint8_t* ormask,
                                    void* stream) {
  FM_REQUIRE(q_mot_lo != nullptr || q_mov_lo != nullptr, "fm_bin_motion_planes: no plane to write (use fm_bin_motion)");
  return bin_motion_impl(frames, T, H, W, sbin, prev_sum, nullptr, nullptr, nullptr, nullptr, 0, col0, energy, rois,
                         roi_energy, last_sum, nullptr, nullptr, ormask, stream, q_mot_hi, q_mot_lo, q_mov_hi, q_mov_lo,
                         ldq);
}

extern "C" int fm_mean_update(float* acc, const long long* tsum, int64_t n, int sbin, int count,
                              int finalize_div, void* stream) {
  using namespace fm;
  FM_REQUIRE(acc != nullptr && n >= 0 && sbin >= 1 && (tsum == nullptr || count >= 1), "fm_mean_update: bad args");
  if (n == 0) return FM_OK;
  mean_update_kernel<<<(unsigned)cdiv(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      acc, tsum, n, (double)sbin * sbin, (double)count, (float)finalize_div);
  FM_LAUNCH_CHECK();
  return FM_OK;
}
