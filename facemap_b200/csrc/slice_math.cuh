// Row-scaled 8-bit slices of float32 values (the operand form of the "sliced" Gram, svd.sliced_gram):
//   a = s0 + s1 + s2 + r,  s_d = k_d * 2^(ex - 8(d+1)),  k_d the nearest integer (|k_0| <= 256, |k_1|, |k_2| <= 128),
//   |r| <= 2^(ex - 25)
// where 2^ex bounds every |a| of the row.  Rounding to nearest (signed digits) keeps the representation error
// unbiased; truncation would bias every squared norm by ~2^-22.  A slice has at most 9 significant bits, so it is
// exact as a TF32 operand and the products of two slices of any two rows are integers (times a power of two) of
// at most 2^16: 240 of them add up exactly in a float32 accumulator, whatever its rounding mode.
// Usable from device code and, for the CPU test that pins it against the NumPy model (tests/test_slice_math.py),
// from plain C++.
#pragma once
#include <math.h>
#include "numpy_sum.cuh"

namespace fm {

// ex with |a| < 2^ex for every |a| <= amax; clamped so that every slice grid 2^(ex-24) .. 2^(ex-8) and its
// reciprocal are normal float32 numbers (a row that small is zero for every purpose of this library).
FM_HD inline int slice_exponent(float amax) {
  int ex = -100;
  if (amax > 0.0f && amax <= 3.0e38f) (void)frexpf(amax, &ex);          // amax = m * 2^ex, 0.5 <= m < 1
  return ex < -100 ? -100 : (ex > 126 ? 126 : ex);
}

struct SliceGrid {
  float q0, q1, q2;        // 2^(ex-8), 2^(ex-16), 2^(ex-24)
  float i0, i1, i2;        // their reciprocals
};

FM_HD inline SliceGrid slice_grid(int ex) {
  SliceGrid g;
  g.q0 = ldexpf(1.0f, ex - 8);
  g.q1 = ldexpf(1.0f, ex - 16);
  g.q2 = ldexpf(1.0f, ex - 24);
  g.i0 = ldexpf(1.0f, 8 - ex);
  g.i1 = ldexpf(1.0f, 16 - ex);
  g.i2 = ldexpf(1.0f, 24 - ex);
  return g;
}

// every operation below is exact: scaling by a power of two, rounding to an integer of at most 2^8, and the
// subtraction of a multiple of the grid from a value less than half a grid step away from it
FM_HD inline void slice3(float a, const SliceGrid& g, float& s0, float& s1, float& s2) {
  s0 = rintf(a * g.i0) * g.q0;
  float r = a - s0;
  s1 = rintf(r * g.i1) * g.q1;
  r = r - s1;
  s2 = rintf(r * g.i2) * g.q2;
}

// Integer digits for tcgen05 kind::i8 (fm_slice_rows_i8): with t = a * 2^(7 - ex), |t| < 128,
//   a ~ 2^(ex - 7) * (q0 + q1 / 2^8 + q2 / 2^16),  q0 = floor(t) in [-128, 127] (s8),  q1 = floor of the next 8 bits
//   in [0, 255] (u8),  q2 = the following 8 bits rounded to nearest in [0, 255] (u8), carries propagated upward.
// 23 bits below the row bound, unbiased.  An entry within 2^-24 of the bound whose rounding would carry out of q0
// saturates at 127 + 255/2^8 + 255/2^16.  Evaluated in double so that every step is exact (t - floor(t) of a small
// negative float32 needs 25 bits).
FM_HD inline void slice3_i8(float a, int ex, signed char& q0, unsigned char& q1, unsigned char& q2) {
  const double t = (double)a * ldexp(1.0, 7 - ex);
  double f0 = floor(t);
  double r = (t - f0) * 256.0;                  // in [0, 256)
  double f1 = floor(r);
  r = (r - f1) * 256.0;                         // in [0, 256)
  double f2 = rint(r);
  if (f2 == 256.0) { f2 = 0.0; f1 += 1.0; }
  if (f1 == 256.0) { f1 = 0.0; f0 += 1.0; }
  if (f0 == 128.0) { f0 = 127.0; f1 = 255.0; f2 = 255.0; }
  q0 = (signed char)(int)f0;
  q1 = (unsigned char)(int)f1;
  q2 = (unsigned char)(int)f2;
}

}  // namespace fm
