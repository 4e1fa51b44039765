// Scalar pieces of the pupil fit that decide ordering / signs / rounding, shared between the CUDA kernel
// and the CPU unit test that pins them against NumPy + LAPACK (tests/test_numpy_sum.py).
#pragma once
#include <math.h>
#include "numpy_sum.cuh"

namespace fm {

// inverse of the symmetric 2x2 [[a, b], [b, c]] by LU with partial pivoting (what numpy.linalg.inv runs)
FM_HD inline void inv2x2(double a, double b, double c, double& i00, double& i01, double& i10,
                                       double& i11) {
  if (fabs(b) > fabs(a)) {           // rows swapped: [[b, c], [a, b]]
    const double l = a / b, u22 = b - l * c;
    // solve for e1 (permuted rhs (0, 1)) and e2 (permuted rhs (1, 0))
    double x2 = 1.0 / u22;
    i10 = x2;
    i00 = (0.0 - c * x2) / b;
    x2 = (0.0 - l) / u22;
    i11 = x2;
    i01 = (1.0 - c * x2) / b;
  } else {
    const double l = b / a, u22 = c - l * b;
    double x2 = (0.0 - l) / u22;
    i10 = x2;
    i00 = (1.0 - b * x2) / a;
    x2 = 1.0 / u22;
    i11 = x2;
    i01 = (0.0 - b * x2) / a;
  }
}


// numpy.median of the float32 values {0 (count hist[256])} U {max(0, (255 - v) - dark_offset) (count hist[v])}:
// ascending order is the zeros, then v = 255 .. 0; an even count averages the two middle values in float32.
FM_HD inline float median_from_hist(const unsigned* hist, float dark_offset) {
  unsigned cnt = 0;
  for (int h = 0; h < 257; ++h) cnt += hist[h];
  if (cnt == 0) return nanf("");
  const unsigned k1 = (cnt - 1) / 2, k2 = cnt / 2;
  float v1 = 0.f, v2 = 0.f;
  unsigned seen = hist[256];
  bool got1 = k1 < seen, got2 = k2 < seen;
  for (int pix = 255; pix >= 0 && !(got1 && got2); --pix) {
    const unsigned c = hist[pix];
    if (c == 0) continue;
    const float v = fmaxf(0.f, fadd_rn((float)(255 - pix), -dark_offset));
    if (!got1 && k1 < seen + c) { v1 = v; got1 = true; }
    if (!got2 && k2 < seen + c) { v2 = v; got2 = true; }
    seen += c;
  }
  return k1 == k2 ? v1 : fadd_rn(v1, v2) * 0.5f;
}

// Eigen-decomposition of the symmetric [[a, b], [b, c]] in the order and with the signs numpy.linalg.eig
// (dgeev -> dlahqr) returns: a negligible off-diagonal deflates (eigenvalues = diagonal, vectors from the
// triangular solve), otherwise dlanv2's rotation.  Columns of [[u00, u01], [u10, u11]] are the vectors.
FM_HD inline void sym2x2_eig_lapack(double a, double b, double c, double& w0, double& w1, double& u00, double& u01,
                                    double& u10, double& u11) {
  bool deflate = b == 0.0;
  if (!deflate && fabs(b) <= 0x1p-52 * (fabs(a) + fabs(c))) {
    const double ab = fabs(b), aa = fmax(fabs(c), fabs(a - c)), bb = fmin(fabs(c), fabs(a - c)), s = aa + ab;
    deflate = ab * (ab / s) <= fmax(1e-290, 0x1p-52 * (bb * (aa / s)));
  }
  if (deflate) {
    w0 = a;
    w1 = c;
    const double x = b == 0.0 ? 0.0 : -b / (a - c), nrm = hypot(x, 1.0);
    u00 = 1.0;
    u10 = 0.0;
    u01 = x / nrm;
    u11 = 1.0 / nrm;
  } else {
    const double p = 0.5 * (a - c), bc = fabs(b);
    const double scale = fmax(fabs(p), bc);
    double z = p / scale * p + bc / scale * bc;
    z = p + copysign(sqrt(scale) * sqrt(z), p);
    w0 = c + z;
    w1 = c - bc / z * bc;
    const double tau = hypot(b, z);
    u00 = z / tau;
    u10 = b / tau;
    u01 = -u10;
    u11 = u00;
  }
}

}  // namespace fm
