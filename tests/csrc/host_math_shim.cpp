// Test shim: exposes the host/device scalar helpers of facemap_b200/csrc/{numpy_sum,pupil_math}.cuh to
// ctypes so the CPU suite can pin them against NumPy / LAPACK (tests/test_numpy_sum.py).  Not shipped.
#include "../../facemap_b200/csrc/pupil_math.cuh"

extern "C" float shim_pairwise_sum_f32(const float* a, int n) { return fm::pairwise_sum_f32(a, n); }

extern "C" void shim_inv2x2(double a, double b, double c, double* out) {
  fm::inv2x2(a, b, c, out[0], out[1], out[2], out[3]);
}

extern "C" float shim_median_from_hist(const unsigned* hist, float dark_offset) {
  return fm::median_from_hist(hist, dark_offset);
}

extern "C" void shim_sym2x2_eig(double a, double b, double c, double* out) {
  fm::sym2x2_eig_lapack(a, b, c, out[0], out[1], out[2], out[3], out[4], out[5]);
}
