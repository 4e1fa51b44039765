// Test shim: exposes the host/device scalar helpers of facemap_b200/csrc/{numpy_sum,pupil_math}.cuh to
// ctypes so the CPU suite can pin them against NumPy / LAPACK (tests/test_numpy_sum.py).  Not shipped.
#include "../../facemap_b200/csrc/pupil_math.cuh"

extern "C" float shim_pairwise_sum_f32(const float* a, int n) { return fm::pairwise_sum_f32(a, n); }

// the split form the kernel uses: list leaves, sum each leaf, fold along the tree
extern "C" float shim_pairwise_sum_split(const float* a, int n, int cap) {
  static int lo[4096], len[4096];
  static float sums[4096];
  const int count = fm::pairwise_list_leaves(n, lo, len, cap);
  if (count > cap) return fm::pairwise_sum_f32(a, n);
  for (int i = 0; i < count; ++i) sums[i] = fm::pairwise_leaf(a + lo[i], len[i]);
  return fm::fadd_rn(0.0f, fm::pairwise_fold_leaves(n, sums));
}

// the flat plan the kernel uses: leaves + fold counts, no tree walk per sum
extern "C" float shim_pairwise_sum_plan(const float* a, int n, int cap) {
  static int lo[4096], len[4096], st[96];
  static unsigned char comb[4096];
  static float sums[4096], vstack[32];
  const int count = fm::pairwise_plan(n, lo, len, comb, cap, st);
  if (count > cap) return fm::pairwise_sum_f32(a, n);
  for (int i = 0; i < count; ++i) sums[i] = fm::pairwise_leaf(a + lo[i], len[i]);
  return fm::fadd_rn(0.0f, fm::pairwise_fold_plan(count, sums, comb, vstack));
}

// float64 flavour (numpy's DOUBLE_pairwise_sum): what fm_motion_ref runs for sbin > 1
extern "C" double shim_pairwise_sum_plan_f64(const double* a, int n) {
  static int lo[1 << 16], len[1 << 16], st[96];
  static unsigned char comb[1 << 16];
  static double sums[1 << 16], vstack[32];
  const int count = fm::pairwise_plan(n, lo, len, comb, 1 << 16, st);
  for (int i = 0; i < count; ++i) sums[i] = fm::pairwise_leaf(a + lo[i], len[i]);
  return fm::fadd_rn(0.0, fm::pairwise_fold_plan(count, sums, comb, vstack));
}

extern "C" int shim_pairwise_leaf_count(int n) {
  static int lo[4096], len[4096];
  return fm::pairwise_list_leaves(n, lo, len, 4096);
}

extern "C" void shim_inv2x2(double a, double b, double c, double* out) {
  fm::inv2x2(a, b, c, out[0], out[1], out[2], out[3]);
}

extern "C" float shim_median_from_hist(const unsigned* hist, float dark_offset) {
  return fm::median_from_hist(hist, dark_offset);
}

extern "C" void shim_sym2x2_eig(double a, double b, double c, double* out) {
  fm::sym2x2_eig_lapack(a, b, c, out[0], out[1], out[2], out[3], out[4], out[5]);
}

// the row slicing of fm_slice_rows, run on the host exactly like one block of the kernel does
#include "../../facemap_b200/csrc/slice_math.cuh"
extern "C" void shim_slice_rows(const float* x, int rows, int cols, float* s0, float* s1, float* s2, int* ex_out) {
  for (int r = 0; r < rows; ++r) {
    float amax = 0.0f;
    for (int c = 0; c < cols; ++c) amax = fmaxf(amax, fabsf(x[(size_t)r * cols + c]));
    const int ex = fm::slice_exponent(amax);
    ex_out[r] = ex;
    const fm::SliceGrid g = fm::slice_grid(ex);
    for (int c = 0; c < cols; ++c) {
      const size_t i = (size_t)r * cols + c;
      fm::slice3(x[i], g, s0[i], s1[i], s2[i]);
    }
  }
}

extern "C" void shim_slice_rows_i8(const float* x, int rows, int cols, signed char* q0, unsigned char* q1,
                                   unsigned char* q2, int* ex_out) {
  for (int r = 0; r < rows; ++r) {
    float amax = 0.0f;
    for (int c = 0; c < cols; ++c) amax = fmaxf(amax, fabsf(x[(size_t)r * cols + c]));
    const int ex = fm::slice_exponent(amax);
    ex_out[r] = ex;
    for (int c = 0; c < cols; ++c) {
      const size_t i = (size_t)r * cols + c;
      fm::slice3_i8(x[i], ex, q0[i], q1[i], q2[i]);
    }
  }
}
