// NumPy's pairwise float32 summation order, usable from device code and (for the CPU unit test that
// pins it against numpy.sum, tests/test_numpy_sum.py) from plain C++.
#pragma once
#if defined(__CUDACC__)
#define FM_HD __host__ __device__
#else
#define FM_HD
#endif

namespace fm {

// one correctly rounded float32 addition that no compiler may fuse or reassociate
FM_HD inline float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}

FM_HD inline double fadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}

// ------------------------------------------------------------------------------------------------
// NumPy pairwise sum (numpy/_core/src/umath/loops_utils.h.src, FLOAT_/DOUBLE_pairwise_sum — the same
// algorithm for both types), run by ONE thread: the result depends on the association order, which is what
// is being reproduced.
// ------------------------------------------------------------------------------------------------
template <class F>
FM_HD inline F pairwise_leaf(const F* a, int n) {
  if (n < 8) {
    F r = (F)-0.0;
    for (int i = 0; i < n; ++i) r = fadd_rn(r, a[i]);
    return r;
  }
  F r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  const int m = n - (n % 8);
  for (int i = 8; i < m; i += 8) {
    r0 = fadd_rn(r0, a[i + 0]);
    r1 = fadd_rn(r1, a[i + 1]);
    r2 = fadd_rn(r2, a[i + 2]);
    r3 = fadd_rn(r3, a[i + 3]);
    r4 = fadd_rn(r4, a[i + 4]);
    r5 = fadd_rn(r5, a[i + 5]);
    r6 = fadd_rn(r6, a[i + 6]);
    r7 = fadd_rn(r7, a[i + 7]);
  }
  F res = fadd_rn(fadd_rn(fadd_rn(r0, r1), fadd_rn(r2, r3)), fadd_rn(fadd_rn(r4, r5), fadd_rn(r6, r7)));
  for (int i = m; i < n; ++i) res = fadd_rn(res, a[i]);
  return res;
}

// Post-order walk of NumPy's split tree over [0, n) with an explicit stack (depth <= log2(n / 128) + 1).
// `leaf(lo, len)` is called for every <=128-element leaf, left to right, and returns its sum.
template <class Leaf>
FM_HD inline float pairwise_walk(int n, Leaf leaf) {
  int lo_s[32], n_s[32];
  float left_s[32];
  unsigned char phase_s[32];
  int sp = 0;
  lo_s[0] = 0;
  n_s[0] = n;
  phase_s[0] = 0;
  for (;;) {
    if (n_s[sp] > 128) {                   // descend into the left half
      int n2 = n_s[sp] / 2;
      n2 -= n2 % 8;
      phase_s[sp] = 1;
      lo_s[sp + 1] = lo_s[sp];
      n_s[sp + 1] = n2;
      phase_s[sp + 1] = 0;
      ++sp;
      continue;
    }
    float val = leaf(lo_s[sp], n_s[sp]);
    // deliver `val` upwards until a node still waits for its right half
    for (;;) {
      if (sp == 0) return val;
      --sp;
      if (phase_s[sp] == 1) {              // left half done: remember it, descend right
        left_s[sp] = val;
        phase_s[sp] = 2;
        int n2 = n_s[sp] / 2;
        n2 -= n2 % 8;
        lo_s[sp + 1] = lo_s[sp] + n2;
        n_s[sp + 1] = n_s[sp] - n2;
        phase_s[sp + 1] = 0;
        ++sp;
        break;
      }
      val = fadd_rn(left_s[sp], val);    // right half done: combine, keep delivering
    }
  }
}

struct PairwiseSerialLeaf {
  const float* a;
  FM_HD float operator()(int lo, int len) const { return pairwise_leaf(a + lo, len); }
};

FM_HD inline float pairwise_tree_f32(const float* a, int n) { return pairwise_walk(n, PairwiseSerialLeaf{a}); }

// The same sum split for parallel evaluation: list the leaves, sum them anywhere (each with pairwise_leaf's
// order), then fold the leaf sums along the tree.
struct PairwiseListLeaf {
  int* lo;
  int* len;
  int cap;
  int* count;
  FM_HD float operator()(int l, int n) const {
    if (*count < cap) {
      lo[*count] = l;
      len[*count] = n;
    }
    ++*count;
    return 0.f;
  }
};

// returns the number of leaves (entries beyond `cap` are not stored: the caller falls back to the serial form)
FM_HD inline int pairwise_list_leaves(int n, int* lo, int* len, int cap) {
  int count = 0;
  pairwise_walk(n, PairwiseListLeaf{lo, len, cap, &count});
  return count;
}

struct PairwiseFoldLeaf {
  const float* sums;
  int* next;
  FM_HD float operator()(int, int) const { return sums[(*next)++]; }
};

FM_HD inline float pairwise_fold_leaves(int n, const float* leaf_sums) {
  int next = 0;
  return pairwise_walk(n, PairwiseFoldLeaf{leaf_sums, &next});
}

// The same tree as a flat plan, for evaluation without a tree walk per sum: leaves left to right, and for
// each leaf the number of pending left partial sums that get folded in right after it ("left + value",
// repeated comb[l] times; a right child completes its parent, which may complete its own parent, ...).
// `st` is caller-provided stack storage (3 x 32 ints) so that device code can keep it out of local memory.
FM_HD inline int pairwise_plan(int n, int* lo, int* len, unsigned char* comb, int cap, int* st) {
  int* st_lo = st;
  int* st_n = st + 32;
  int* st_c = st + 64;
  int sp = 1, count = 0;
  st_lo[0] = 0;
  st_n[0] = n;
  st_c[0] = 0;
  while (sp > 0) {
    --sp;
    int l = st_lo[sp], m = st_n[sp], c = st_c[sp];
    while (m > 128) {                  // push the right half, descend into the left one
      int n2 = m / 2;
      n2 -= n2 % 8;
      st_lo[sp] = l + n2;
      st_n[sp] = m - n2;
      st_c[sp] = c + 1;
      ++sp;
      m = n2;
      c = 0;
    }
    if (count < cap) {
      lo[count] = l;
      len[count] = m;
      comb[count] = (unsigned char)c;
    }
    ++count;
  }
  return count;
}

// fold the leaf sums of a plan; `vstack` holds 32 values
template <class F>
FM_HD inline F pairwise_fold_plan(int count, const F* sums, const unsigned char* comb, F* vstack) {
  int sp = 0;
  for (int l = 0; l < count; ++l) {
    F v = sums[l];
    for (int k = 0; k < (int)comb[l]; ++k) v = fadd_rn(vstack[--sp], v);
    vstack[sp++] = v;
  }
  return vstack[0];
}

// numpy.add.reduce starts from the identity 0 and adds the pairwise result to it (so an empty or all -0
// input gives +0)
FM_HD inline float pairwise_sum_f32(const float* a, int n) { return fadd_rn(0.0f, pairwise_tree_f32(a, n)); }

}  // namespace fm
