"""The scalar helpers the ROI kernels share with the host (facemap_b200/csrc/numpy_sum.cuh,
pupil_math.cuh), compiled for the CPU with g++ and pinned against NumPy / LAPACK:

  pairwise_sum_f32     == numpy.sum of a contiguous float32 array, bit for bit, for every length
  inv2x2               == numpy.linalg.inv to rounding
  median_from_hist     == numpy.median of the float32 values the histogram stands for, bit for bit
  sym2x2_eig_lapack    == numpy.linalg.eig (order of the eigenvalues, signs of the vectors)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "libshim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(HERE, "csrc", "host_math_shim.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.shim_pairwise_sum_f32.restype = C.c_float
    lib.shim_pairwise_sum_f32.argtypes = [C.c_void_p, C.c_int]
    lib.shim_pairwise_sum_split.restype = C.c_float
    lib.shim_pairwise_sum_split.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.shim_pairwise_leaf_count.argtypes = [C.c_int]
    lib.shim_pairwise_sum_plan.restype = C.c_float
    lib.shim_pairwise_sum_plan.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.shim_pairwise_sum_plan_f64.restype = C.c_double
    lib.shim_pairwise_sum_plan_f64.argtypes = [C.c_void_p, C.c_int]
    lib.shim_inv2x2.argtypes = [C.c_double] * 3 + [C.c_void_p]
    lib.shim_median_from_hist.restype = C.c_float
    lib.shim_median_from_hist.argtypes = [C.c_void_p, C.c_float]
    lib.shim_sym2x2_eig.argtypes = [C.c_double] * 3 + [C.c_void_p]
    return lib


def test_pairwise_sum_matches_numpy_bitwise(shim):
    rs = np.random.RandomState(0)
    sizes = list(range(0, 300)) + [511, 512, 513, 1000, 1023, 1024, 1025, 4095, 4104, 8191, 8192, 8193,
                                    10007, 16384, 20001, 40000, 65537, 100003, 250000]
    for n in sizes:
        for scale in (1.0, 1e-3, 300.0):
            a = (rs.rand(n) * scale).astype(np.float32)
            if n % 3 == 0:
                a[rs.rand(n) < 0.5] = 0          # the pupil weights are half zeros after thresholding
            got = shim.shim_pairwise_sum_f32(a.ctypes.data, n)
            want = a.sum()
            assert np.float32(got).tobytes() == np.float32(want).tobytes(), (n, scale, got, want)
            # the leaf-list / fold split used for the block-parallel evaluation (cap 1024 leaves, serial beyond)
            got = shim.shim_pairwise_sum_split(a.ctypes.data, n, 1024)
            assert np.float32(got).tobytes() == np.float32(want).tobytes(), ("split", n, scale, got, want)
            got = shim.shim_pairwise_sum_plan(a.ctypes.data, n, 512)          # flat plan (what the pupil kernel runs)
            assert np.float32(got).tobytes() == np.float32(want).tobytes(), ("plan", n, scale, got, want)
    # leaves hold 64..128 elements once the array is split at all: 1024 leaves cover >= 65536 elements
    for n in (129, 1000, 32768, 65536):
        assert shim.shim_pairwise_leaf_count(n) <= max(1, n // 64)


def test_pairwise_sum_f64_matches_numpy_bitwise(shim):
    rs = np.random.RandomState(5)
    for n in [1, 7, 8, 9, 127, 128, 129, 332, 540, 1000, 4096, 19200, 34080, 100003, 1310720]:
        a = rs.rand(n) * 3.0 + rs.rand(n) * 1e-9
        got = shim.shim_pairwise_sum_plan_f64(a.ctypes.data, n)
        want = a.sum()
        assert np.float64(got).tobytes() == np.float64(want).tobytes(), (n, got, want)


def test_inv2x2_matches_numpy(shim):
    rs = np.random.RandomState(1)
    out = np.zeros(4)
    for _ in range(2000):
        A = rs.randn(2, 2) * rs.choice([0.1, 1, 30])
        S = A @ A.T + 0.01 * np.eye(2)
        if rs.rand() < 0.2:
            S[0, 1] = S[1, 0] = S[0, 0] * 1.5 * np.sign(rs.randn())     # forces the pivoted branch (|b| > |a|)
            S[1, 1] = 3.5 * S[0, 0]
        shim.shim_inv2x2(S[0, 0], S[0, 1], S[1, 1], out.ctypes.data)
        ref = np.linalg.inv(S)
        assert np.allclose(out.reshape(2, 2), ref, rtol=1e-12, atol=0), (S, out, ref)


def test_median_from_hist_matches_numpy_bitwise(shim):
    rs = np.random.RandomState(2)
    for trial in range(400):
        off = np.float32(255.0 - rs.choice([60.0, 90.27, 140.25, 200.5, 254.9]))
        hist = np.zeros(257, np.uint32)
        k = rs.randint(0, 6)
        vals = rs.randint(0, 256, size=k) if k else np.array([], int)
        for v in vals:
            if np.maximum(np.float32(0), np.float32(255 - v) - off) > 0:
                hist[v] += rs.randint(1, 5)
        hist[256] = rs.randint(0, 4)
        parts = [np.zeros(hist[256], np.float32)]
        for v in range(256):
            parts.append(np.full(hist[v], np.maximum(np.float32(0), np.float32(255 - v) - off), np.float32))
        x = np.concatenate(parts)
        got = np.float32(shim.shim_median_from_hist(hist.ctypes.data, off))
        if x.size == 0:
            assert np.isnan(got)
        else:
            want = np.median(x)
            assert got.tobytes() == np.float32(want).tobytes(), (trial, hist.nonzero(), got, want)


def test_sym2x2_eig_matches_lapack_order_and_signs(shim):
    rs = np.random.RandomState(3)
    out = np.zeros(6)
    for i in range(20000):
        A = rs.randn(2, 2) * rs.choice([0.1, 1, 10, 100])
        S = A @ A.T + 0.01 * np.eye(2)
        k = i % 9
        if k < 6:
            S[0, 1] = S[1, 0] = S[0, 1] * 10.0 ** (-8 - 2 * k)      # down to the deflation regime
        if i % 13 == 0:
            S[0, 1] = S[1, 0] = 0.0
        w, u = np.linalg.eig(S)
        shim.shim_sym2x2_eig(S[0, 0], S[0, 1], S[1, 1], out.ctypes.data)
        assert np.allclose(out[:2], w, rtol=1e-13, atol=0), (S, out, w)
        assert np.allclose(out[2:].reshape(2, 2), u, rtol=0, atol=1e-13), (S, out, u)
