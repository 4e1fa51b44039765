"""The 8-bit row slicing behind the sliced merge Gram (facemap_b200/csrc/slice_math.cuh, fm_slice_rows), compiled for
the CPU and checked for the properties the exactness argument rests on:

  * S0 + S1 + S2 reproduces X to 2^-25 of the bound 2^e_row of its row (round to nearest, unbiased), exactly for the
    entries within a factor two of it; the sum and every partial remainder are exact float32 operations;
  * every slice is an integer of at most 2^8 times its grid 2^(e_row - 8(d+1)): an exact TF32 operand (top 19 bits
    of the word), and dot products of 240 slice pairs are exact in float32;
  * the Gram assembled from the pair products  S0 S0^T + (S0 S1^T + S1 S0^T + S1 S1^T) + ...  in float64 equals the
    float64 Gram of X to ~1e-14 of its diagonal scale (what scripts/model_accuracy.py `passes` relies on)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim_slice") / "libshim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(HERE, "csrc", "host_math_shim.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.shim_slice_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def slice_with_shim(shim, X):
    X = np.ascontiguousarray(X, np.float32)
    rows, cols = X.shape
    S = [np.empty_like(X) for _ in range(3)]
    ex = np.empty(rows, np.int32)
    shim.shim_slice_rows(X.ctypes.data, rows, cols, S[0].ctypes.data, S[1].ctypes.data, S[2].ctypes.data, ex.ctypes.data)
    return S, ex


def graded_rows(rs, rows, cols):
    """rows whose norms span six decades, like stacked chunk components U_i S_i"""
    X = rs.randn(rows, cols) * np.logspace(3, -3, rows)[:, None]
    X[:, ::7] *= 1e-3                                     # entries far below the row maximum
    return X.astype(np.float32)


def test_slices_are_exact_and_cover_24_bits(shim):
    rs = np.random.RandomState(0)
    X = graded_rows(rs, 64, 1000)
    X[5] = 0                                              # an all-zero row
    X[6, :] = 0
    X[6, 17] = -3.5                                       # a single non-zero entry
    X[7] = np.float32(2.0) ** rs.randint(-30, 30, size=1000)          # exact powers of two (the row max IS 2^(ex-1))
    (S0, S1, S2), ex = slice_with_shim(shim, X)
    amax = np.abs(X).max(axis=1)
    nz = amax > 0
    assert np.all(amax[nz] < np.exp2(ex[nz].astype(np.float64))) and np.all(amax[nz] >= np.exp2(ex[nz] - 1.0))
    X64 = X.astype(np.float64)
    total = S0.astype(np.float64) + S1.astype(np.float64) + S2.astype(np.float64)
    grid = np.exp2(ex.astype(np.float64))[:, None]
    assert np.all(np.abs(X64 - total) <= grid * 2.0 ** -25)
    big = np.abs(X64) >= grid / 2                         # ulp(x) = 2^(ex-24): nothing is left over
    assert np.array_equal(total[big], X64[big])
    for d, S in enumerate((S0, S1, S2)):
        q = grid * 2.0 ** (-8 * (d + 1))
        k = S.astype(np.float64) / q
        assert np.array_equal(k, np.round(k)) and np.abs(k).max() <= (256 if d == 0 else 128), d
        assert np.array_equal(S.view(np.uint32) & np.uint32(0x1FFF), np.zeros_like(S, np.uint32)), "not a TF32 value"
    assert not S0[5].any() and not S1[5].any() and not S2[5].any()
    assert S0[6, 17] == np.float32(-3.5) and not S1[6].any()
    # unbiased: the representation error has no preferred sign relative to the entry
    err = (X64 - total)[~big & (X64 != 0)]
    assert abs(np.mean(np.sign(err * X64[~big & (X64 != 0)]))) < 0.05


def test_partial_dot_products_of_240_slice_pairs_are_exact_in_float32(shim):
    """what makes fm_gemm_tn on (S0, S0) with K partials <= 240 exact: a float32 accumulator, whatever its rounding,
    never rounds while |sum| * 2^-(e_i + e_j - 16) stays below 2^24"""
    rs = np.random.RandomState(1)
    X = graded_rows(rs, 48, 240 * 5)
    X[3] = 1000.0                                         # a constant row: every digit is the same, sums are maximal
    X[4] = -X[3]
    (S0, _, _), ex = slice_with_shim(shim, X)
    for k0 in range(0, X.shape[1], 240):
        a = S0[:, k0:k0 + 240]
        exact = a.astype(np.float64) @ a.astype(np.float64).T
        acc = np.zeros((48, 48), np.float32)
        for k in range(0, 240, 8):                        # 8-wide MMA steps, float32 accumulation
            acc = (acc.astype(np.float64) + a[:, k:k + 8].astype(np.float64) @ a[:, k:k + 8].astype(np.float64).T
                   ).astype(np.float32)
        assert np.array_equal(acc.astype(np.float64), exact)
        unit = np.exp2(ex[:, None].astype(np.float64) + ex[None, :] - 16)
        assert np.abs(exact / unit).max() < 2 ** 24


def test_gram_from_slice_pair_products(shim):
    rs = np.random.RandomState(2)
    X = graded_rows(rs, 96, 4000)
    (S0, S1, S2), _ = slice_with_shim(shim, X)
    P = [s.astype(np.float64) for s in (S0, S1, S2)]
    G = P[0] @ P[0].T
    for lo, hi in ((P[0], P[1]), (P[0], P[2]), (P[1], P[2])):       # the passes of svd.sliced_gram
        G += lo @ hi.T + hi @ lo.T + hi @ hi.T
    T = P[0] + P[1] + P[2]
    want = T @ T.T + P[2] @ P[2].T                        # S2 S2^T is counted twice (2^-32 of the diagonal scale)
    d = np.sqrt(np.diag(want))
    assert np.abs(G - want).max() <= 1e-13 * d.max() ** 2
    full = X.astype(np.float64) @ X.astype(np.float64).T
    # representation error: 2^-25 of the row bound per entry, unbiased -> averages down with sqrt(K)
    assert np.abs((G - full) / np.outer(d, d)).max() < 2e-8


def test_i8_digits(shim):
    """slice3_i8: s8 / u8 / u8 digits, 23 bits below the row bound, last digit to nearest with carries, saturation
    instead of a carry out of the leading digit"""
    shim.shim_slice_rows_i8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rs = np.random.RandomState(3)
    X = graded_rows(rs, 40, 3000)
    X[2] = 0
    X[3, :] = np.float32(1.0)                              # row bound 2, t = 64 exactly
    X[4, :5] = [np.nextafter(np.float32(4.0), np.float32(0)), -np.nextafter(np.float32(4.0), np.float32(0)),
                np.float32(-4.0 + 2.0 ** -20), np.float32(2.0 ** -30), np.float32(-(2.0 ** -30))]
    X[4, 5:] = 0
    rows, cols = X.shape
    q0 = np.empty((rows, cols), np.int8)
    q1 = np.empty((rows, cols), np.uint8)
    q2 = np.empty((rows, cols), np.uint8)
    ex = np.empty(rows, np.int32)
    shim.shim_slice_rows_i8(X.ctypes.data, rows, cols, q0.ctypes.data, q1.ctypes.data, q2.ctypes.data, ex.ctypes.data)
    scale = np.exp2(ex.astype(np.float64) - 7)[:, None]
    rec = scale * (q0.astype(np.float64) + q1 / 256.0 + q2 / 65536.0)
    X64 = X.astype(np.float64)
    err = (X64 - rec) / scale                              # in units of q0
    sat = (q0 == 127) & (q1 == 255) & (q2 == 255)
    assert np.abs(err[~sat]).max() <= 2.0 ** -17 + 1e-12   # half a unit of the last digit
    assert np.abs(err[sat]).max() <= 2.0 ** -16            # saturated entries lose at most one unit
    assert abs(err[np.abs(X64) > 0].mean()) < 2.0 ** -20   # unbiased
    assert not q0[2].any() and not q1[2].any() and not q2[2].any()
    assert (q0[3] == 64).all() and not q1[3].any()
    # the entry just below the bound 4 = 2^2 rounds up into a carry chain: saturated, not wrapped to -128
    assert (q0[4, 0], q1[4, 0], q2[4, 0]) == (127, 255, 255)
    assert q0[4, 1] == -128 and q1[4, 1] == 0 and q2[4, 1] == 0          # -3.9999998 rounds to -4 = -128 units: representable
    assert q0[4, 3] == 0 and q0[4, 4] in (-1, 0)           # tiny values: 0 or -1 + 255/256 + ... (floor digits)
    # reference implementation in NumPy (scripts/model_accuracy.py:slice_rows_i8), digit for digit
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "scripts"))
    from model_accuracy import slice_rows_i8

    (m0, m1, m2), mscale = slice_rows_i8(X)
    nz = np.abs(X).max(axis=1) > 0
    assert np.array_equal(m0[nz], q0[nz]) and np.array_equal(m1[nz], q1[nz]) and np.array_equal(m2[nz], q2[nz])
    assert np.array_equal(mscale[nz], scale[nz, 0])
