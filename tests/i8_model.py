"""NumPy restatement of the integer projection path (csrc/project_i8.cu, fm_bin_motion_planes), test infrastructure only.

Every function follows the kernel's arithmetic step by step so that the GPU result can be compared bit for bit:
  planes_of          value = 256 * hi + lo
  class_sums         one int64 sum per weight class j = d + e of the ND x 3 plane products (the TMEM accumulators;
                     `chain` splits K like the kernel and checks that every chain fits an int32)
  project_i8_model   Horner combination in int64, one float64 multiply by alpha * 2^(ex - 23), one float64 subtraction
                     of the bias, one rounding to float32
The mask digits come from scripts/model_accuracy.py:slice_rows_i8, which tests/test_slice_math.py pins digit for digit
against csrc/slice_math.cuh."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
from model_accuracy import slice_rows_i8  # noqa: E402

KCHAIN = 16384


def planes_of(D):
    D = np.asarray(D, np.int64)
    assert D.min() >= 0 and D.max() < 65536
    return (D >> 8).astype(np.uint8), (D & 255).astype(np.uint8)


def mask_digits(U):
    """U [k, p] float32 -> (Q0 int8, Q1 uint8, Q2 uint8, exps int32) like kernels.slice_rows_i8"""
    (q0, q1, q2), scale = slice_rows_i8(U)
    exps = np.rint(np.log2(scale)).astype(np.int32) + 7
    return q0.astype(np.int8), q1.astype(np.uint8), q2.astype(np.uint8), exps


def class_sums(D_planes, Q, chain=KCHAIN):
    """D_planes: most significant first, 1 or 2 uint8 [M, K]; Q: (Q0, Q1, Q2) [N, K].  Returns int64 [chains, NC, M, N]."""
    nd, K = len(D_planes), D_planes[0].shape[1]
    nc = nd + 2
    out = []
    for k0 in range(0, K, chain):
        k1 = min(K, k0 + chain)
        cls = [0] * nc
        for d, Dp in enumerate(D_planes):
            for e, Qe in enumerate(Q):
                # float64 BLAS is exact here: |partial sums| < 2^53
                prod = Dp[:, k0:k1].astype(np.float64) @ Qe[:, k0:k1].astype(np.float64).T
                cls[d + e] = cls[d + e] + prod
        cls = np.stack([np.asarray(c) for c in cls]).astype(np.int64)
        assert np.abs(cls).max() < 2 ** 31, "a chain overflows the int32 accumulator"
        out.append(cls)
    return np.stack(out)


def project_i8_model(D, U, alpha, avg=None, chain=KCHAIN, bias=None):
    """float32 V [M, N] exactly as fm_project_i8 forms it from integer D [M, K] and float32 masks U [N, K].
    bias: the float64 vector handed to the kernel (default: bias_of(U, avg) when avg is given)."""
    D = np.asarray(D, np.int64)
    hi, lo = planes_of(D)
    planes = [lo] if D.max() < 256 else [hi, lo]
    q0, q1, q2, exps = mask_digits(U)
    cls = class_sums(planes, (q0, q1, q2), chain)
    nc = cls.shape[1]
    acc = np.zeros(cls.shape[2:], np.int64)
    for c in range(cls.shape[0]):
        for j in range(nc):
            acc += cls[c, j] * (1 << (8 * (nc - 1 - j)))
    sc = np.float64(alpha) * np.exp2(exps.astype(np.float64) - 23)
    v = acc.astype(np.float64) * sc[None, :]
    if bias is None and avg is not None:
        bias = bias_of(U, avg)
    if bias is not None:
        v = v - np.asarray(bias, np.float64)[None, :]
    return v.astype(np.float32)


def bias_of(U, avg):
    return U.astype(np.float64) @ np.asarray(avg, np.float64)


def block_sums(frames, sbin):
    """S [T, Lyb * Lxb] int64: sums of sbin x sbin blocks of uint8 frames [T, H, W] (facemap/process.py:34-43 before
    the division)"""
    T, H, W = frames.shape
    Lyb, Lxb = H // sbin, W // sbin
    f = frames[:, :Lyb * sbin, :Lxb * sbin].astype(np.int64)
    return f.reshape(T, Lyb, sbin, Lxb, sbin).sum(axis=(2, 4)).reshape(T, Lyb * Lxb)
