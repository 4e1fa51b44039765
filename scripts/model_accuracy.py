"""CPU model of the arithmetic of the CUDA SVD path (no GPU needed): which rounding step sets the error floor of the
tail singular values, and what an exact-integer Gram would buy.

The oracle pipeline (oracle/facemap_oracle.py, `run_oracle`) is run with its `svd_centered` swapped for a NumPy
model of what facemap_b200/svd.py launches:

    centred Gram  G = A A^T - p m m^T         fm_gemm_tn (3xTF32, chained) + fm_gram_assemble (float64)
    eigensolve                                 float64 (cuSOLVER)
    W -> float32, sign rule                    fm_topk_components
    left vectors  Y = A^T W - 1 (m^T W)        fm_gemm_tn + fm_splitk_reduce (float64), stored float32

`emu_gemm` restates the tensor-core arithmetic measured in round 1 (scripts/probe_trunc.py):
kind::tf32 reads the top 19 bits of each operand word; per 8-wide K step the kernel issues lo*hi, hi*lo, hi*hi;
every MMA adds into the fp32 TMEM accumulator with TRUNCATION; a chain of 128 K is drained into float32 registers
(round to nearest); split-K partials (<= 320 K for Gram matrices, <= 256 K for left vectors) are summed in float64.

    python scripts/model_accuracy.py [--H 96 --W 128 --N 2600 --sbin 2] [--variants current,chunk_exact,...]

Variants (which GEMMs are modelled as tensor-core arithmetic, which as exact float64 on the float32 operands):
    current       everything emulated (the round-1 product configuration)
    chunk_exact   chunk Gram exact (integer slicing), the rest emulated
    grams_exact   chunk and merge Gram exact, left vectors emulated
    all_exact     every GEMM exact; only the float32 storage of W, Y and U remains
    ozaki3[_o]    Grams from three 8-bit row slices, every pair product with d + e <= o exact (default all)
    i8 / i8m      Grams from s8/u8/u8 floor digits with exact integer products (the kind::i8 plan), both levels / merge
    passes / passesm   the four-pass scheme that fits the existing kernels (see Model.gemm), both levels / merge only
    sliced2       both Grams as  top-8-bit row slice (exact, one MMA per K step)  +  one 3-MMA pass on the remainder
    sliced3       the same with a second remainder pass;  sliced2m / sliced3m: merge Gram only
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facemap_b200.synth import SynthVideo  # noqa: E402
from oracle import facemap_oracle as orc  # noqa: E402

KCHAIN_GRAM, KCHAIN_LEFT, BK, CH = 320, 256, 16, 8


def trunc19(x32):
    """what kind::tf32 sees of a float32 word: sign, exponent, top 10 mantissa bits"""
    return (np.ascontiguousarray(x32, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def add_trunc(acc32, addend64):
    """fp32 accumulator + exact addend, rounded TOWARD ZERO to float32 (the TMEM accumulation)"""
    s = acc32.astype(np.float64) + addend64
    y = s.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(s)
    if over.any():
        y[over] = np.nextafter(y[over], np.float32(0))
    return y


def emu_products(pairs, kpart):
    """sum over `pairs` [(A [M,K], B [N,K]) float64 planes already holding what the MMA reads] of A @ B^T in the
    modelled accumulation: per 8-wide K step one truncating add per pair (in list order), TMEM chains of CH k-blocks
    drained into float32 registers, split-K partials of <= kpart summed in float64."""
    M, K = pairs[0][0].shape
    N = pairs[0][1].shape[0]
    KB = -(-K // BK)
    ksplit = max(1, min(-(-K // kpart), KB))
    C = np.zeros((M, N), np.float64)
    for z in range(ksplit):
        kb0, kb1 = (z * KB) // ksplit, ((z + 1) * KB) // ksplit
        reg = np.zeros((M, N), np.float32)
        for c0 in range(kb0, kb1, CH):
            acc = np.zeros((M, N), np.float32)
            for kb in range(c0, min(kb1, c0 + CH)):
                for k0 in range(kb * BK, min(K, (kb + 1) * BK), 8):
                    s = slice(k0, min(K, k0 + 8))
                    for (P, Q) in pairs:
                        acc = add_trunc(acc, P[:, s] @ Q[:, s].T)
            reg = reg + acc                                   # float32, round to nearest
        C += reg.astype(np.float64)
    return C


def emu_gemm(A, B, kpart, upper_only=False):
    """C = A @ B^T for float32 A [M,K], B [N,K] as the product kernel evaluates it (lo*hi, hi*lo, hi*hi)."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    Ah, Bh = trunc19(A), trunc19(B)
    Al, Bl = trunc19(A - Ah), trunc19(B - Bh)
    Ah, Al, Bh, Bl = (x.astype(np.float64) for x in (Ah, Al, Bh, Bl))
    return emu_products([(Al, Bh), (Ah, Bl), (Ah, Bh)], kpart)


def slice_rows(A, bits=8):
    """A [n, K] float32 -> (A0, R): A0 = every entry rounded to the nearest multiple of 2^(e_row - bits), 2^e_row
    bounding the row (exact in TF32; products of two such rows sum exactly in fp32 over ~2^(24 - 2 bits) terms),
    R = A - A0 exactly (float32).  The arithmetic of facemap_b200/csrc/slice_math.cuh."""
    A = np.ascontiguousarray(A, np.float32)
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax.astype(np.float64), 1e-300))) + 1, 0.0)        # |a| < 2^e
    q = np.exp2(e - bits).astype(np.float32)[:, None]
    A0 = (np.rint(A / q) * q).astype(np.float32)
    R = (A - A0).astype(np.float32)
    assert np.array_equal(A0.astype(np.float64) + R.astype(np.float64), A.astype(np.float64))
    return A0, R


def slice_rows_like(R, A, bits):
    """next slice of the remainder R on the row grid of A: multiples of 2^(e_row(A) - bits)"""
    amax = np.abs(np.ascontiguousarray(A, np.float32)).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax.astype(np.float64), 1e-300))) + 1, 0.0)
    q = np.exp2(e - bits).astype(np.float32)[:, None]
    S = (np.rint(R / q) * q).astype(np.float32)
    return S, (R - S).astype(np.float32)


def slice_rows_i8(A):
    """A [n, K] float32 -> ([Q0, Q1, Q2] float64 integer planes, scale [n]): A ~ scale * (Q0 + Q1/2^8 + Q2/2^16) with
    scale = 2^(e_row - 7), |A| < 2^e_row; Q0 = floor in [-128, 127], Q1 = floor in [0, 255], Q2 to nearest in [0, 255]
    with carries (23 bits below the row bound, unbiased).  The arithmetic of csrc/slice_math.cuh:slice3_i8."""
    A = np.ascontiguousarray(A, np.float32)
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax.astype(np.float64), 1e-300))) + 1, -100.0)
    scale = np.exp2(e - 7)
    t = A.astype(np.float64) / scale[:, None]                # exact: power-of-two scaling, |t| < 128
    q0 = np.floor(t)
    r = (t - q0) * 256.0
    q1 = np.floor(r)
    r = (r - q1) * 256.0
    q2 = np.rint(r)                                          # last digit to nearest (unbiased), carries propagated
    c = q2 == 256
    q2[c] = 0
    q1[c] += 1
    c = q1 == 256
    q1[c] = 0
    q0[c] += 1
    c = q0 == 128                                            # only for |a| within 2^-24 of the row bound: saturate
    q0[c], q1[c], q2[c] = 127, 255, 255
    assert q0.min() >= -128 and q0.max() <= 127 and q1.min() >= 0 and q1.max() <= 255 and 0 <= q2.min() and q2.max() <= 255
    return [q0, q1, q2], scale


def sliced_gram(A, passes, bits=8):
    """Gram A A^T as: exact top-slice product (single MMA per K step, partials of 2^(24-2 bits) K) + the 3-MMA
    kernel on (hi = remainder, lo = top slice), optionally once more on the remainder's remainder."""
    A0, R = slice_rows(A, bits)
    A0d = A0.astype(np.float64)
    G = emu_products([(A0d, A0d)], 240)
    exact = A0d @ A0d.T
    assert np.array_equal(G, exact), "top-slice Gram is not exact in the modelled arithmetic"
    R1 = trunc19(R)
    G += emu_products([(A0d, R1.astype(np.float64)), (R1.astype(np.float64), A0d),
                       (R1.astype(np.float64), R1.astype(np.float64))], 4096)
    if passes >= 3:
        R2 = trunc19(R - R1).astype(np.float64)
        G += emu_products([(A0d, R2), (R2, A0d), (R2, R2)], 4096)
    return G


def exact_gemm(A, B, *_a, **_k):
    return np.asarray(A, np.float64) @ np.asarray(B, np.float64).T


class Model:
    def __init__(self, variant):
        self.variant = variant
        self.calls = 0

    def gemm(self, role, A, B, kpart):
        exact = {"current": (), "chunk_exact": ("chunk.gram",), "grams_exact": ("chunk.gram", "merge.gram"),
                 "all_exact": ("chunk.gram", "merge.gram", "chunk.left", "merge.left")}.get(self.variant, ())
        if self.variant.startswith("passes") and role.endswith(".gram") and \
                (role == "merge.gram" or not self.variant.endswith("m")):
            # the scheme that fits the existing kernels: three 8-bit row slices Y0, Y1, Y2;
            #   pass A  Y0 Y0^T                      one MMA per K step, partials of 256 K: exact
            #   pass B  Y0 Y1^T + Y1 Y0^T + Y1 Y1^T  the 3-MMA kernel with (lo, hi) = (Y0, Y1): inexact at 2^-8 scale
            #   pass C  the same with (Y0, Y2);  pass D with (Y1, Y2)  (Y2 Y2^T counted twice: 2^-32, harmless)
            A0, R = slice_rows(A, 8)
            A1, R = slice_rows_like(R, A, 16)
            A2, R = slice_rows_like(R, A, 24)
            Y0, Y1, Y2 = (x.astype(np.float64) for x in (A0, A1, A2))
            G = emu_products([(Y0, Y0)], 240)
            assert np.array_equal(G, Y0 @ Y0.T)
            for lo, hi in ((Y0, Y1), (Y0, Y2), (Y1, Y2)):
                G += emu_products([(lo, hi), (hi, lo), (hi, hi)], 4096)
            return G
        if self.variant.startswith("i8") and role.endswith(".gram") and \
                (role == "merge.gram" or not self.variant.endswith("m")):
            # integer digits for tcgen05 kind::i8 (fm_slice_rows_i8 + fm_gemm_i8_tn, exact int32 accumulation):
            # y ~ 2^(e-7) (q0 + q1 2^-8 + q2 2^-16), q0 = floor in [-128, 127] (s8), q1, q2 in [0, 255] (u8)
            Q, scale = slice_rows_i8(A)
            G = np.zeros((A.shape[0], A.shape[0]))
            for d in range(3):
                for e in range(3):
                    G += (Q[d] @ Q[e].T) * 2.0 ** (-8 * (d + e))
            return G * np.outer(scale, scale)
        if self.variant.startswith("ozaki") and role.endswith(".gram"):
            # ozaki3 / ozaki4: every pair product of 3 (4) 8-bit row slices exact = the exact Gram of the rows cut to
            # 24 (32) bits below their row maximum
            A0, R = slice_rows(A, 8)
            A1, R = slice_rows_like(R, A, 16)
            A2, R = slice_rows_like(R, A, 24)
            planes = [x.astype(np.float64) for x in (A0, A1, A2)]
            if self.variant == "ozaki4":
                planes.append(slice_rows_like(R, A, 32)[0].astype(np.float64))
            # ozaki3_2 / ozaki3_3: only the pair products with d + e <= 2 (3) are formed
            order = int(self.variant.split("_")[1]) if "_" in self.variant else 2 * len(planes)
            G = np.zeros((A.shape[0], A.shape[0]))
            for d, P in enumerate(planes):
                for e, Q in enumerate(planes):
                    if d + e <= order:
                        G += P @ Q.T
            return G
        if self.variant.startswith("sliced") and role.endswith(".gram"):
            # sliced2 / sliced3: both levels; sliced2m / sliced3m: merge level only
            if not self.variant.endswith("m") or role == "merge.gram":
                return sliced_gram(A, int(self.variant[6]))
        return (exact_gemm if role in exact else emu_gemm)(A, B, kpart)

    def svd_centered(self, X, k, svd=None):
        """stand-in for oracle.svd_centered(X [p, n], k): the Gram route of facemap_b200/svd.py."""
        self.calls += 1
        X = np.asarray(X, np.float32)
        p, n = X.shape
        level = "merge" if n > 1000 or self.in_merge else "chunk"
        A = np.ascontiguousarray(X.T)                                   # [n, p]: rows = frames / stacked components
        m = A.astype(np.float64).mean(axis=1)
        G = self.gemm(level + ".gram", A, A, KCHAIN_GRAM) - p * np.outer(m, m)
        G = 0.5 * (G + G.T)
        lam, W = np.linalg.eigh(G)
        lam, W = lam[::-1][:k], W[:, ::-1][:, :k]
        S = np.sqrt(np.maximum(lam, 0.0))
        idx = np.argmax(np.abs(W), axis=0)
        sg = np.sign(W[idx, np.arange(k)])
        W = W * np.where(sg == 0, 1.0, sg)[None, :]
        Wt = np.ascontiguousarray(W.T.astype(np.float32))               # [k, n]
        bias = Wt.astype(np.float64) @ m                                # [k]
        Yt = self.gemm(level + ".left", Wt, np.ascontiguousarray(A.T), KCHAIN_LEFT) - bias[:, None]      # [k, p]
        if level == "chunk":
            Y = Yt.T.astype(np.float32).astype(np.float64)              # U S, stored float32
            return Y / S[None, :], S, None                              # the caller multiplies by S again
        U = (Yt / S[:, None]).T.astype(np.float32)
        return U, S.astype(np.float32), None


def run_variant(video, sbin, variant, movSVD):
    model = Model(variant)
    model.in_merge = False
    saved = orc.svd_centered
    nsegs = orc.stage2_plan(video.shape[0])[1]
    per_level = (1 + int(movSVD)) * nsegs

    def patched(X, k, svd=None):
        model.in_merge = model.calls >= per_level
        return model.svd_centered(X, k)

    orc.svd_centered = patched
    try:
        return orc.run_oracle([video], sbin=sbin, motSVD=True, movSVD=movSVD, svd="model")
    finally:
        orc.svd_centered = saved


def report(name, o, ref, movSVD):
    out = {}
    for sk in ("motSv",) + (("movSv",) if movSVD else ()):
        S, Sr = o[sk].astype(np.float64), ref[sk].astype(np.float64)
        rel = np.abs(S - Sr) / Sr
        out[sk] = [float(rel[:k].max()) for k in (10, 50, 100, 250, 500)]
        print(f"{name:12s} {sk}: max rel err over [0:10] [0:50] [0:100] [0:250] [0:500] = "
              + "  ".join(f"{x:.1e}" for x in out[sk]), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=96)
    ap.add_argument("--W", type=int, default=128)
    ap.add_argument("--N", type=int, default=2600)
    ap.add_argument("--sbin", type=int, default=2)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--mov", action="store_true")
    ap.add_argument("--variants", default="current,chunk_exact,grams_exact,all_exact")
    a = ap.parse_args()
    v = SynthVideo(a.H, a.W, a.N, seed=a.seed).all_frames()
    t = time.time()
    ref = orc.run_oracle([v], sbin=a.sbin, motSVD=True, movSVD=a.mov, svd="exact")
    Sr = ref["motSv"].astype(np.float64)
    print(f"oracle (exact float64 svdecon): {time.time() - t:.1f} s;  S_k/S_0 at 10/50/100/250/499 = "
          + " ".join(f"{Sr[i] / Sr[0]:.1e}" for i in (10, 50, 100, 250, 499)), flush=True)
    for name in a.variants.split(","):
        t = time.time()
        o = run_variant(v, a.sbin, name, a.mov)
        report(name, o, ref, a.mov)
        print(f"             ({time.time() - t:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
