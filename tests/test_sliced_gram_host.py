"""svd.sliced_gram's orchestration (which planes go into which pass, K partition, the stack of partials) on the CPU,
with NumPy stand-ins for the kernels that restate their contracts:

  fm_slice_rows     the arithmetic of csrc/slice_math.cuh (pinned in tests/test_slice_math.py)
  fm_gemm_tn        impl 0 with both lo planes: per K partial (16-wide k-blocks dealt as in the kernel) the float32
                    value of  lo.hi^T + hi.lo^T + hi.hi^T  on the top 19 bits of every word; tiles entirely below the
                    diagonal are left untouched (poisoned with NaN here)
  fm_gram_assemble  float64 sum of the partials, upper triangle mirrored, minus p m m^T

The stand-in REFUSES to round in pass A: it asserts that every partial of S0 S0^T is exactly representable in float32,
which is the property the whole scheme rests on (a truncating accumulator never rounds an exact sum)."""
import math

import numpy as np
import torch

from facemap_b200 import svd


def _trunc19(x):
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


class FakeKernels:
    def __init__(self):
        self.calls = []

    @staticmethod
    def round_up(x, m):
        return (int(x) + m - 1) // m * m

    @staticmethod
    def row_means(X):
        return torch.from_numpy(X.numpy().astype(np.float64).mean(axis=1))

    @staticmethod
    def slice_rows(X, out):
        A = X.numpy()
        amax = np.abs(A).max(axis=1)
        e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax, 1e-300))) + 1, -100.0)
        R = A.astype(np.float32)
        for d, pl in enumerate(out):
            q = np.exp2(e - 8 * (d + 1)).astype(np.float32)[:, None]
            S = (np.rint(R / q) * q).astype(np.float32)
            R = (R - S).astype(np.float32)
            pl.copy_(torch.from_numpy(S))
        return tuple(out)

    def gemm_tn(self, A, B, C_out=None, *, ksplit=1, workspace=None, upper_only=False, impl=0, B_lo=None, A_lo=None,
                rscale=None, rbias=None):
        assert upper_only and A is B and A_lo is B_lo
        assert (impl == 0 and A_lo is not None) or (impl == 10 and A_lo is None)
        hi = _trunc19(A.numpy()).astype(np.float64)
        lo = _trunc19(A_lo.numpy()).astype(np.float64) if A_lo is not None else np.zeros_like(hi)
        n, K = hi.shape
        assert n > 128, "the chain kernel serves outputs wider than 128 only"
        KB = -(-K // 16)
        ksplit = min(ksplit, -(-K // 32))
        outs = [C_out] if ksplit == 1 else [workspace[z] for z in range(ksplit)]
        if ksplit > 1:
            assert workspace.shape[0] == ksplit and workspace.stride(0) == n * workspace.stride(1)
        exact_pass = not lo.any()
        for z, dst in enumerate(outs):
            k0, k1 = 16 * ((z * KB) // ksplit), min(K, 16 * (((z + 1) * KB) // ksplit))
            h, l = hi[:, k0:k1], lo[:, k0:k1]
            P = l @ h.T + h @ l.T + h @ h.T
            P32 = P.astype(np.float32)
            if exact_pass:
                assert k1 - k0 <= 240 and np.array_equal(P32.astype(np.float64), P), "pass A partial would round"
            full = np.full((n, dst.shape[1]), np.nan, np.float32)
            for i0 in range(0, n, 128):
                for j0 in range(0, n, 256):
                    if j0 + 255 >= i0:                                  # the kernel skips tiles below the diagonal
                        full[i0:i0 + 128, j0:min(n, j0 + 256)] = P32[i0:i0 + 128, j0:min(n, j0 + 256)]
            dst.copy_(torch.from_numpy(full))
        self.calls.append((ksplit, exact_pass))

    @staticmethod
    def gram_assemble(part, ksplit, n, mean, ncols, upper_only, out=None):
        P = part[:ksplit].numpy().astype(np.float64)[:, :, :n]
        iu = np.triu_indices(n)
        assert upper_only and np.isfinite(P[:, iu[0], iu[1]]).all(), "an upper-triangle entry was never written"
        G = np.triu(P.sum(axis=0))
        G = G + np.triu(G, 1).T
        m = mean.numpy()
        G -= ncols * np.outer(m, m)
        G = torch.from_numpy(G)
        if out is not None:
            out.copy_(G)
            return out
        return G


class FakeI8Kernels(FakeKernels):
    """stand-ins for fm_slice_rows_i8 / fm_gemm_i8_tn / fm_gram_assemble_i8 (contracts in include/facemap_b200.h)"""

    @staticmethod
    def slice_rows_i8(X, out):
        import os
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
        from model_accuracy import slice_rows_i8

        (q0, q1, q2), scale = slice_rows_i8(X.numpy())
        out[0].copy_(torch.from_numpy(q0.astype(np.int8)))
        out[1].copy_(torch.from_numpy(q1.astype(np.uint8)))
        out[2].copy_(torch.from_numpy(q2.astype(np.uint8)))
        assert out[0].stride(0) % 16 == 0 and out[0].stride(0) == out[1].stride(0) == out[2].stride(0)
        return out[0], out[1], out[2], torch.from_numpy((np.log2(scale) + 7).astype(np.int32))

    def gemm_i8_tn(self, A, B, C_out=None, *, ksplit=1, workspace=None, upper_only=False):
        a, b = A.numpy().astype(np.int64), B.numpy().astype(np.int64)
        n, K = a.shape
        KB = -(-K // 64)
        ksplit = min(ksplit, KB)
        assert -(-KB // ksplit) * 64 <= 32768, "int32 accumulator would overflow"
        outs = [C_out] if ksplit == 1 else [workspace[z] for z in range(ksplit)]
        for z, dst in enumerate(outs):
            k0, k1 = 64 * ((z * KB) // ksplit), min(K, 64 * (((z + 1) * KB) // ksplit))
            Pz = a[:, k0:k1] @ b[:, k0:k1].T
            assert np.abs(Pz).max() < 2 ** 31
            full = np.full((n, dst.shape[1]), -(2 ** 31), np.int64)          # poison: never-written entries
            for i0 in range(0, n, 128):
                for j0 in range(0, b.shape[0], 256):
                    if not upper_only or j0 + 255 >= i0:
                        full[i0:i0 + 128, j0:min(b.shape[0], j0 + 256)] = Pz[i0:i0 + 128, j0:min(b.shape[0], j0 + 256)]
            dst.copy_(torch.from_numpy(full.astype(np.int32)))
        self.calls.append(("i8", ksplit, upper_only, str(A.dtype), str(B.dtype)))

    @staticmethod
    def gram_assemble_i8(P, ksplit, n, exps, mean, ncols, out=None, rowside=None):
        S = P[:, :ksplit].numpy().astype(np.int64).sum(axis=1)[:, :, :n]     # [6, n, n]
        iu = np.triu_indices(n)
        poison = -(2 ** 31) * ksplit
        assert (S[:3][:, iu[0], iu[1]] != poison).all() and (S[3:] != poison).all()
        br = (S[0] + (S[3] + S[3].T) / 2.0 ** 8 + (S[1] + S[4] + S[4].T) / 2.0 ** 16 + (S[5] + S[5].T) / 2.0 ** 24
              + S[2] / 2.0 ** 32)
        e = exps.numpy().astype(np.float64)
        G = np.triu(br * np.exp2(e[:, None] + e[None, :] - 14))
        G = G + np.triu(G, 1).T
        m = mean.numpy()
        if rowside is None:
            G = torch.from_numpy(G - ncols * np.outer(m, m))
        else:                                                   # row-side form: - q 1^T - 1 q^T + mu.mu
            q, mu = rowside[0].numpy(), rowside[1].numpy()
            G = torch.from_numpy(G - (q[:, None] + q[None, :]) + float(mu @ mu))
        if out is not None:
            out.copy_(G)
            return out
        return G


def test_i8_gram_matches_float64_gram(monkeypatch):
    rs = np.random.RandomState(1)
    n, p = 300, 1000
    base = rs.randn(40, p)
    Y = ((rs.randn(n, 40) @ base) * np.logspace(2, -3, n)[:, None] + 1e-4 * rs.randn(n, p) + 0.3).astype(np.float32)
    Y64 = Y.astype(np.float64)
    m = Y64.mean(axis=1)
    want = Y64 @ Y64.T - p * np.outer(m, m)
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    for kmax, ks_want in ((32768, 1), (256, 4)):                        # one partial; forced split-K
        fake = FakeI8Kernels()
        monkeypatch.setattr(svd, "K", fake)
        monkeypatch.setattr(svd, "I8_KMAX", kmax)
        ws = _workspace()
        ws.gram_mode = "i8"
        X = ws.matrix("Y", n, p)
        X.copy_(torch.from_numpy(Y))
        mean, G = svd.centred_gram(ws, X, ws.timer, "stage2.chunk")
        err = np.abs(G.numpy() - want) / np.outer(d, d)
        # 2^-24 of the row bound per entry, unbiased, averaged over K = 1000
        assert err.max() < 5e-8 and np.median(err) < 5e-9, (err.max(), np.median(err))
        assert np.array_equal(G.numpy(), G.numpy().T) and np.array_equal(mean.numpy(), m)
        assert [c[1] for c in fake.calls] == [ks_want] * 6
        assert [c[2] for c in fake.calls] == [True, True, True, False, False, False]
        assert [c[3:] for c in fake.calls][3] == ("torch.int8", "torch.uint8")


def test_i8_gram_rowside_matches_float64(monkeypatch):
    """the merge of a small ROI: Y [p, c] with fewer pixels than stacked components, centred over its columns"""
    rs = np.random.RandomState(2)
    p, c = 120, 700
    Y = (rs.randn(p, 30) @ rs.randn(30, c) * np.logspace(1, -2, c)[None, :] + 0.2).astype(np.float32)
    Y64 = Y.astype(np.float64)
    mu = Y64.mean(axis=0)
    Yc = Y64 - mu[None, :]
    want = Yc @ Yc.T
    fake = FakeI8Kernels()
    fake.row_dot = staticmethod(lambda M, w: torch.from_numpy(M.numpy().astype(np.float64) @ w.numpy()))
    monkeypatch.setattr(svd, "K", fake)
    ws = _workspace()
    ws.gram_mode = "i8"
    X = ws.matrix("Y", p, c)
    X.copy_(torch.from_numpy(Y))
    q, G = svd.i8_gram(ws, X, ws.timer, "stage2.roi_merge", rowside_mu=torch.from_numpy(mu))
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    assert np.allclose(q.numpy(), Y64 @ mu, rtol=1e-13)
    # digits carry 2^-24 of each ROW's largest entry; with the columns graded over three decades a row's norm sits in
    # a few of them, so the bound relative to the row norms is looser than for the chunk matrices above
    assert (np.abs(G.numpy() - want) / np.outer(d, d)).max() < 5e-7
    assert np.array_equal(G.numpy(), G.numpy().T)


def _workspace():
    ws = svd.SvdWorkspace.__new__(svd.SvdWorkspace)
    ws.device, ws._ws, ws._bufs = torch.device("cpu"), None, {}
    ws.timer = svd.StageTimer(enabled=False)
    ws.gram_impl, ws.gram_mode, ws.sliced_single = 0, "sliced", False
    return ws


def test_sliced_gram_matches_float64_gram(monkeypatch):
    fake = FakeKernels()
    monkeypatch.setattr(svd, "K", fake)
    rs = np.random.RandomState(0)
    n, p = 300, 1000                                       # p not a multiple of 16 or 240: ragged last partial
    base = rs.randn(40, p)
    Y = (rs.randn(n, 40) @ base) * np.logspace(2, -3, n)[:, None] + 1e-4 * rs.randn(n, p)
    Y = Y.astype(np.float32)
    ws = _workspace()
    X = ws.matrix("Y", n, p)
    X.copy_(torch.from_numpy(Y))
    mean, G = svd.sliced_gram(ws, X, ws.timer, "merge")
    Y64 = Y.astype(np.float64)
    m = Y64.mean(axis=1)
    want = Y64 @ Y64.T - p * np.outer(m, m)
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    err = np.abs(G.numpy() - want) / np.outer(d, d)
    assert np.array_equal(mean.numpy(), m)
    # what is left: the 2^-25 (of the row bound) representation error of the slices, unbiased, averaged over K = 1000
    # (19 200 in the merge this serves), and the float32 rounding of passes B-D (2^-8 of an entry)
    assert err.max() < 5e-8 and np.median(err) < 5e-9, (err.max(), np.median(err))
    assert np.array_equal(G.numpy(), G.numpy().T)
    ks_a = math.ceil(math.ceil(p / 16) / svd.SLICE_KBLOCKS)
    assert [c[1] for c in fake.calls] == [True, False, False, False] and fake.calls[0][0] == ks_a
    # the one-GEMM Gram of the product path, same stand-in arithmetic, for scale: ~1e-7
    hi = _trunc19(Y).astype(np.float64)
    lo = _trunc19(Y - _trunc19(Y)).astype(np.float64)
    one = (lo @ hi.T + hi @ lo.T + hi @ hi.T).astype(np.float32).astype(np.float64) - p * np.outer(m, m)
    err_one = np.abs(one - want) / np.outer(d, d)
    print("sliced max/median", err.max(), np.median(err), "one GEMM max/median", err_one.max(), np.median(err_one))
    assert np.median(err_one) > 3 * np.median(err)


def test_sliced_gram_single_product_pass(monkeypatch):
    fake = FakeKernels()
    monkeypatch.setattr(svd, "K", fake)
    ws = _workspace()
    Y = np.random.RandomState(5).randn(200, 777).astype(np.float32)
    X = ws.matrix("Y", 200, 777)
    X.copy_(torch.from_numpy(Y))
    _, G0 = svd.sliced_gram(ws, X, ws.timer, "merge")
    ws.sliced_single = True
    _, G1 = svd.sliced_gram(ws, X, ws.timer, "merge")
    assert np.array_equal(G0.numpy(), G1.numpy()) and "gram.zero" in ws._bufs


def test_gram_routing_by_mode():
    ws = _workspace()
    for mode, label, n, impl, want in (("sliced", "stage2.merge", 3750, 0, True), ("sliced", "stage2.chunk", 999, 0, True),
                                       ("sliced", "stage2.roi", 999, 0, True), ("sliced", "stage2.roi_merge", 100, 0, False),
                                       ("sliced_merge", "stage2.merge", 3750, 0, True),
                                       ("sliced_merge", "stage2.roi_merge", 500, 0, True),
                                       ("sliced_merge", "stage2.chunk", 999, 0, False),
                                       ("sliced", "stage2.merge", 3750, 9, False), ("tf32x3", "stage2.merge", 3750, 0, False),
                                       ("i8", "stage2.merge", 3750, 0, False)):
        ws.gram_mode, ws.gram_impl = mode, impl
        assert svd.gram_is_sliced(ws, label, n) is want, (mode, label, n, impl)
    for mode, label, want in (("i8", "stage2.chunk", True), ("i8", "stage2.roi_merge", True), ("i8_merge", "stage2.chunk", False),
                              ("i8_merge", "stage2.merge", True), ("sliced", "stage2.merge", False), ("tf32x3", "stage2.merge", False)):
        ws.gram_mode = mode
        assert svd.gram_is_i8(ws, label) is want, (mode, label)


def test_centred_gram_dispatches_to_sliced(monkeypatch):
    fake = FakeKernels()
    monkeypatch.setattr(svd, "K", fake)
    ws = _workspace()
    X = ws.matrix("Y", 200, 480)
    X.copy_(torch.from_numpy(np.random.RandomState(3).randn(200, 480).astype(np.float32)))
    out = torch.empty((200, 200), dtype=torch.float64)
    mean, G = svd.centred_gram(ws, X, ws.timer, "stage2.chunk", out=out)
    assert G is out and [c[1] for c in fake.calls] == [True, False, False, False]
    x = X.numpy().astype(np.float64)
    want = x @ x.T - 480 * np.outer(x.mean(axis=1), x.mean(axis=1))
    assert np.abs(G.numpy() - want).max() < 1e-6 * np.abs(want).max()
