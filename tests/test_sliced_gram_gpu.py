"""The sliced merge Gram (svd.sliced_gram, fm_slice_rows) on the device.

WRITTEN WITHOUT A GPU (round 1 had spent its GPU budget): these tests have not run on a B200 yet, so they are opt-in —
`FM_EXPERIMENTAL=1 python -m pytest tests/test_sliced_gram_gpu.py -m gpu`.  The first one that has to hold is the
hardware property the scheme rests on: a tcgen05 kind::tf32 MMA chain whose products and partial sums are integers
(times a power of two) below 2^24 accumulates without rounding.  CPU-side evidence: tests/test_slice_math.py,
tests/test_sliced_gram_host.py, scripts/model_accuracy.py (`passesm`)."""
import os

import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc

pytestmark = pytest.mark.gpu


def _graded(rs, n, p):
    base = rs.randn(40, p)
    Y = (rs.randn(n, 40) @ base) * np.logspace(2, -3, n)[:, None] + 1e-4 * rs.randn(n, p)
    return Y.astype(np.float32)


def _numpy_slices(A):
    amax = np.abs(A).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax, 1e-300))) + 1, -100.0)
    R, out = A.astype(np.float32), []
    for d in range(3):
        q = np.exp2(e - 8 * (d + 1)).astype(np.float32)[:, None]
        S = (np.rint(R / q) * q).astype(np.float32)
        R = (R - S).astype(np.float32)
        out.append(S)
    return out


def test_slice_rows_matches_numpy_bitwise(cuda_device):
    from facemap_b200 import kernels as K

    rs = np.random.RandomState(0)
    for n, p in ((7, 33), (300, 1000), (64, 19200)):
        Y = _graded(rs, n, p)
        Y[n // 2] = 0
        X = K.alloc_matrix(n, p, cuda_device)
        X.copy_(torch.from_numpy(Y))
        planes = K.slice_rows(X)
        for got, want in zip(planes, _numpy_slices(Y)):
            assert np.array_equal(got.cpu().numpy(), want)


def test_top_slice_product_is_exact_on_the_tensor_cores(cuda_device):
    """pass A: partials of <= 240 K of S0 S0^T leave the chained tcgen05 kernel as exact integers"""
    from facemap_b200 import kernels as K

    rs = np.random.RandomState(1)
    n, p = 300, 16 * 15 * 7 + 40
    Y = _graded(rs, n, p)
    Y[3] = 1000.0                                          # constant rows: the largest sums the digits allow
    Y[4] = -1000.0
    X = K.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(Y))
    S0 = K.slice_rows(X)[0]
    Z = K.alloc_matrix(n, p, cuda_device, zero=True)
    KB = -(-p // 16)
    ks = -(-KB // 15)
    part = torch.full((ks, n, 320), float("nan"), dtype=torch.float32, device=cuda_device)
    K.gemm_tn(S0, S0, ksplit=ks, workspace=part, upper_only=True, impl=0, B_lo=Z, A_lo=Z)
    s0 = S0.cpu().numpy().astype(np.float64)
    got = part.cpu().numpy().astype(np.float64)
    iu = np.triu_indices(n)
    for z in range(ks):
        k0, k1 = 16 * ((z * KB) // ks), min(p, 16 * (((z + 1) * KB) // ks))
        want = s0[:, k0:k1] @ s0[:, k0:k1].T
        assert np.array_equal(got[z][:, :n][iu], want[iu]), z


def test_single_product_kernel_equals_zero_lo_plane(cuda_device):
    """fm_gemm_tn impl 10 (one MMA per K step) == impl 0 with all-zero lo planes, bit for bit, on slice operands"""
    from facemap_b200 import kernels as K

    rs = np.random.RandomState(4)
    n, p = 520, 3000
    X = K.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(_graded(rs, n, p)))
    S0, S1, _ = K.slice_rows(X)
    Z = K.alloc_matrix(n, p, cuda_device, zero=True)
    for A, B, upper in ((S0, S0, True), (S0, S1, False)):
        Zb = Z
        a = torch.full((7, n, 544), float("nan"), dtype=torch.float32, device=cuda_device)
        b = torch.full((7, n, 544), float("nan"), dtype=torch.float32, device=cuda_device)
        K.gemm_tn(A, B, ksplit=7, workspace=a, upper_only=upper, impl=0, B_lo=Zb, A_lo=Z)
        K.gemm_tn(A, B, ksplit=7, workspace=b, upper_only=upper, impl=10)
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
        c = torch.empty((n, 544), dtype=torch.float32, device=cuda_device)
        K.gemm_tn(A, B, c, impl=10)
        want = A.cpu().numpy().astype(np.float64) @ B.cpu().numpy().astype(np.float64).T
        got = c[:, :n].cpu().numpy().astype(np.float64)
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


def test_sliced_gram_against_float64(cuda_device):
    from facemap_b200 import kernels as K, svd

    rs = np.random.RandomState(2)
    n, p = 700, 5000
    Y = _graded(rs, n, p)
    X = K.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(Y))
    ws = svd.SvdWorkspace(cuda_device, gram_mode="sliced")
    mean, G = svd.sliced_gram(ws, X, ws.timer, "merge")
    ws.gram_mode = "tf32x3"
    mean1, G1 = svd.centred_gram(ws, X, ws.timer, "merge")
    Y64 = Y.astype(np.float64)
    m = Y64.mean(axis=1)
    want = Y64 @ Y64.T - p * np.outer(m, m)
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    err = np.abs(G.cpu().numpy() - want) / np.outer(d, d)
    err1 = np.abs(G1.cpu().numpy() - want) / np.outer(d, d)
    print("sliced max/median", err.max(), np.median(err), "one GEMM max/median", err1.max(), np.median(err1))
    assert err.max() < 3e-8 and np.median(err) < 3e-9
    assert np.median(err1) > 3 * np.median(err)
    ws.gram_mode, ws.sliced_single = "sliced", True
    _, G2 = svd.sliced_gram(ws, X, ws.timer, "merge")
    assert torch.equal(G, G2)


def test_pipeline_with_sliced_merge_gram_resolves_all_singular_values(cuda_device):
    """north-star tolerance (1e-4 relative) on ALL 500 singular values against the exact oracle, not only on the
    components above the float32-Gram floor"""
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    H, W, N, sbin = 96, 128, 2600, 2
    v = SynthVideo(H, W, N, seed=5).all_frames()
    o = orc.run_oracle([v], sbin=sbin, motSVD=True, movSVD=True, svd="exact")
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device)])
    p = process.process_source(src, sbin=sbin, movSVD=True, gram_mode="sliced")
    p1 = process.process_source(src, sbin=sbin, movSVD=True, gram_mode="tf32x3")
    for sk in ("motSv", "movSv"):
        S_ref = o[sk].astype(np.float64)
        rel = np.abs(p[sk].astype(np.float64) - S_ref) / S_ref
        rel1 = np.abs(p1[sk].astype(np.float64) - S_ref) / S_ref
        print(sk, "sliced", [float(rel[:k].max()) for k in (10, 100, 250, 500)],
              "one GEMM", [float(rel1[:k].max()) for k in (10, 100, 250, 500)])
        assert rel.max() <= 1e-4, sk
    assert np.array_equal(p["motion"][0], o["motion"][0]) and np.array_equal(p["avgmotion"][0], o["avgmotion"][0])
