"""fm_gemm_i8_tn (tcgen05.mma kind::i8, int32 accumulation) against exact integer arithmetic.

First run on a B200 at the start of round 2 (19 passed, profiles/r02a_first_gpu_run.md); part of the regular `-m gpu` run.
Exact equality everywhere — integer work; ragged shapes, both signednesses, split-K partials, upper-triangle skipping,
a K range long enough to approach the int32 limit."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _operand(rs, rows, K, signed, device, extreme=False):
    lo, hi = (-128, 128) if signed else (0, 256)
    a = rs.randint(lo, hi, size=(rows, K))
    if extreme:
        a[:] = hi - 1 if not signed else lo
    ld = -(-K // 16) * 16
    buf = torch.zeros((rows, ld), dtype=torch.int8 if signed else torch.uint8, device=device)
    buf[:, :K] = torch.from_numpy(a.astype(np.int8 if signed else np.uint8)).to(device)
    return a.astype(np.int64), buf[:, :K]


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 1000), (999, 999, 19200), (130, 257, 77), (1, 129, 33)])
@pytest.mark.parametrize("a_signed,b_signed", [(False, False), (False, True), (True, True)])
def test_gemm_i8_exact(cuda_device, M, N, K, a_signed, b_signed):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(M + 7 * N + K)
    a, A = _operand(rs, M, K, a_signed, cuda_device)
    b, B = _operand(rs, N, K, b_signed, cuda_device)
    want = (a.astype(np.float64) @ b.astype(np.float64).T).astype(np.int64)      # exact: |sums| < 2^53, BLAS speed
    ldc = -(-N // 4) * 4
    C = torch.full((M, ldc), -77, dtype=torch.int32, device=cuda_device)
    K_.gemm_i8_tn(A, B, C)
    assert np.array_equal(C[:, :N].cpu().numpy().astype(np.int64), want)
    ks = 3
    if K >= 64 * ks:
        ws = torch.full((ks, M, ldc), -77, dtype=torch.int32, device=cuda_device)
        K_.gemm_i8_tn(A, B, ksplit=ks, workspace=ws)
        assert np.array_equal(ws[:, :, :N].cpu().numpy().astype(np.int64).sum(axis=0), want)


def test_gemm_i8_upper_only_and_int32_range(cuda_device):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(0)
    n, K = 700, 32768                                       # 255^2 * 32768 = 2 130 739 200 < 2^31
    a, A = _operand(rs, n, K, False, cuda_device)
    a[5], a[6] = 255, 255
    A[5], A[6] = 255, 255
    C = torch.full((n, 704), -1, dtype=torch.int32, device=cuda_device)
    K_.gemm_i8_tn(A, A, C, upper_only=True)
    got = C[:, :n].cpu().numpy().astype(np.int64)
    want = (a.astype(np.float64) @ a.astype(np.float64).T).astype(np.int64)
    iu = np.triu_indices(n)
    assert np.array_equal(got[iu], want[iu]) and got[5, 6] == 255 * 255 * K
    assert (got[256:, :256] == -1).all()                    # the tile below the diagonal was skipped
    with pytest.raises(Exception):
        K_.gemm_i8_tn(torch.zeros((128, 40000), dtype=torch.uint8, device=cuda_device),
                      torch.zeros((256, 40000), dtype=torch.uint8, device=cuda_device), C)   # would overflow int32


def _graded(rs, n, p):
    base = rs.randn(40, p)
    Y = (rs.randn(n, 40) @ base) * np.logspace(2, -3, n)[:, None] + 1e-4 * rs.randn(n, p) + 0.3
    return Y.astype(np.float32)


def _model_digits(Y):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from model_accuracy import slice_rows_i8

    return slice_rows_i8(Y)


def test_slice_rows_i8_matches_model(cuda_device):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(0)
    for n, p in ((7, 33), (300, 1000), (64, 19200)):
        Y = _graded(rs, n, p)
        X = K_.alloc_matrix(n, p, cuda_device)
        X.copy_(torch.from_numpy(Y))
        Q0, Q1, Q2, exps = K_.slice_rows_i8(X)
        (m0, m1, m2), scale = _model_digits(Y)
        assert np.array_equal(Q0.cpu().numpy(), m0.astype(np.int8))
        assert np.array_equal(Q1.cpu().numpy(), m1.astype(np.uint8)) and np.array_equal(Q2.cpu().numpy(), m2.astype(np.uint8))
        assert np.array_equal(exps.cpu().numpy(), (np.log2(scale) + 7).astype(np.int32))


def test_i8_gram_against_float64(cuda_device):
    from facemap_b200 import kernels as K_, svd

    rs = np.random.RandomState(2)
    n, p = 700, 5000
    Y = _graded(rs, n, p)
    X = K_.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(Y))
    ws = svd.SvdWorkspace(cuda_device, gram_mode="i8")
    mean, G = svd.centred_gram(ws, X, ws.timer, "stage2.merge")
    Y64 = Y.astype(np.float64)
    m = Y64.mean(axis=1)
    want = Y64 @ Y64.T - p * np.outer(m, m)
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    err = np.abs(G.cpu().numpy() - want) / np.outer(d, d)
    print("i8 gram max/median rel err", err.max(), np.median(err))
    assert err.max() < 3e-8 and np.median(err) < 3e-9
    # against the digits' own exact Gram: nothing but float64 rounding
    (q0, q1, q2), scale = _model_digits(Y)
    T = scale[:, None] * (q0 + q1 / 256.0 + q2 / 65536.0)
    exact = T @ T.T - p * np.outer(m, m)
    assert np.abs(G.cpu().numpy() - exact).max() <= 1e-12 * d.max() ** 2


def test_i8_gram_in_column_blocks_and_blocked_left_vectors(cuda_device, monkeypatch):
    """views too wide to slice at once (sbin 1 on a large view) go through the Gram in column blocks with float64
    accumulation (fm_gram_accumulate_i8) and through the left-vector GEMMs in pixel blocks: forced here on a small
    operand by shrinking the block sizes, against the one-block results and float64"""
    from facemap_b200 import kernels as K_, svd

    rs = np.random.RandomState(7)
    n, p, k = 300, 5000, 40
    Y = _graded(rs, n, p)
    X = K_.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(Y))
    ws = svd.SvdWorkspace(cuda_device, gram_mode="i8")
    _, G1 = svd.centred_gram(ws, X, ws.timer, "stage2.chunk")
    G1 = G1.clone()
    monkeypatch.setattr(svd, "I8_BLOCK_COLS", 1536)              # 4 blocks, the last one ragged (5000 = 3 * 1536 + 392)
    _, G4 = svd.centred_gram(ws, X, ws.timer, "stage2.chunk")
    Y64 = Y.astype(np.float64)
    m = Y64.mean(axis=1)
    want = Y64 @ Y64.T - p * np.outer(m, m)
    d = np.sqrt(np.diag(Y64 @ Y64.T))
    for G in (G1, G4):
        assert (np.abs(G.cpu().numpy() - want) / np.outer(d, d)).max() < 3e-8
    assert np.array_equal(G4.cpu().numpy(), G4.cpu().numpy().T)
    # per-block digits differ from whole-row digits (every block has its own row exponents): equal to rounding only
    assert (np.abs(G4.cpu().numpy() - G1.cpu().numpy()) / np.outer(d, d)).max() < 3e-8
    # left vectors (U S)^T = W^T X - (W^T m) 1^T in pixel blocks == in one piece, bit for bit (same K chains per pixel)
    lam = torch.linalg.eigvalsh(G1)
    evec = torch.linalg.eigh(G1)[1].t().contiguous()
    mean = K_.row_means(X)
    out1 = K_.alloc_matrix(k, p, cuda_device)
    svd.left_components(ws, X, evec, lam, mean, k, out1, ws.timer, "stage2.chunk")
    monkeypatch.setattr(svd, "LEFT_BLOCK_PIXELS", 2048)
    out3 = K_.alloc_matrix(k, p, cuda_device)
    svd.left_components(ws, X, evec, lam, mean, k, out3, ws.timer, "stage2.chunk")
    assert torch.equal(out1, out3)
    W = evec.flip(0)[:k].cpu().numpy()
    want_us = W @ (Y64 - m[:, None])
    sg = np.sign(np.sum(want_us * out1.cpu().numpy(), axis=1))
    assert np.abs(out1.cpu().numpy() * sg[:, None] - want_us).max() <= 1e-5 * np.abs(want_us).max()    # 3xTF32 GEMM


def test_pipeline_with_i8_grams_resolves_all_singular_values(cuda_device):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo
    from oracle import facemap_oracle as orc

    H, W, N, sbin = 96, 128, 2600, 2
    v = SynthVideo(H, W, N, seed=5).all_frames()
    o = orc.run_oracle([v], sbin=sbin, motSVD=True, movSVD=True, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=sbin, movSVD=True,
                               gram_mode="i8")
    for sk in ("motSv", "movSv"):
        S_ref = o[sk].astype(np.float64)
        rel = np.abs(p[sk].astype(np.float64) - S_ref) / S_ref
        print(sk, "i8", [float(rel[:k].max()) for k in (10, 100, 250, 500)])
        assert rel.max() <= 1e-4, sk
