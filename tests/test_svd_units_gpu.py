"""Unit parity of the SVD building blocks on the GPU against float64 NumPy:
chunk_components / merge_components (column-side Gram) and merge_components_rowside (p x p Gram for wide
matrices) must all be the centred SVD of sklearn's PCA (facemap/utils.py:773-775) with its sign rule."""
import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc

pytestmark = pytest.mark.gpu


def _spectrum_matrix(p, c, seed, decay=0.03):
    rng = np.random.RandomState(seed)
    k = min(p, c)
    Q1, _ = np.linalg.qr(rng.randn(p, k))
    Q2, _ = np.linalg.qr(rng.randn(c, k))
    s = 100.0 * np.exp(-decay * np.arange(k)) + 0.5
    # column means small against the signal, as in the path (the chunk matrices have the running averages
    # removed, the stacked components are centred): the Gram form removes the mean by a rank-1 correction
    # and would lose digits to cancellation under a mean much larger than the deviations
    return ((Q1 * s) @ Q2.T + rng.randn(p, c) * 0.01 + rng.randn(1, c) * 0.05).astype(np.float32)


def _check(Ut, U, sval, Y, k, s_floor=2e-2):
    u_ref, s_ref, _ = orc.svd_centered(Y, k, "exact")
    S, Sr = sval.cpu().numpy().astype(np.float64), s_ref.astype(np.float64)
    big = Sr >= s_floor * Sr[0]
    assert (np.abs(S - Sr)[big] / Sr[big]).max() <= 1e-4
    Ug = U.cpu().numpy().astype(np.float64)
    assert torch.equal(Ut.T.contiguous(), U)
    lead = min(8, k)
    dots = (Ug[:, :lead] * u_ref[:, :lead].astype(np.float64)).sum(axis=0)
    assert np.abs(dots - 1).max() < 1e-5, dots        # same vectors AND same sign (svd_flip rule)


@pytest.mark.parametrize("p,c,k", [(900, 300, 120), (2000, 750, 500)])
def test_merge_components_column_side(cuda_device, p, c, k):
    from facemap_b200 import kernels as K, svd

    Y = _spectrum_matrix(p, c, seed=p)
    Yt = K.alloc_matrix(c, p, cuda_device)
    Yt.copy_(torch.from_numpy(np.ascontiguousarray(Y.T)))
    ws = svd.SvdWorkspace(cuda_device)
    Ut, U, sval = svd.merge_components(ws, Yt, k, svd.StageTimer(False))
    torch.cuda.synchronize()
    _check(Ut, U, sval, Y, k)


@pytest.mark.parametrize("p,c,k", [(285, 500, 284), (240, 480, 239), (700, 1500, 500)])
def test_merge_components_row_side_equals_column_side(cuda_device, p, c, k):
    """wide matrices (small ROI): the p x p formulation and the c x c formulation are the same SVD."""
    from facemap_b200 import kernels as K, svd

    Y = _spectrum_matrix(p, c, seed=c)
    Yt = K.alloc_matrix(c, p, cuda_device)
    Yt.copy_(torch.from_numpy(np.ascontiguousarray(Y.T)))
    ws = svd.SvdWorkspace(cuda_device)
    tm = svd.StageTimer(False)
    Ut_r, U_r, s_r = svd.merge_components_rowside(ws, Yt, k, tm)
    torch.cuda.synchronize()
    _check(Ut_r, U_r, s_r, Y, k)
    # column-side on the same data (bypass the dispatch)
    Wt, sval, rbias, rscale = svd.gram_eig(ws, Yt, k, tm, True, "t")
    Ymat = K.alloc_matrix(p, c, cuda_device)
    K.transpose(Yt, Ymat)
    Ut_c = K.alloc_matrix(k, p, cuda_device)
    ws.gemm(Wt, Ymat, Ut_c, rscale=rscale, rbias=rbias, chain=svd.KCHAIN_LEFT)
    torch.cuda.synchronize()
    lead = min(20, k)
    a, b = Ut_r[:lead].double(), Ut_c[:lead].double()
    assert float((a - b).abs().max()) < 1e-4
    big = s_r >= 2e-2 * s_r[0]
    assert float(((s_r - sval).abs() / s_r)[big].max()) <= 1e-4


def test_chunk_components_match_pca(cuda_device):
    """(U S)^T of one chunk: svdecon(X^T, k) with rows = frames (process.py:246-251)."""
    from facemap_b200 import kernels as K, svd

    n, p, k = 999, 1536, 250
    Xh = _spectrum_matrix(n, p, seed=3)
    X = K.alloc_matrix(n, p, cuda_device)
    X.copy_(torch.from_numpy(Xh))
    out = K.alloc_matrix(k, p, cuda_device)
    svd.chunk_components(svd.SvdWorkspace(cuda_device), X, k, out, svd.StageTimer(False))
    torch.cuda.synchronize()
    u, s, _ = orc.svd_centered(np.ascontiguousarray(Xh.T), k, "exact")
    ref = (u.astype(np.float64) * s.astype(np.float64)).T           # [k, p]
    got = out.cpu().numpy().astype(np.float64)
    lead = 40
    assert np.abs(got[:lead] - ref[:lead]).max() <= 2e-5 * np.abs(ref[:lead]).max()
    nr, ng = np.linalg.norm(ref, axis=1), np.linalg.norm(got, axis=1)
    big = nr >= 2e-2 * nr[0]
    assert (np.abs(ng - nr)[big] / nr[big]).max() <= 1e-4
