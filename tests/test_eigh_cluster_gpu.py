"""fm_eigh_topk_batched (csrc/eigh_cluster.cu: the hand-written batched eigensolver of the per-chunk Gram matrices)
against numpy.linalg.eigh and its NumPy restatement (tests/eigh_model.py): eigenvalues, residuals, orthogonality, the
output conventions fm_topk_components relies on, and the self-check on matrices the method cannot separate."""
import numpy as np
import pytest
import torch

from tests import eigh_model as M
from tests.test_eigh_model import graded_gram

pytestmark = pytest.mark.gpu


def _solve(G, k, device):
    from facemap_b200 import kernels as K

    s = K.Solver()
    Gd = torch.from_numpy(np.ascontiguousarray(G)).to(device)
    lam, W = s.eigh_topk_batched(Gd.clone(), k)
    torch.cuda.synchronize()
    flags = [st.cpu().numpy() for st in s._status]
    s._status = []
    return lam.cpu().numpy(), W.cpu().numpy(), flags[0]


@pytest.mark.parametrize("n,p,k,batch", [(2, 10, 2, 1), (5, 40, 3, 2), (37, 400, 37, 3), (128, 1500, 50, 2), (300, 4000, 120, 2),
                                         (999, 6000, 250, 3), (1024, 3000, 64, 1)])
def test_top_eigenpairs_match_lapack(cuda_device, n, p, k, batch):
    rs = np.random.RandomState(n + k)
    G = np.stack([graded_gram(rs, n, p, strong=max(1, min(24, n // 3))) for _ in range(batch)])
    lam, W, flags = _solve(G, k, cuda_device)
    assert lam.shape == (batch, k) and W.shape == (batch, k, n) and not flags.any(), flags
    for b in range(batch):
        ref = np.linalg.eigvalsh(G[b])[-k:]                      # ascending, like the output
        assert np.abs(lam[b] - ref).max() <= 2e-13 * ref[-1], (b, np.abs(lam[b] - ref).max() / ref[-1])
        R = G[b] @ W[b].T - W[b].T * lam[b][None, :]
        assert np.abs(R).max() <= 1e-12 * ref[-1], (b, np.abs(R).max() / ref[-1])
        assert np.abs(W[b] @ W[b].T - np.eye(k)).max() <= 1e-9, b


def test_matches_the_numpy_model(cuda_device):
    """same algorithm, same start vectors: eigenvalues to rounding, eigenvectors up to sign"""
    rs = np.random.RandomState(11)
    G = graded_gram(rs, 90, 1200, strong=12)
    lam, W, flags = _solve(G[None], 30, cuda_device)
    lam_m, W_m = M.eigh_topk(G, 30)                              # descending
    assert not flags.any()
    assert np.abs(lam[0][::-1] - lam_m).max() <= 1e-13 * lam_m[0]
    cos = np.abs(np.sum(W[0][::-1] * W_m, axis=1))
    assert np.abs(cos - 1).max() <= 1e-9


def test_rank_deficient_matrix(cuda_device):
    """a ROI with fewer pixels than frames: the non-zero eigenpairs are right and the self-check stays quiet (pairs in
    the numerically-zero cluster are not held to orthogonality: their U S columns vanish whatever they are)"""
    rs = np.random.RandomState(2)
    X = rs.standard_normal((200, 40))
    X -= X.mean(axis=1, keepdims=True)
    G = X @ X.T
    lam, W, flags = _solve(G[None], 60, cuda_device)
    ref = np.linalg.eigvalsh(G)[-60:]
    assert not flags.any(), flags
    assert np.abs(lam[0] - ref).max() <= 1e-12 * ref[-1]
    top = W[0][-39:]
    assert np.abs(top @ top.T - np.eye(39)).max() <= 1e-8
    assert np.abs(G @ W[0].T - W[0].T * lam[0][None, :]).max() <= 1e-11 * ref[-1]


def test_repeated_eigenvalues_raise_the_flag(cuda_device):
    """two identical diagonal blocks: every eigenvalue is double, independent inverse iterations return the same vector
    twice — the self-check must say so (the host then repeats the pass on cuSOLVER)"""
    rs = np.random.RandomState(4)
    B = graded_gram(rs, 40, 500, strong=6)
    G = np.zeros((80, 80))
    G[:40, :40] = B
    G[40:, 40:] = B
    _, _, flags = _solve(G[None], 30, cuda_device)
    assert flags[0] & 1


def test_pipeline_falls_back_to_cusolver_when_flagged(cuda_device, monkeypatch):
    """process_source repeats the pass with eig_mode='cusolver' when the self-check fires"""
    from facemap_b200 import kernels as K
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    v = SynthVideo(32, 48, 2100, seed=3).all_frames()
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device)])
    want = process.process_source(src, sbin=2, eig_mode="cusolver")
    got = process.process_source(src, sbin=2, eig_mode="cluster")
    S, Sr = got["motSv"].astype(np.float64), want["motSv"].astype(np.float64)
    assert (np.abs(S - Sr) / Sr).max() <= 1e-5                   # both solvers, same decomposition
    calls = []
    orig = K.Solver.eigh_topk_batched

    def flagged(self, G, k, scratch=None):
        out = orig(self, G, k, scratch=scratch)
        self._status[-1].fill_(1)                                # pretend the self-check fired
        calls.append(k)
        return out

    monkeypatch.setattr(K.Solver, "eigh_topk_batched", flagged)
    again = process.process_source(src, sbin=2, eig_mode="cluster")
    assert calls and np.array_equal(again["motSv"], want["motSv"]) and np.array_equal(again["motSVD"][0], want["motSVD"][0])
