"""The NumPy restatement of the hand-written eigensolver (tests/eigh_model.py: Householder tridiagonalisation, Sturm
bisection, inverse iteration with partial pivoting, back-transformation) against numpy.linalg.eigh — the algorithm
csrc/eigh_cluster.cu runs, checked where no GPU is needed."""
import numpy as np
import pytest

from tests import eigh_model as M


def graded_gram(rs, n, p, strong=24):
    """a chunk-like Gram matrix: a few strong components over a flat noise floor, rows centred"""
    X = rs.standard_normal((n, p))
    w = np.ones(p)
    w[:strong] = np.logspace(2.3, 0.3, strong)
    X = X * w
    X -= X.mean(axis=1, keepdims=True)
    return X @ X.T


@pytest.mark.parametrize("n,p,k", [(5, 40, 5), (64, 900, 20), (150, 2500, 60)])
def test_model_matches_lapack(n, p, k):
    rs = np.random.RandomState(n)
    G = graded_gram(rs, n, p, strong=min(24, n // 2))
    lam, W = M.eigh_topk(G, k)
    ref = np.linalg.eigvalsh(G)[::-1][:k]
    assert np.abs(lam - ref).max() <= 1e-13 * ref[0]
    assert np.abs(G @ W.T - W.T * lam[None, :]).max() <= 1e-13 * ref[0]
    assert np.abs(W @ W.T - np.eye(k)).max() <= 1e-11


def test_tridiagonal_form_is_similar():
    rs = np.random.RandomState(3)
    G = graded_gram(rs, 40, 300, strong=8)
    d, e, tau, V = M.tridiagonalize(G)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(G)).max() <= 1e-13 * np.abs(G).max()
    # Q from the reflectors: Q^T G Q = T
    Q = M.back_transform(np.eye(40), tau, V).T
    assert np.abs(Q.T @ G @ Q - T).max() <= 1e-12 * np.abs(G).max()


def test_rank_deficient_matrix_keeps_its_range():
    """fewer columns than rows: the non-zero eigenpairs are right, the null-space vectors are whatever they are"""
    rs = np.random.RandomState(5)
    X = rs.standard_normal((60, 12))
    G = X @ X.T
    lam, W = M.eigh_topk(G, 20)
    ref = np.linalg.eigvalsh(G)[::-1][:20]
    assert np.abs(lam - ref).max() <= 1e-13 * ref[0]
    assert np.abs(W[:12] @ W[:12].T - np.eye(12)).max() <= 1e-10
    assert np.abs(G @ W.T - W.T * lam[None, :]).max() <= 1e-12 * ref[0]
