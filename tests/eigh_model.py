"""NumPy restatement of the hand-written symmetric eigensolver (facemap_b200/csrc/eigh_cluster.cu): the top-k eigenpairs
of a dense symmetric matrix by (1) Householder tridiagonalisation, lower variant (LAPACK dsytd2 conventions), (2) Sturm
bisection for the k largest eigenvalues of the tridiagonal matrix, (3) inverse iteration with partial pivoting for
their eigenvectors (dlagtf / dlagts), (4) back-transformation with the reflectors.  Test infrastructure: the CUDA kernels
are held to this model and to numpy.linalg.eigh (tests/test_eigh_model.py, tests/test_eigh_cluster_gpu.py)."""
import numpy as np


def tridiagonalize(A):
    """A symmetric [n, n] -> (d [n], e [n-1], tau [n-1], V [n, n]) with Q^T A Q = T, Q = H_0 H_1 ... H_{n-2},
    H_i = I - tau_i v_i v_i^T, v_i[:i+1] = 0, v_i[i+1] = 1, v_i[i+2:] = V[i+2:, i]."""
    A = np.array(A, np.float64)
    n = A.shape[0]
    d, e, tau = np.zeros(n), np.zeros(max(n - 1, 0)), np.zeros(max(n - 1, 0))
    V = np.zeros((n, n))
    for i in range(n - 1):
        x = A[i + 1:, i].copy()
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0.0:
            t, beta, v = 0.0, alpha, np.zeros_like(x)
            v[0] = 1.0
        else:
            beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
            t = (beta - alpha) / beta
            v = x / (alpha - beta)
            v[0] = 1.0
        e[i], tau[i] = beta, t
        V[i + 1:, i] = v
        d[i] = A[i, i]
        if t != 0.0:
            S = A[i + 1:, i + 1:]
            p = t * (S @ v)
            w = p - (0.5 * t * (p @ v)) * v
            S -= np.outer(v, w) + np.outer(w, v)
    d[n - 1] = A[n - 1, n - 1]
    return d, e, tau, V


def sturm_count(d, e2, x, pivmin):
    """number of eigenvalues of the tridiagonal (d, e) strictly below each x[j] (vectorised over the shifts)"""
    x = np.atleast_1d(np.asarray(x, np.float64))
    q = d[0] - x
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    c = (q < 0).astype(np.int64)
    for i in range(1, len(d)):
        q = d[i] - x - e2[i - 1] / q
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        c += q < 0
    return c


def gershgorin(d, e):
    n = len(d)
    ea = np.abs(np.r_[0.0, e]) + np.abs(np.r_[e, 0.0])
    lo, hi = float((d - ea).min()), float((d + ea).max())
    tnorm = max(abs(lo), abs(hi))
    tiny = np.finfo(np.float64).tiny
    pivmin = max(tiny * max(1.0, float((e * e).max()) if e.size else 0.0), tiny)
    pad = 2.0 * tnorm * np.finfo(np.float64).eps * n + 2.0 * pivmin
    return lo - pad, hi + pad, pivmin


def top_eigenvalues(d, e, k, iters=64):
    """k largest eigenvalues (descending) by bisection from the Gershgorin interval: `iters` halvings for every
    eigenvalue at once (the interval shrinks by 2^-64: below one ulp of the matrix norm)"""
    n = len(d)
    e2 = e * e
    lo0, hi0, pivmin = gershgorin(d, e)
    lo, hi = np.full(k, lo0), np.full(k, hi0)
    want = n - 1 - np.arange(k)             # j-th largest = ascending index n-1-j:  count(x) <= n-1-j  <=>  x <= lam_j
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        below = sturm_count(d, e2, mid, pivmin) <= want
        lo = np.where(below, mid, lo)
        hi = np.where(below, hi, mid)
    return 0.5 * (lo + hi)


def start_vector(n, j):
    """deterministic start vector of eigenvector j: entries in (-1, 1) from an integer hash (the kernel's)"""
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B1) + np.uint64(j + 1) * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    return (h.astype(np.float64) + 0.5) / 2147483648.0 - 1.0


def inverse_iteration(d, e, lam, iters=3):
    """eigenvectors [k, n] of the tridiagonal (d, e) for the (accurate) eigenvalues lam [k]: Gaussian elimination with
    partial pivoting of T - lam I (LAPACK dlagtf) and `iters` solves from a fixed start vector (dlagts), vectorised
    over the eigenvalues"""
    lam = np.atleast_1d(np.asarray(lam, np.float64))
    n, k = len(d), len(lam)
    eps = np.finfo(np.float64).eps
    a = d[None, :] - lam[:, None]                     # diagonal                       [k, n]
    b = np.tile(e, (k, 1))                            # super-diagonal                 [k, n-1]
    c = np.tile(e, (k, 1))                            # sub-diagonal -> multipliers
    d2 = np.zeros((k, max(n - 2, 0)))                 # second super-diagonal created by row interchanges
    swap = np.zeros((k, max(n - 1, 0)), bool)
    tnorm = max(np.abs(d).max(), np.abs(e).max() if n > 1 else 0.0, 1e-300)
    tol = eps * tnorm
    for i in range(n - 1):
        keep = np.abs(a[:, i]) >= np.abs(c[:, i])
        ai = np.where(keep & (a[:, i] == 0.0), tol, a[:, i])
        # no interchange
        m0 = c[:, i] / np.where(keep, ai, 1.0)
        # interchange rows i and i+1
        ci = np.where(keep, 1.0, c[:, i])
        m1 = ai / ci
        nxt = a[:, i + 1]
        a_next = np.where(keep, nxt - m0 * b[:, i], b[:, i] - m1 * nxt)
        if i < n - 2:
            d2[:, i] = np.where(keep, 0.0, b[:, i + 1])
            b[:, i + 1] = np.where(keep, b[:, i + 1], -m1 * b[:, i + 1])
        b[:, i] = np.where(keep, b[:, i], nxt)
        a[:, i] = np.where(keep, ai, c[:, i])
        a[:, i + 1] = a_next
        c[:, i] = np.where(keep, m0, m1)
        swap[:, i] = ~keep
    x = np.stack([start_vector(n, j) for j in range(k)])
    for _ in range(iters):
        for i in range(n - 1):                        # forward: the row operations applied to the right-hand side
            xi, xn = x[:, i].copy(), x[:, i + 1].copy()
            x[:, i] = np.where(swap[:, i], xn, xi)
            x[:, i + 1] = np.where(swap[:, i], xi - c[:, i] * xn, xn - c[:, i] * xi)
        for i in range(n - 1, -1, -1):                # backward substitution with U = (a, b, d2), tiny pivots replaced
            t = x[:, i].copy()
            if i + 1 < n:
                t -= b[:, i] * x[:, i + 1]
            if i + 2 < n:
                t -= d2[:, i] * x[:, i + 2]
            piv = a[:, i]
            piv = np.where(np.abs(piv) < tol, np.where(piv < 0, -tol, tol), piv)
            x[:, i] = t / piv
        x /= np.abs(x).max(axis=1, keepdims=True)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def back_transform(Z, tau, V):
    """rows of Z [k, n] are eigenvectors of T; returns Q Z^T as rows"""
    Z = np.array(Z, np.float64)
    n = Z.shape[1]
    for i in range(n - 2, -1, -1):
        if tau[i] == 0.0:
            continue
        v = V[i + 1:, i]
        s = Z[:, i + 1:] @ v
        Z[:, i + 1:] -= tau[i] * np.outer(s, v)
    return Z


def eigh_topk(A, k):
    """(lam [k] descending, W [k, n] rows = eigenvectors)"""
    d, e, tau, V = tridiagonalize(A)
    lam = top_eigenvalues(d, e, k)
    return lam, back_transform(inverse_iteration(d, e, lam), tau, V)
