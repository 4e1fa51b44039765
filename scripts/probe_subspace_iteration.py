"""CPU study (NumPy float64, no GPU): could an iterative, GEMM-shaped solver replace cusolverDnXsyevdx for the merge
problem "top 500 of the 3 750 x 3 750 Gram of the stacked chunk components" (facemap/process.py:264-295)?

Builds the real merge problem of the path — 15 chunks x 250 components of a synthetic video with configs[1]'s binned
geometry (19 200 pixels), chunk components in float32 like the reference stores them — solves it exactly
(numpy.linalg.eigh), then runs block subspace iteration with Rayleigh-Ritz on the same Gram matrix and records after how
many multiplications by G the path's own gates hold:

  (a) all 500 singular values within 1e-4 relative (the north star),
  (b) every gap-resolved mask (tests/helpers.gap_resolved) within sin(angle) 1e-3.

Two variants: plain powers of G, and a Chebyshev filter that damps [lambda_min, lambda_cut] (cut = the Ritz value just
below the block), each with an orthonormalisation after every `per` multiplications.  The cost model at the bottom
turns the counts into B200 time with the measured rates of this repo (profiles/r02_eig.md).

usage: python scripts/probe_subspace_iteration.py [motion|movie] [H W sbin]      (~10 min on 8 cores)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facemap_b200.synth import SynthVideo           # noqa: E402
from oracle import facemap_oracle as orc            # noqa: E402
from tests.helpers import gap_resolved              # noqa: E402

K_KEEP = 500


def merge_problem(kind, H, W, sbin, N=16000):
    """Y [p, 3750] float32: the stacked U S blocks of the 15 chunks (oracle, Gram back-end), and the centred Gram."""
    video = SynthVideo(H, W, N, seed=0)              # the generator of bench.py's workload (64 modes + noise)
    nt0, nsegs, nc, tf = orc.stage2_plan(N)
    Lyb, Lxb, cols = orc.binned_geometry([H], [W], sbin)
    frames = video.all_frames()
    src = orc.ArrayVideo([frames])
    avgframe, avgmotion = orc.stage1_means(src, sbin)
    blocks = []
    for t in tf:
        ims = src.get(int(t), nt0)
        X_mot, X_mov, _, _ = orc.chunk_matrices(ims, sbin, Lyb, Lxb, cols, avgframe, avgmotion, kind == "motion",
                                                kind == "movie")
        X = X_mot if kind == "motion" else X_mov
        u, s, _ = orc.svd_centered(X.T, min(nc, X.shape[-1]), "gram")
        blocks.append((u * s).astype(np.float32))
    Y = np.ascontiguousarray(np.concatenate(blocks, axis=1)).astype(np.float64)
    Yc = Y - Y.mean(axis=0)
    return Yc, Yc.T @ Yc


def gates(G, Yc, lam_ref, W_ref, Q):
    """Rayleigh-Ritz on span(Q); returns (max rel error of the 500 singular values, worst resolved mask angle)."""
    H = Q.T @ (G @ Q)
    th, Z = np.linalg.eigh((H + H.T) / 2)
    th, Z = th[::-1][:K_KEEP], Z[:, ::-1][:, :K_KEEP]
    Wr = Q @ Z
    S, S_ref = np.sqrt(np.maximum(th, 0)), np.sqrt(lam_ref[:K_KEEP])
    s_err = float(np.max(np.abs(S - S_ref) / S_ref))
    res = np.flatnonzero(gap_resolved(S_ref, 1e-3)[0])
    # masks U = Yc w / s: the angle between masks equals the angle between the eigenvectors w
    cosv = np.abs(np.sum(Wr[:, res] * W_ref[:, res], axis=0))
    ang = float(np.max(np.sqrt(np.maximum(0.0, 1.0 - cosv ** 2)))) if res.size else 0.0
    return s_err, ang, th


def iterate(G, Yc, lam_ref, W_ref, block, scheme, per, max_mults=120, seed=0):
    n = G.shape[0]
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, block)))
    mults, log = 0, []
    lam_min = float(lam_ref[-1])
    while mults < max_mults:
        if scheme == "power":
            for _ in range(per):
                Q = G @ Q
                mults += 1
        else:                                   # Chebyshev of degree `per` damping [lam_min, cut]
            H = Q.T @ (G @ Q)
            cut = float(np.sort(np.linalg.eigvalsh((H + H.T) / 2))[0]) if mults else float(np.trace(G) / n)
            c, e = (cut + lam_min) / 2, (cut - lam_min) / 2
            Y0, Y1 = Q, (G @ Q - c * Q) / e
            mults += 1
            for _ in range(per - 1):
                Y0, Y1 = Y1, 2 * (G @ Y1 - c * Y1) / e - Y0
                mults += 1
                sc = np.linalg.norm(Y1, axis=0).max()                          # keep the numbers finite
                Y0, Y1 = Y0 / sc, Y1 / sc
            Q = Y1
        Q, _ = np.linalg.qr(Q)
        s_err, ang, _ = gates(G, Yc, lam_ref, W_ref, Q)
        log.append((mults, s_err, ang))
        if s_err <= 1e-4 and ang <= 1e-3:
            break
    return log


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "motion"
    H, W, sbin = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (240, 320, 2)
    t0 = time.time()
    cache = f"/tmp/probe_si_{kind}_{H}x{W}_s{sbin}.npy"
    if os.path.exists(cache):
        Yc = np.load(cache)
        G = Yc.T @ Yc
    else:
        Yc, G = merge_problem(kind, H, W, sbin)
        np.save(cache, Yc)
    lam, Wv = np.linalg.eigh(G)
    lam, Wv = lam[::-1], Wv[:, ::-1]
    S = np.sqrt(np.maximum(lam, 0))
    out = {"kind": kind, "pixels": int(Yc.shape[0]), "n": int(G.shape[0]), "build_s": round(time.time() - t0, 1),
           "S_k_over_S_0": {str(k): float(S[k] / S[0]) for k in (10, 50, 100, 250, 499, 500, 600, 750, 1000, 2000, 3749)},
           "gap_resolved_components": int(np.count_nonzero(gap_resolved(S[:K_KEEP], 1e-3)[0])), "runs": []}
    print(json.dumps({k: v for k, v in out.items() if k != "runs"}), flush=True)
    for block, scheme, per in ((600, "power", 1), (750, "power", 1), (1000, "power", 1), (1500, "power", 1),
                               (750, "power", 2), (750, "cheb", 4), (750, "cheb", 8), (1000, "cheb", 8)):
        t1 = time.time()
        log = iterate(G, Yc, lam, Wv, block, scheme, per)
        mults, s_err, ang = log[-1]
        run = {"block": block, "scheme": scheme, "mults_between_orthonormalisations": per, "multiplications_by_G": mults,
               "orthonormalisations": len(log), "converged": bool(s_err <= 1e-4 and ang <= 1e-3),
               "sv_rel_err": s_err, "worst_resolved_mask_angle": ang,
               "trace": [(m, float(f"{a:.2e}"), float(f"{b:.2e}")) for m, a, b in log[:: max(1, len(log) // 8)]],
               "seconds": round(time.time() - t1, 1)}
        # B200 cost model: float64 GEMM 3750 x 3750 x block at 30 TFLOP/s; orthonormalisation (Cholesky QR: Gram of Q,
        # potrf on `block`, trsm) ~ 2 n b^2 / 30e12 + 19 us x block for the factorisation's column chain; one
        # Rayleigh-Ritz eigensolve of order `block` at the measured 19 us per column + its two GEMMs
        n = G.shape[0]
        gemm = 2.0 * n * n * block / 30e12
        orth = 2.0 * 2 * n * block * block / 30e12 + 19e-6 * block
        rr = 19e-6 * block + 2 * 2.0 * n * block * block / 30e12
        run["modelled_b200_ms"] = round(1e3 * (mults * gemm + len(log) * orth + rr), 1)
        out["runs"].append(run)
        print(json.dumps(run), flush=True)
    os.makedirs(os.path.join(ROOT, "profiles", "r02_logs"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_logs", f"probe_subspace_iteration_{kind}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
