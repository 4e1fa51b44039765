"""GPU probe: the merge eigenproblem (top 500 of the 3 750 x 3 750 centred Gram of the stacked chunk components,
facemap/process.py:264-295) by block subspace iteration with Rayleigh-Ritz instead of cusolverDnXsyevdx.

The CPU study (scripts/probe_subspace_iteration.py) shows that on the benchmark's spectrum 6 multiplications by G with
a block of 750 vectors (4 with 1 000) bring all 500 singular values within 2e-5 and every gap-resolved vector within
1e-7.  This probe builds the same problem with the product's own stages on a B200, times the pieces such a solver is
made of — float64 GEMM (cuBLAS through torch.matmul), Cholesky-QR (syrk + potrf + trsm), the Rayleigh-Ritz eigensolve of
the projected 750 x 750 problem on cuSOLVER and on the hand-written cluster solver (fm_eigh_topk_batched) — and the whole
loop against torch.linalg.eigh / the path's syevdx, with the path's accuracy gates.  Research probe: nothing of it is on
the product path.

usage: python scripts/probe_merge_subspace.py [nframes]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facemap_b200 import svd                         # noqa: E402
from facemap_b200.frames import DeviceArraySource    # noqa: E402
from facemap_b200.synth import SynthVideo            # noqa: E402
from tests.helpers import gap_resolved               # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
K_KEEP = 500
out = {}


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), r


def merge_gram(nframes):
    H, W, sbin = 480, 640, 4
    frames = SynthVideo(H, W, nframes, seed=0).all_frames_torch(dev)
    src = DeviceArraySource([frames])
    geo = svd.Geometry([H], [W], sbin, None)
    timer = svd.StageTimer(enabled=False)
    ws = svd.SvdWorkspace(dev, timer=timer)
    src.plan_access(*svd.access_plan(src, None), dev)
    avgframe, avgmotion = svd.stage1_means(src, geo, dev, timer)
    Yt, ks, nsegs = svd.stage2_chunks(src, geo, avgframe, avgmotion, ["mot"], True, ws, dev, timer)
    Y = Yt["mot"][0].double()                                   # [3750, p]: (U S)^T of the 15 chunks
    ws.solver.check()
    Yc = Y - Y.mean(dim=1, keepdim=True)
    G = Yc @ Yc.T
    del frames, src, Yt, Y, Yc
    torch.cuda.empty_cache()
    return (G + G.T) / 2, ws


def cholqr(Z, passes):
    for _ in range(passes):
        S = Z.T @ Z
        L = torch.linalg.cholesky(S)
        Z = torch.linalg.solve_triangular(L, Z.T, upper=False).T.contiguous()
    return Z


def subspace(G, block, mults, passes, rr, ws, seed=0):
    """`mults` multiplications by G, Cholesky-QR (`passes` passes) after each, then Rayleigh-Ritz (rr = 'cusolver' |
    'cluster') -> (lam [500] descending, W [n, 500])"""
    n = G.shape[0]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Q = cholqr(torch.randn((n, block), dtype=torch.float64, device=dev, generator=g), 1)
    for _ in range(mults):
        Q = cholqr(G @ Q, passes)
    GQ = G @ Q
    Hm = Q.T @ GQ
    Hm = (Hm + Hm.T) / 2
    if rr == "cusolver":
        th, Z = torch.linalg.eigh(Hm)
        th, Z = th[-K_KEEP:].flip(0), Z[:, -K_KEEP:].flip(1)
    else:
        lam, Wt = ws.solver.eigh_topk_batched(Hm[None].contiguous(), K_KEEP)       # ascending, rows = vectors
        th, Z = lam[0].flip(0), Wt[0].flip(0).T
    return th, Q @ Z


def accuracy(lam_ref, W_ref, th, Wr):
    S, S_ref = th.clamp_min(0).sqrt().cpu().numpy(), lam_ref[:K_KEEP].clamp_min(0).sqrt().cpu().numpy()
    s_err = float(np.max(np.abs(S - S_ref) / S_ref))
    res = np.flatnonzero(gap_resolved(S_ref, 1e-3)[0])
    cosv = (Wr[:, res] * W_ref[:, res]).sum(0).abs().cpu().numpy()
    ang = float(np.max(np.sqrt(np.maximum(0.0, 1.0 - cosv ** 2)))) if res.size else 0.0
    return s_err, ang, int(res.size)


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
    t0 = time.time()
    G, ws = merge_gram(nframes)
    n = G.shape[0]
    out["problem"] = {"n": n, "nframes": nframes, "build_s": round(time.time() - t0, 1)}
    # the reference solves
    t_eigh, (lam_all, W_all) = timed(lambda: torch.linalg.eigh(G), reps=2)
    lam_ref, W_ref = lam_all.flip(0), W_all.flip(1)
    S = lam_ref.clamp_min(0).sqrt()
    out["spectrum_S_k_over_S_0"] = {str(k): float(S[k] / S[0]) for k in (10, 100, 250, 499, 600, 750, 1000, 2000, n - 1)}
    out["torch_linalg_eigh_3750_ms"] = round(t_eigh, 2)
    t_sx, _ = timed(lambda: ws.solver.syevdx_top(G.clone(), K_KEEP), reps=2)
    out["syevdx_top500_ms_incl_clone"] = round(t_sx, 2)
    # the pieces
    pieces = {}
    for b in (750, 1000):
        Q = torch.randn((n, b), dtype=torch.float64, device=dev)
        t, Z = timed(lambda: G @ Q)
        pieces[f"dgemm_{n}x{n}x{b}_ms"] = round(t, 3)
        pieces[f"dgemm_{n}x{n}x{b}_tflops"] = round(2.0 * n * n * b / (t * 1e-3) / 1e12, 1)
        t, Sm = timed(lambda: Z.T @ Z)
        pieces[f"syrk_as_gemm_{b}_ms"] = round(t, 3)
        t, L = timed(lambda: torch.linalg.cholesky(Sm))
        pieces[f"cholesky_{b}_ms"] = round(t, 3)
        t, _ = timed(lambda: torch.linalg.solve_triangular(L, Z.T, upper=False))
        pieces[f"trsm_{b}_ms"] = round(t, 3)
        t, _ = timed(lambda: torch.linalg.qr(Z), reps=2)
        pieces[f"householder_qr_{n}x{b}_ms"] = round(t, 3)
        Hm = Q.T @ (G @ Q)
        Hm = (Hm + Hm.T) / 2
        t, _ = timed(lambda: torch.linalg.eigh(Hm), reps=2)
        pieces[f"rr_eigh_cusolver_{b}_ms"] = round(t, 3)
        try:
            t, _ = timed(lambda: ws.solver.eigh_topk_batched(Hm[None].clone(), K_KEEP), reps=2)
            pieces[f"rr_eigh_cluster_top500_of_{b}_ms"] = round(t, 3)
            ws.solver.check()
        except Exception as e:      # noqa: BLE001
            pieces[f"rr_eigh_cluster_top500_of_{b}_ms"] = repr(e)[:200]
    out["pieces"] = pieces
    print(json.dumps(out), flush=True)
    # the whole loop
    runs = []
    for block, mults, passes, rr in ((750, 4, 1, "cusolver"), (750, 6, 1, "cusolver"), (750, 6, 2, "cusolver"),
                                     (750, 6, 1, "cluster"), (750, 8, 1, "cluster"), (1000, 4, 1, "cluster"),
                                     (1000, 4, 2, "cusolver"), (600, 10, 1, "cluster")):
        try:
            t, (th, Wr) = timed(lambda: subspace(G, block, mults, passes, rr, ws), reps=2)
            ws.solver.check()
            s_err, ang, nres = accuracy(lam_ref, W_ref, th, Wr)
            orth = float((Wr.T @ Wr - torch.eye(K_KEEP, dtype=torch.float64, device=dev)).abs().max())
            run = {"block": block, "multiplications": mults, "cholqr_passes": passes, "rayleigh_ritz": rr,
                   "ms": round(t, 2), "sv_rel_err_all_500": s_err, "worst_resolved_vector_angle": ang,
                   "resolved": nres, "orthogonality": orth, "gates_hold": bool(s_err <= 1e-4 and ang <= 1e-3)}
        except Exception as e:      # noqa: BLE001
            run = {"block": block, "multiplications": mults, "cholqr_passes": passes, "rayleigh_ritz": rr,
                   "error": repr(e)[:200]}
        runs.append(run)
        print(json.dumps(run), flush=True)
    out["runs"] = runs
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_merge_subspace.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
