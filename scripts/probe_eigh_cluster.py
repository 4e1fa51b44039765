"""GPU probe: the hand-written batched eigensolver (fm_eigh_topk_batched) against cuSOLVER at the path's chunk size —
time per stage-2 group (15 x 999^2, top 250) and for the batches other rank counts leave per GPU."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
out = {}
s = K.Solver()


def graded(n, cols, batch):
    X = torch.randn((batch, n, cols), device=dev, dtype=torch.float64)
    w = torch.ones(cols, device=dev, dtype=torch.float64)
    w[:64] = torch.logspace(2.3, 0.3, 64, device=dev, dtype=torch.float64)
    X = X * w
    X = X - X.mean(dim=2, keepdim=True)
    return X @ X.transpose(-1, -2)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for n, k, batch in ((999, 250, 15), (999, 250, 8), (999, 250, 2), (999, 250, 1), (999, 250, 30), (999, 250, 120), (499, 250, 15)):
    G0 = graded(n, 4000, batch)
    G = G0.clone()
    scratch = torch.empty(K.eigh_topk_scratch_bytes(n, batch, k), dtype=torch.uint8, device=dev)

    def mine():
        G.copy_(G0)
        return s.eigh_topk_batched(G, k, scratch=scratch)

    def copy_only():
        G.copy_(G0)

    t_copy = timed(copy_only)
    t = timed(mine) - t_copy
    lam, W = mine()
    torch.cuda.synchronize()
    ref = torch.linalg.eigvalsh(G0[0])[-k:]
    err = float((lam[0] - ref).abs().max() / ref[-1])
    R = G0[0] @ W[0].T - W[0].T * lam[0][None, :]
    res = float(R.abs().max() / ref[-1])
    orth = float((W[0] @ W[0].T - torch.eye(k, device=dev, dtype=torch.float64)).abs().max())

    def lib():
        G.copy_(G0)
        return s.syevd_batched(G) if batch >= 5 else [s.syevd(G[i]) for i in range(batch)]

    t_lib = timed(lib, reps=2) - t_copy if batch <= 30 else None
    s.check()
    key = f"n{n}_k{k}_batch{batch}"
    out[key] = {"cluster_ms": t, "cusolver_ms": t_lib, "lam_err": err, "residual": res, "orth": orth}
    print(key, out[key], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_eigh_cluster.json", "w"), indent=1)
