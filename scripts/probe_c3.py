"""GPU probe for BASELINE configs[3] (1280x1024, sbin 1): the integer projection at K = 1 310 720, the reference-arithmetic
motion sums (fm_motion_ref, sbin 1 fast path), and cusolverDnXsyevBatched on the single merge-sized matrix."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from facemap_b200 import kernels as K, svd

dev = torch.device("cuda", 0)
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


g = torch.Generator(device=dev).manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "proj"):
    for (T, p, k) in ((4480, 1310720, 500), (4480, 262144, 500), (1152, 1310720, 500)):
        ld = K.round_up(p, 16)
        lo = torch.randint(0, 256, (T, ld), dtype=torch.uint8, device=dev, generator=g)[:, :p]
        U = K.alloc_matrix(k, p, dev)
        U.copy_(torch.randn((k, p), device=dev, generator=g) / p ** 0.5)
        Q0, Q1, Q2, exps = K.slice_rows_i8(U)
        bias = K.row_dot(U, torch.rand(p, device=dev, generator=g, dtype=torch.float64))
        V = torch.empty((T, k), dtype=torch.float32, device=dev)
        for impl in (0, 1):
            ms = timed(lambda: K.project_i8((None, lo), (Q0, Q1, Q2), exps, 1.0, bias, V, impl=impl))
            out[f"project_i8_T{T}_p{p}_impl{impl}"] = {"ms": ms, "algorithmic_tflops": 2.0 * T * p * k / (ms * 1e-3) / 1e12}
            print(f"project_i8 T={T} p={p} impl={impl}: {ms:.2f} ms, {2.0 * T * p * k / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic", flush=True)
        del lo, U, Q0, Q1, Q2, V
        torch.cuda.empty_cache()
if which in ("all", "motion"):
    T, H, W = 512, 1024, 1280
    frames = torch.randint(0, 256, (T, H, W), dtype=torch.uint8, device=dev, generator=g)
    plan = K.pairwise_plan(H * W, dev)
    blocks = svd.motion_ref_blocks(H * W, plan[0].numel(), 1)
    scratch = K.motion_ref_scratch(H, W, 1, plan[0].numel(), blocks, dev)
    res = torch.zeros(T, dtype=torch.float64, device=dev)
    ms = timed(lambda: K.motion_ref(frames, 1, plan, scratch, blocks, res))
    out["motion_ref_1280x1024_512frames"] = {"ms": ms, "us_per_frame": 1e3 * ms / T, "blocks": blocks,
                                             "GBps_frames": (T * H * W) / (ms * 1e-3) / 1e9}
    print("motion_ref sbin 1:", out["motion_ref_1280x1024_512frames"], flush=True)
    gplan = K.pairwise_plan(H * W, dev, groups=True)
    ldq = K.round_up(H * W, 16)
    pl2 = torch.empty((T, ldq), dtype=torch.uint8, device=dev)
    e2 = torch.zeros(T, dtype=torch.int64, device=dev)
    res2 = torch.zeros(T, dtype=torch.float64, device=dev)
    for label, pp in (("leaf-by-leaf fold", plan), ("exact-group fold", gplan)):
        ms2 = timed(lambda: K.motion_ref_planes(frames, pp, scratch, blocks, res2, pl2[:, :H * W], energy=e2))
        out["motion_ref_planes_" + label.replace(" ", "_")] = {"ms": ms2, "us_per_frame": 1e3 * ms2 / T,
                                                               "GBps_frames_read_once_plus_plane": (2 * T * H * W) / (ms2 * 1e-3) / 1e9}
        print(f"fm_motion_ref_planes ({label}): {ms2:.2f} ms, {1e3 * ms2 / T:.2f} us per frame; equal to fm_motion_ref:",
              bool(torch.equal(res2[:T - 1], res[:T - 1])), flush=True)
    fr = frames[:5].cpu().numpy().astype(np.float32).reshape(5, -1)
    ref = np.abs(np.diff(fr, axis=0)).sum(axis=-1)
    got = res[:4].cpu().numpy()
    print("  bit-exact vs numpy float32 pairwise:", np.array_equal(got, ref.astype(np.float64)))
    ld = K.round_up(H * W, 16)
    pl = torch.empty((T, ld), dtype=torch.uint8, device=dev)
    energy = torch.zeros(T, dtype=torch.int64, device=dev)
    prev = torch.zeros(H * W, dtype=torch.int32, device=dev)
    ms = timed(lambda: K.bin_motion_planes(frames, 1, prev_sum=prev, q_mot=(None, pl[:, :H * W]), energy=energy))
    out["k1_planes_1280x1024_512frames"] = {"ms": ms, "GBps_algorithmic": T * H * W * 2 / (ms * 1e-3) / 1e9}
    print("K1 planes sbin 1:", out["k1_planes_1280x1024_512frames"], flush=True)
    del frames, pl
if which in ("all", "eig"):
    s = K.Solver()
    n = 3750
    X = torch.randn(n, 6000, device=dev, dtype=torch.float64, generator=g)
    G1 = X @ X.T
    for label, fn in (("syevdx_top500", lambda G: s.syevdx_top(G, 500)), ("syevBatched_batch1", lambda G: s.syevd_batched(G.unsqueeze(0)))):
        try:
            G = G1.clone()

            def run():
                G.copy_(G1)
                fn(G)
            out["merge_" + label + "_ms"] = timed(run, reps=2)
            print("merge", label, out["merge_" + label + "_ms"], "ms", flush=True)
        except Exception as e:      # noqa: BLE001
            print("merge", label, "failed:", repr(e)[:200])
    s.check()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/probe_c3_{which}.json", "w"), indent=1)
