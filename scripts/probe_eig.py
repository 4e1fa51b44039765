"""GPU probe: cuSOLVER eigensolve timings at the path's sizes — serial, batched, several streams (one solver each;
issued from one thread and from one thread per stream), range solves (top k only), float32.
Writes gpurun_out/probe_eig.json (the source of profiles/r02_eig.md)."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
out = {}


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
    return best


def graded(n, cols, batch=None):
    """Gram matrices with a spectrum like the path's: a few strong components over a flat noise floor"""
    shape = (n, cols) if batch is None else (batch, n, cols)
    X = torch.randn(shape, device=dev, dtype=torch.float64)
    w = torch.ones(cols, device=dev, dtype=torch.float64)
    w[:64] = torch.logspace(2, 0.3, 64, device=dev, dtype=torch.float64)
    X = X * w
    return X @ X.transpose(-1, -2)


n, B = 999, 15
G0 = graded(n, 4000, B)
s = K.Solver()
G = G0.clone()


def serial():
    G.copy_(G0)
    for i in range(B):
        s.syevd(G[i])


out["chunk_serial_syevd_ms"] = timed(serial)
print(f"serial {B} x syevd({n}) fp64: {out['chunk_serial_syevd_ms']:.1f} ms", flush=True)


def serial_x():
    G.copy_(G0)
    for i in range(B):
        s.syevdx_top(G[i], 250)


out["chunk_serial_syevdx250_ms"] = timed(serial_x)
print(f"serial {B} x syevdx({n}, top 250) fp64: {out['chunk_serial_syevdx250_ms']:.1f} ms", flush=True)


def batched():
    G.copy_(G0)
    s.syevd_batched(G)


out["chunk_batched_ms"] = timed(batched)
print(f"batched {B} x syev({n}) fp64: {out['chunk_batched_ms']:.1f} ms", flush=True)

for ns in (2, 4, 8, 15):
    solvers = [K.Solver() for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for variant in ("syevd", "syevdx250"):
        def one_thread():
            G.copy_(G0)
            main = torch.cuda.current_stream()
            for st in streams:
                st.wait_stream(main)
            for i in range(B):
                with torch.cuda.stream(streams[i % ns]):
                    if variant == "syevd":
                        solvers[i % ns].syevd(G[i])
                    else:
                        solvers[i % ns].syevdx_top(G[i], 250)
            for st in streams:
                main.wait_stream(st)

        def work(tid):
            with torch.cuda.stream(streams[tid]):
                for i in range(tid, B, ns):
                    if variant == "syevd":
                        solvers[tid].syevd(G[i])
                    else:
                        solvers[tid].syevdx_top(G[i], 250)

        def threads():
            G.copy_(G0)
            torch.cuda.synchronize()
            th = [threading.Thread(target=work, args=(i,)) for i in range(ns)]
            [x.start() for x in th]
            [x.join() for x in th]

        a, b = timed(one_thread), timed(threads)
        out[f"chunk_{variant}_{ns}streams_one_thread_ms"] = a
        out[f"chunk_{variant}_{ns}streams_threads_ms"] = b
        print(f"{ns} streams, {B} x {variant}({n}): one issuing thread {a:.1f} ms, one thread per stream {b:.1f} ms", flush=True)
    for sv in solvers:
        sv.check()
    del solvers

# correctness of the asynchronous path: eigenvalues of the last multi-stream run against torch
ref = torch.linalg.eigvalsh(G0[3])
G.copy_(G0)
lam = s.syevd(G[3])
s.check()
out["lam_err"] = float((lam - ref).abs().max() / ref.abs().max())
print("lam err", out["lam_err"])

# the merge-sized problem
n = 3750
G1 = graded(n, 6000)
Gm = G1.clone()
for label, fn in (("merge_syevd_ms", lambda: s.syevd(Gm)),
                  ("merge_syevdx_top500_ms", lambda: s.syevdx_top(Gm, 500)),
                  ("merge_syevd_fp32_ms", lambda: s.syevd(Gm, use_fp32=True))):
    def run():
        Gm.copy_(G1)
        fn()
    out[label] = timed(run, reps=2)
    print(f"{label}: {out[label]:.1f} ms", flush=True)
# the merge solve next to concurrent chunk-sized solves on other streams (does the big one slow down?)
solvers = [K.Solver() for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]


def merge_plus_chunks():
    Gm.copy_(G1)
    G.copy_(G0)
    main = torch.cuda.current_stream()
    for st in streams:
        st.wait_stream(main)
    s.syevdx_top(Gm, 500)
    for i in range(8):
        with torch.cuda.stream(streams[i % 4]):
            solvers[i % 4].syevd(G[i])
    for st in streams:
        main.wait_stream(st)


out["merge_top500_plus_8_chunk_solves_ms"] = timed(merge_plus_chunks, reps=2)
print(f"syevdx(3750, 500) with 8 syevd(999) on 4 other streams: {out['merge_top500_plus_8_chunk_solves_ms']:.1f} ms")
# host time of one asynchronous solve (does the call return before the device is done?)
Gm.copy_(G1)
torch.cuda.synchronize()
t = time.perf_counter()
s.syevdx_top(Gm, 500)
out["merge_top500_host_return_ms"] = (time.perf_counter() - t) * 1e3
torch.cuda.synchronize()
out["merge_top500_total_ms"] = (time.perf_counter() - t) * 1e3
print(f"syevdx(3750, 500): host returns after {out['merge_top500_host_return_ms']:.1f} ms, done after {out['merge_top500_total_ms']:.1f} ms")
G.copy_(G0)
torch.cuda.synchronize()
t = time.perf_counter()
s.syevd(G[0])
out["chunk_syevd_host_return_ms"] = (time.perf_counter() - t) * 1e3
torch.cuda.synchronize()
out["chunk_syevd_total_ms"] = (time.perf_counter() - t) * 1e3
print(f"syevd(999): host returns after {out['chunk_syevd_host_return_ms']:.1f} ms, done after {out['chunk_syevd_total_ms']:.1f} ms")
s.check()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_eig.json", "w"), indent=1)
