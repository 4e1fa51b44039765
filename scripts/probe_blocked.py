"""GPU probe: the projection at BASELINE configs[3]'s K = 1 310 720 from plain planes (row pitch 1.31 MB) and from
K-blocked copies (fm_retile_u8 + fm_project_i8_blocked), retiling included."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
T, p, k = 4480, 1310720, 500


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


lo = torch.randint(0, 256, (T, p), dtype=torch.uint8, device=dev, generator=g)
U = K.alloc_matrix(k, p, dev)
U.copy_(torch.randn((k, p), device=dev, generator=g) / p ** 0.5)
Q0, Q1, Q2, exps = K.slice_rows_i8(U)
bias = K.row_dot(U, torch.rand(p, device=dev, generator=g, dtype=torch.float64))
V = torch.empty((T, k), dtype=torch.float32, device=dev)
Vb = torch.empty_like(V)
ms = timed(lambda: K.project_i8((None, lo), (Q0, Q1, Q2), exps, 1.0, bias, V))
print(f"plain planes: {ms:.2f} ms, {2.0 * T * p * k / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic", flush=True)
for kb in (4096, 16384, 1024):
    Qb = tuple(K.retile_u8(q, kb) for q in (Q0, Q1, Q2))
    nblk = -(-p // kb)
    buf = torch.empty((nblk, T, kb), dtype=torch.uint8, device=dev)
    ms_r = timed(lambda: K.retile_u8(lo, kb, out=buf))
    ms_p = timed(lambda: K.project_i8_blocked((None, buf), Qb, exps, 1.0, bias, Vb, p))
    torch.cuda.synchronize()
    print(f"kb {kb:6d}: retile {ms_r:.2f} ms ({2 * T * p / (ms_r * 1e-3) / 1e12:.2f} TB/s), projection {ms_p:.2f} ms = "
          f"{2.0 * T * p * k / (ms_p * 1e-3) / 1e12:.0f} TFLOP/s algorithmic, together {2.0 * T * p * k / ((ms_r + ms_p) * 1e-3) / 1e12:.0f}; "
          f"same bits as plain: {bool(torch.equal(V, Vb))}", flush=True)
    del buf, Qb
    torch.cuda.empty_cache()
