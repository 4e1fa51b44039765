"""GPU probe: is the projection's slowdown at K = 1 310 720 a matter of the operands' ROW PITCH (one 2 MB page per row:
TLB reach) rather than of K?  Same T x 500 x 262 144 problem with the operands stored at a pitch of 262 144 and of
1 310 720 bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
T, p, k = 4480, 262144, 500


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


U = K.alloc_matrix(k, p, dev)
U.copy_(torch.randn((k, p), device=dev, generator=g) / p ** 0.5)
bias = K.row_dot(U, torch.rand(p, device=dev, generator=g, dtype=torch.float64))
V = torch.empty((T, k), dtype=torch.float32, device=dev)
for pitch_a, pitch_b in ((p, p), (1310720, p), (p, 1310720), (1310720, 1310720), (1310720 + 4096, 1310720 + 4096)):
    lo = torch.randint(0, 256, (T, pitch_a), dtype=torch.uint8, device=dev, generator=g)[:, :p]
    q = torch.empty((3, k, pitch_b), dtype=torch.uint8, device=dev)
    Q0, Q1, Q2, exps = K.slice_rows_i8(U, out=(q[0].view(torch.int8)[:, :p], q[1][:, :p], q[2][:, :p]))
    ms = timed(lambda: K.project_i8((None, lo), (Q0, Q1, Q2), exps, 1.0, bias, V))
    print(f"frame-plane pitch {pitch_a:8d} B, digit pitch {pitch_b:8d} B: {ms:6.2f} ms, {2.0 * T * p * k / (ms * 1e-3) / 1e12:5.0f} TFLOP/s algorithmic", flush=True)
    del lo, q
    torch.cuda.empty_cache()
