"""Timing of fm_project_i8 (kind::i8, exact) against the 3xTF32 projection at the BASELINE configs[1] batch shape
(p = 19 200, 500 components) and the configs[3] crop (sbin 1: one frame plane).  Run on a B200:
  python scripts/probe_project_i8.py  ->  one JSON line per shape.  EXPERIMENTAL kernel, see DESIGN.md §7.2(a)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facemap_b200 import kernels as K  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def main():
    dev = torch.device("cuda", 0)
    for (T, p, k, sbin) in ((18944, 19200, 500, 4), (18944, 19200, 500, 1), (4736, 262144, 500, 1)):
        g = torch.Generator(device=dev).manual_seed(0)
        ld = K.round_up(p, 16)
        lo = torch.randint(0, 256, (T, ld), dtype=torch.uint8, device=dev, generator=g)[:, :p]
        hi = torch.randint(0, 16, (T, ld), dtype=torch.uint8, device=dev, generator=g)[:, :p] if sbin > 1 else None
        U = K.alloc_matrix(k, p, dev)
        U.copy_(torch.randn((k, p), device=dev, generator=g) / p ** 0.5)
        Q0, Q1, Q2, exps = K.slice_rows_i8(U)
        bias = K.row_dot(U, torch.rand(p, device=dev, generator=g, dtype=torch.float64))
        V = torch.empty((T, k), dtype=torch.float32, device=dev)
        ms_i8 = timed(lambda: K.project_i8((hi, lo), (Q0, Q1, Q2), exps, 1.0 / sbin ** 2, bias, V))
        Vp = torch.empty((T, k), dtype=torch.float32, device=dev)
        ms_pair = timed(lambda: K.project_i8((hi, lo), (Q0, Q1, Q2), exps, 1.0 / sbin ** 2, bias, Vp, impl=1))
        same = bool(torch.equal(V, Vp))
        X = K.alloc_matrix(T, p, dev)
        X.copy_(lo.float() + (0 if hi is None else 256 * hi.float()))
        Ulo = K.split_lo(U)
        V2 = torch.empty((T, k), dtype=torch.float32, device=dev)
        ms_tf = timed(lambda: K.gemm_tn(X, U, V2, B_lo=Ulo, impl=0))
        flop = 2.0 * T * p * k
        print(json.dumps({"T": T, "p": p, "k": k, "sbin": sbin, "ms_i8": ms_i8, "ms_i8_pair": ms_pair, "pair_equals_single": same,
                          "ms_tf32x3": ms_tf,
                          "algorithmic_tflops_i8": flop / ms_i8 / 1e9, "algorithmic_tflops_i8_pair": flop / ms_pair / 1e9, "algorithmic_tflops_tf32x3": flop / ms_tf / 1e9,
                          "operand_bytes_i8": T * p * (1 if hi is None else 2) + 3 * k * p,
                          "operand_bytes_tf32x3": 4 * T * p + 8 * k * p}), flush=True)


if __name__ == "__main__":
    main()
