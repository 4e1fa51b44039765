"""Triage of the integer tensor-core kernels on their first GPU run (fm_gemm_i8_tn, fm_project_i8 impl 0 / 1).

The opt-in tests only say "differs"; this prints HOW, with structured operands whose wrong answers point at one layer:
  K coverage      ones x ramp(k)            -> sum k            (a K step skipped or read twice: descriptor advance)
  row mapping     A[m, 0] = m % 128         -> C[m, n] = m%128  (TMEM lane / swizzle row mix-up)
  column mapping  B[n, 0] = n % 128         -> C[m, n] = n%128  (N placement, column halves of the epilogue warps)
  signedness      0xFF as s8 / u8                               (a_format / b_format bits of the instruction descriptor)
  class weights   one frame plane x one digit plane at a time   (accumulator per weight class, Horner weights, scaling)
Every case runs in its own try block and reports instead of asserting.  python scripts/triage_i8.py [impl ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facemap_b200 import kernels as K  # noqa: E402

dev = torch.device(os.environ.get("FM_TRIAGE_DEVICE", "cuda:0"))      # "cpu" only for the dry run with stand-in kernels


def pitched(a):
    rows, cols = a.shape
    buf = torch.zeros((rows, -(-cols // 16) * 16), dtype=torch.from_numpy(a[:0]).dtype, device=dev)
    buf[:, :cols] = torch.from_numpy(a).to(dev)
    return buf[:, :cols]


def report(tag, got, want):
    got, want = np.asarray(got), np.asarray(want)
    bad = got != want
    if not bad.any():
        print(f"ok    {tag}")
        return
    i = np.argwhere(bad)[0]
    print(f"FAIL  {tag}: {bad.sum()} of {bad.size} wrong; first at {tuple(int(x) for x in i)}: got {got[tuple(i)]} want "
          f"{want[tuple(i)]}; rows wrong {int(bad.any(axis=1).sum())}/{bad.shape[0]}, cols wrong "
          f"{int(bad.any(axis=0).sum())}/{bad.shape[1]}; got values {np.unique(got)[:8]} want {np.unique(want)[:8]}")


def guarded(tag, fn):
    try:
        fn()
        if dev.type == "cuda":
            torch.cuda.synchronize()
    except Exception as e:                                   # noqa: BLE001 - triage prints everything
        print(f"ERROR {tag}: {type(e).__name__}: {str(e)[:300]}")
        return False
    return True


def gemm_cases():
    for (M, N, Kd) in ((128, 256, 32), (128, 256, 64), (128, 256, 448), (300, 520, 200)):
        def run(A, B, tag, want):
            C = torch.full((M, -(-N // 4) * 4), -7, dtype=torch.int32, device=dev)
            if guarded(tag, lambda: K.gemm_i8_tn(pitched(A), pitched(B), C)):
                report(tag, C[:, :N].cpu().numpy(), want)
        ones_a, ones_b = np.ones((M, Kd), np.uint8), np.ones((N, Kd), np.uint8)
        ramp = np.tile((np.arange(Kd) % 200).astype(np.uint8), (N, 1))
        run(ones_a, ramp, f"gemm_i8 {M}x{N}x{Kd} K coverage", np.full((M, N), int(ramp[0].astype(np.int64).sum())))
        a = np.zeros((M, Kd), np.uint8)
        a[:, 0] = np.arange(M) % 128
        run(a, ones_b, f"gemm_i8 {M}x{N}x{Kd} row mapping", np.tile((np.arange(M) % 128)[:, None], (1, N)))
        b = np.zeros((N, Kd), np.uint8)
        b[:, 0] = np.arange(N) % 128
        run(ones_a, b, f"gemm_i8 {M}x{N}x{Kd} column mapping", np.tile((np.arange(N) % 128)[None, :], (M, 1)))
        ff = np.full((M, Kd), -1, np.int8)
        run(ff, ones_b, f"gemm_i8 {M}x{N}x{Kd} A = 0xFF as s8", np.full((M, N), -Kd))
        run(ff.view(np.uint8), ones_b, f"gemm_i8 {M}x{N}x{Kd} A = 0xFF as u8", np.full((M, N), 255 * Kd))
        run(ones_a, np.full((N, Kd), -1, np.int8), f"gemm_i8 {M}x{N}x{Kd} B = 0xFF as s8", np.full((M, N), -Kd))


def project_cases(impl):
    for (M, N, Kd) in ((128, 128, 64), (256, 128, 64), (300, 500, 1000), (256, 256, 16384 + 128)):
        exps = torch.full((N,), 7, dtype=torch.int32, device=dev)          # 2^(ex - 7) = 1: V = acc / 2^16
        zeros_d, zeros_q = np.zeros((M, Kd), np.uint8), np.zeros((N, Kd), np.uint8)
        for d in (0, 1):
            for e in (0, 1, 2):
                Dp = [zeros_d, zeros_d]
                Dp[d] = np.ones((M, Kd), np.uint8)
                Q = [zeros_q.view(np.int8), zeros_q, zeros_q]
                Q[e] = np.ones((N, Kd), np.int8 if e == 0 else np.uint8)
                V = torch.full((M, -(-N // 4) * 4), float("nan"), dtype=torch.float32, device=dev)
                tag = f"project_i8 impl {impl} {M}x{N}x{Kd} plane {d} x digit {e}"
                want = np.float32(Kd * 256.0 ** (1 - d) * 256.0 ** (-e))
                if guarded(tag, lambda: K.project_i8((pitched(Dp[0]), pitched(Dp[1])), tuple(pitched(q) for q in Q), exps,
                                                     1.0, None, V, impl=impl)):
                    report(tag, V[:, :N].cpu().numpy(), np.full((M, N), want, np.float32))
        # row / column mapping through the low plane and the leading digit
        d = np.zeros((M, Kd), np.uint8)
        d[:, 0] = np.arange(M) % 128
        q0 = np.zeros((N, Kd), np.int8)
        q0[:, 0] = 1
        V = torch.full((M, -(-N // 4) * 4), float("nan"), dtype=torch.float32, device=dev)
        tag = f"project_i8 impl {impl} {M}x{N}x{Kd} row mapping"
        if guarded(tag, lambda: K.project_i8((None, pitched(d)), (pitched(q0), pitched(zeros_q), pitched(zeros_q)), exps, 1.0,
                                             None, V, impl=impl)):
            report(tag, V[:, :N].cpu().numpy(), np.tile((np.arange(M) % 128)[:, None], (1, N)).astype(np.float32))
        q0[:, 0] = np.arange(N) % 100
        d[:, 0] = 1
        tag = f"project_i8 impl {impl} {M}x{N}x{Kd} column mapping"
        if guarded(tag, lambda: K.project_i8((None, pitched(d)), (pitched(q0), pitched(zeros_q), pitched(zeros_q)), exps, 1.0,
                                             None, V, impl=impl)):
            report(tag, V[:, :N].cpu().numpy(), np.tile((np.arange(N) % 100)[None, :], (M, 1)).astype(np.float32))


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "0", "1"]
    if "gemm" in what:
        gemm_cases()
    for impl in (0, 1):
        if str(impl) in what:
            project_cases(impl)
