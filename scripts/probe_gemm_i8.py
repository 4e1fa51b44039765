"""GPU probe (round-2 opener): times the one-product kernels written without a GPU at the end of round 1 —
fm_gemm_i8_tn (kind::i8) and fm_gemm_tn impl 10 (kind::tf32, one MMA per K step) — next to the product 3xTF32 kernel
(impl 0, pre-split planes) on the shapes of BASELINE configs[1], and prints bytes-staged and MMA-issue rates so the
L2-fed-or-MMA-bound question can be read off.  Exactness is covered by the opt-in tests; this only
checks one entry per run.

    python scripts/probe_gemm_i8.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def timed(f, reps=5):
    for _ in range(2):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def tiles(M, N, upper):
    mt, nt = -(-M // 128), -(-N // 256)
    return sum(1 for i in range(mt) for j in range(nt) if (not upper) or (j + 1) * 256 - 1 >= i * 128)


def report(tag, ms, M, N, Kd, upper, bytes_per_k_per_tile, mma_per_k32):
    t = tiles(M, N, upper)
    staged = t * Kd * bytes_per_k_per_tile                     # shared-memory fill traffic (L2 -> SMEM)
    mmas = t * (Kd / 32) * mma_per_k32                         # 128x256 MMA instructions (K = 32 bytes each)
    print(f"{tag:34s} M{M} N{N} K{Kd} upper{int(upper)}: {ms:7.3f} ms   staged {staged / ms / 1e9:6.2f} TB/s   "
          f"{mmas / ms / 1e6:7.1f} G MMA/s   {2.0 * M * N * Kd * (0.5 if upper else 1) / ms / 1e9:7.1f} TFLOP/s (one product)",
          flush=True)


def run(M, N, Kd, upper, ksplit):
    # --- i8 ---
    A8 = torch.randint(0, 256, (M, K.round_up(Kd, 16)), dtype=torch.uint8, device=dev)[:, :Kd]
    B8 = torch.randint(-128, 128, (N, K.round_up(Kd, 16)), dtype=torch.int8, device=dev)[:, :Kd]
    ldc = K.round_up(N, 4)
    ks8 = max(ksplit, -(-Kd // 32768))
    if ks8 > 1:
        W8 = torch.empty((ks8, M, ldc), dtype=torch.int32, device=dev)
        f8 = lambda: K.gemm_i8_tn(A8, B8, ksplit=ks8, workspace=W8, upper_only=upper)
    else:
        C8 = torch.empty((M, ldc), dtype=torch.int32, device=dev)
        f8 = lambda: K.gemm_i8_tn(A8, B8, C8, upper_only=upper)
    ms = timed(f8)
    got = (W8.sum(0) if ks8 > 1 else C8)[0, N - 1].item()
    want = int((A8[0].long() * B8[N - 1].long()).sum().item())
    assert got == want, ("i8 entry", got, want)
    report(f"i8 one product (ksplit {ks8})", ms, M, N, Kd, upper, (128 + 256) * 1, 1)
    # --- tf32, one product (impl 10) and the product 3xTF32 kernel with pre-split planes (impl 0) ---
    A = K.alloc_matrix(M, Kd, dev)
    B = K.alloc_matrix(N, Kd, dev)
    A.normal_()
    B.normal_()
    Alo, Blo = K.split_lo(A), K.split_lo(B)
    ldw = K.round_up(N, 32)
    kk = max(ksplit, 2)
    W = torch.empty((kk, M, ldw), dtype=torch.float32, device=dev)
    if N > 128:
        ms = timed(lambda: K.gemm_tn(A, B, ksplit=kk, workspace=W, upper_only=upper, impl=10))
        report(f"tf32 one product, impl 10 (ksplit {kk})", ms, M, N, Kd, upper, (128 + 256) * 4, 4)
    ms = timed(lambda: K.gemm_tn(A, B, ksplit=kk, workspace=W, upper_only=upper, impl=0, B_lo=Blo, A_lo=Alo))
    report(f"3xTF32 pre-split, impl 0 (ksplit {kk})", ms, M, N, Kd, upper, (128 + 256) * 8, 12)


for shape in ((3750, 3750, 19200, True, 1), (3750, 3750, 19200, False, 1), (999, 999, 19200, True, 10),
              (18944, 500, 19200, False, 1), (2368, 500, 19200, False, 4)):
    run(*shape)
