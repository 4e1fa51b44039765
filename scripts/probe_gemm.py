"""GPU probe: tcgen05 3xTF32 GEMM variants vs float64, with layout diagnostics.  Usage:
    python scripts/probe_gemm.py <impl> [quick]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import kernels as K

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda", 0)
torch.manual_seed(0)

def run(M, N, Kd, ksplit=1, show=False):
    A = K.alloc_matrix(M, Kd, dev); B = K.alloc_matrix(N, Kd, dev)
    A.copy_(torch.randn(M, Kd)); B.copy_(torch.randn(N, Kd))
    C = K.alloc_matrix(M, N, dev, zero=True)
    K.gemm_tn(A, B, C, impl=impl)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    err = (C.double() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print(f"impl {impl} M{M} N{N} K{Kd}: max rel err {rel:.3e}", flush=True)
    if rel > 1e-4 and show:
        bad = (err > 1e-3 * ref.abs().max()).nonzero()
        print("  bad entries:", bad.shape[0], "of", M * N, " first:", bad[:8].tolist())
        print("  C[0,:8]  ", C[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
        # which single-precision variant does it look like?
        tf = lambda x: (x.view(torch.int32) & -8192).view(torch.float32)
        r1 = tf(A).double() @ tf(B).double().T
        print("  vs 1xTF32 rel", float((C.double() - r1).abs().max() / ref.abs().max()))
    return rel

ok = True
for shp in [(128, 128, 32), (128, 256, 32), (128, 256, 64), (128, 256, 128), (256, 512, 256), (100, 90, 50), (250, 700, 999)]:
    ok &= run(*shp, show=True) < 1e-5
if len(sys.argv) > 2 and sys.argv[2] == "quick":
    sys.exit(0 if ok else 1)
# timing on the production shapes
PRESPLIT = len(sys.argv) > 2 and sys.argv[2] == "presplit"
def bench(M, N, Kd, ksplit=1, reps=5, upper=False):
    A = K.alloc_matrix(M, Kd, dev); B = K.alloc_matrix(N, Kd, dev)
    A.normal_(); B.normal_()
    Blo = K.split_lo(B) if PRESPLIT else None
    Alo = K.split_lo(A) if (len(sys.argv) > 3 and sys.argv[3] == "both") else None
    C = K.alloc_matrix(M, N, dev)
    ws = torch.empty((ksplit, M, K.round_up(N, 32)), device=dev) if ksplit > 1 else None
    f = (lambda: K.gemm_tn(A, B, ksplit=ksplit, workspace=ws, upper_only=upper, impl=impl, B_lo=Blo, A_lo=Alo)) if ksplit > 1 else (lambda: K.gemm_tn(A, B, C, impl=impl, upper_only=upper, B_lo=Blo, A_lo=Alo))
    for _ in range(2): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = 2.0 * M * N * Kd * (0.5 if upper else 1.0)
    print(f"impl {impl} M{M} N{N} K{Kd} ksplit{ksplit} upper{int(upper)}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s useful (x3 issued = {3*fl/ms/1e9:.0f})", flush=True)
bench(18944, 500, 19200)
bench(18944, 500, 19200, ksplit=16)
bench(999, 999, 19200, ksplit=16, upper=True)
bench(250, 19200, 999)
bench(3750, 3750, 19200, ksplit=2, upper=True)
bench(500, 19200, 3750)
sys.exit(0 if ok else 1)
