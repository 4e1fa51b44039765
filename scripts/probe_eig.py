"""GPU probe: cuSOLVER syevd timings (fp64 vs fp32) at the path's sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import kernels as K
dev = torch.device("cuda", 0)
s = K.Solver()
for n in (999, 3750):
    X = torch.randn(n, 4000, device=dev, dtype=torch.float64)
    G0 = X @ X.T
    for fp32 in (False, True):
        for rep in range(3):
            G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
            lam = s.syevd(G, use_fp32=fp32); torch.cuda.synchronize(); dt = time.perf_counter() - t
        ref = torch.linalg.eigvalsh(G0)
        print(f"syevd n={n} fp32={fp32}: {dt*1e3:.1f} ms, max rel lam err {float(((lam-ref).abs()/ref.abs().max()).max()):.2e}", flush=True)
        t = time.perf_counter(); torch.linalg.eigh(G0.float() if fp32 else G0); torch.cuda.synchronize(); print("   torch.linalg.eigh:", round((time.perf_counter()-t)*1e3,1), "ms")
