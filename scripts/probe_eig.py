"""GPU probe: cuSOLVER syevd timings at the path's sizes: serial, batched, multi-stream."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import kernels as K
dev = torch.device("cuda", 0)
s = K.Solver()
n, B = 999, 15
X = torch.randn(B, n, 4000, device=dev, dtype=torch.float64)
G0 = X @ X.transpose(1, 2)
ref = torch.linalg.eigvalsh(G0[3])
for rep in range(2):
    G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
    lams = [s.syevd(G[i]) for i in range(B)]
    torch.cuda.synchronize(); print(f"serial {B} x syevd({n}) fp64: {(time.perf_counter()-t)*1e3:.1f} ms", flush=True)
try:
    for rep in range(2):
        G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
        lam = s.syevd_batched(G)
        torch.cuda.synchronize(); print(f"batched {B} x syev({n}) fp64: {(time.perf_counter()-t)*1e3:.1f} ms  lam err {float((lam[3]-ref).abs().max()/ref.abs().max()):.2e}", flush=True)
    # eigenvector check: G0[3] v = lam v for the top one
    v = G[3][n - 1]; print("  top eigvec residual", float((G0[3] @ v - lam[3][n - 1] * v).norm() / lam[3][n - 1]))
except Exception as e:
    print("batched failed:", e)
# multi-stream / multi-thread: cusolver calls release the GIL (ctypes), one solver + stream per thread
for nthreads in (2, 4, 8):
    solvers = [K.Solver() for _ in range(nthreads)]
    streams = [torch.cuda.Stream() for _ in range(nthreads)]
    G = G0.clone(); torch.cuda.synchronize()
    def work(tid):
        with torch.cuda.stream(streams[tid]):
            for i in range(tid, B, nthreads):
                solvers[tid].syevd(G[i])
    for rep in range(2):
        G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        [x.start() for x in th]; [x.join() for x in th]
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{nthreads} threads/streams {B} x syevd({n}) fp64: {dt*1e3:.1f} ms", flush=True)
n = 3750
X = torch.randn(n, 6000, device=dev, dtype=torch.float64); G0 = X @ X.T
for rep in range(2):
    G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter(); s.syevd(G); torch.cuda.synchronize()
    print(f"syevd({n}) fp64: {(time.perf_counter()-t)*1e3:.1f} ms", flush=True)
# batched API on the single merge-sized matrix
try:
    for rep in range(2):
        G = G0.clone().unsqueeze(0).contiguous(); torch.cuda.synchronize(); t = time.perf_counter()
        lam = s.syevd_batched(G); torch.cuda.synchronize()
        print(f"XsyevBatched({n}) batch=1 fp64: {(time.perf_counter()-t)*1e3:.1f} ms", flush=True)
except Exception as e:
    print("batched 3750 failed:", e)
# range solve: only the top 500 of 3750
try:
    for rep in range(2):
        G = G0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
        lam = s.syevdx_top(G, 500); torch.cuda.synchronize()
        print(f"syevdx_top({n}, 500) fp64: {(time.perf_counter()-t)*1e3:.1f} ms", flush=True)
    ref = torch.linalg.eigvalsh(G0)[-500:]
    v = G[499]
    print("  lam err", float((lam - ref).abs().max() / ref.abs().max()), " top eigvec residual", float((G0 @ v - lam[499] * v).norm() / lam[499]),
          " orth", float((G[:500] @ G[:500].T - torch.eye(500, device=dev, dtype=torch.float64)).abs().max()))
except Exception as e:
    print("syevdx failed:", e)
