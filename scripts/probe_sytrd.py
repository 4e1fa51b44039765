"""GPU probe: how cuSOLVER's 70 ms for syevdx(3750, top 500) split between the tridiagonalisation (cusolverDnDsytrd,
called directly through ctypes) and the rest (tridiagonal eigensolve + back-transformation)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
from facemap_b200 import _lib
_lib.load()                      # libfacemap_b200.so brings libcusolver in (RTLD_GLOBAL)
lib = C.CDLL(None)
h = C.c_void_p()
assert lib.cusolverDnCreate(C.byref(h)) == 0
for n in (999, 3750):
    X = torch.randn(n, 6000, device=dev, dtype=torch.float64)
    G0 = X @ X.T
    G = G0.clone()
    d = torch.empty(n, dtype=torch.float64, device=dev)
    e = torch.empty(n, dtype=torch.float64, device=dev)
    tau = torch.empty(n, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    lwork = C.c_int()
    P = C.c_void_p
    rc = lib.cusolverDnDsytrd_bufferSize(h, 0, n, P(G.data_ptr()), n, P(d.data_ptr()), P(e.data_ptr()), P(tau.data_ptr()), C.byref(lwork))
    assert rc == 0, rc
    work = torch.empty(lwork.value, dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(3):
        G.copy_(G0)
        torch.cuda.synchronize()
        t = time.perf_counter()
        rc = lib.cusolverDnDsytrd(h, 0, n, P(G.data_ptr()), n, P(d.data_ptr()), P(e.data_ptr()), P(tau.data_ptr()),
                                  P(work.data_ptr()), lwork.value, P(info.data_ptr()))
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
        assert rc == 0 and int(info.item()) == 0
    s = K.Solver()
    bx = 1e9
    k = 500 if n == 3750 else 250
    for _ in range(2):
        G.copy_(G0)
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.syevdx_top(G, k)
        torch.cuda.synchronize()
        bx = min(bx, (time.perf_counter() - t) * 1e3)
    s.check()
    print(f"n = {n}: cusolverDnDsytrd {best:.1f} ms; syevdx(top {k}) {bx:.1f} ms -> everything after the tridiagonalisation {bx - best:.1f} ms", flush=True)
