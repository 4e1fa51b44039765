"""GPU probe: whole path vs the live oracle with an error profile.  Usage:
    python scripts/probe_pipeline.py <gemm_impl> <eig_fp32>"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facemap_b200 import process, kernels as K
from facemap_b200.frames import DeviceArraySource
from facemap_b200.synth import SynthVideo
from oracle import facemap_oracle as orc
from tests.helpers import subspace_sin, sign_align

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eig32 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda", 0)
H, W, N, sbin = 96, 128, 2600, 2
v = SynthVideo(H, W, N, seed=5).all_frames()
t = time.time(); o = orc.run_oracle([v], sbin=sbin, motSVD=True, movSVD=True, svd="exact"); print("oracle s", time.time() - t)
timing = {}
t = time.time()
p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(dev)]), sbin=sbin, motSVD=True, movSVD=True,
                           gemm_impl=impl, eig_fp32=eig32, timing=timing)
print("cuda s", time.time() - t)
print("device ms:", {k: round(x, 2) for k, x in sorted(timing["device_ms"].items())})
print("avgframe exact", np.array_equal(p["avgframe"][0], o["avgframe"][0]), "avgmotion exact", np.array_equal(p["avgmotion"][0], o["avgmotion"][0]),
      "motion exact", np.array_equal(p["motion"][0], o["motion"][0]))
for mk, vk, sk in (("motMask", "motSVD", "motSv"), ("movMask", "movSVD", "movSv")):
    S, Sr = p[sk].astype(np.float64), o[sk].astype(np.float64)
    rel = np.abs(S - Sr) / Sr
    print(mk, "S ratio at 10,50,100,250,499:", [float(Sr[i] / Sr[0]) for i in (10, 50, 100, 250, 499)])
    print(mk, "S rel err max over [0:10],[0:50],[0:100],[0:250],[0:500]:", [float(rel[:k].max()) for k in (10, 50, 100, 250, 500)])
    U, Ur = p[mk][0], o[mk][0]
    print(mk, "orth err", float(np.abs(U.astype(np.float64).T @ U.astype(np.float64) - np.eye(U.shape[1])).max()))
    for k in (1, 2, 3, 5, 8, 12, 16, 24, 32, 64):
        print(mk, f"k={k} sin(angle)={subspace_sin(U[:, :k], Ur[:, :k]):.2e}", end="; ")
    print()
    Ua, sg = sign_align(U[:, :32], Ur[:, :32])
    V, Vr = p[vk][0][:, :32] * sg[None, :], o[vk][0][:, :32]
    tr = np.linalg.norm(V - Vr, axis=0) / np.linalg.norm(Vr, axis=0)
    print(mk, "sign agree", int((sg > 0).sum()), "/32 trace rel err first 16:", np.array2string(tr[:16], precision=1))
