"""Small end-to-end workload for `compute-sanitizer --tool memcheck`: every kernel of the library on small
shapes (K1 vector + generic paths, GEMM variants, helpers, whole path with ROIs)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facemap_b200 import kernels as K, process
from facemap_b200.frames import DeviceArraySource
from facemap_b200.synth import SynthVideo
from tests.helpers import motion_roi

dev = torch.device("cuda", 0)
for (H, W, sbin) in [(32, 48, 4), (30, 50, 3), (16, 32, 1)]:
    fr = torch.randint(0, 256, (20, H, W), dtype=torch.uint8, device=dev)
    p = (H // sbin) * (W // sbin)
    X = K.alloc_matrix(19, p, dev)
    e = torch.zeros(19, dtype=torch.int64, device=dev)
    re = torch.zeros((1, 19), dtype=torch.int64, device=dev)
    K.bin_motion(fr, sbin, x_mot=X, x_mov=K.alloc_matrix(19, p, dev), energy=e, rois=[(1, 3, 1, 4)], roi_energy=re)
for impl in (0, 1, 3, 7):
    A = K.alloc_matrix(130, 100, dev); B = K.alloc_matrix(270, 100, dev); A.normal_(); B.normal_()
    C = K.alloc_matrix(130, 270, dev)
    K.gemm_tn(A, B, C, impl=impl)
    ws = torch.zeros((3, 130, 288), device=dev)
    K.gemm_tn(A, B, ksplit=3, workspace=ws, impl=impl)
    K.splitk_reduce(ws, 3, C)
Blo = K.split_lo(B)
K.gemm_tn(A, B, C, B_lo=Blo)
K.gemm_tn(A, B, C, B_lo=Blo, A_lo=K.split_lo(A))
v = SynthVideo(48, 64, 2100, seed=1, nmodes=8).all_frames()
rois = [motion_roi(0, 4, 40, 6, 50)]
p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(dev)]), sbin=4, motSVD=True, movSVD=True, rois=copy.deepcopy(rois))
torch.cuda.synchronize()
print("sanitize workload ok", p["motSVD"][0].shape, p["motSv"].shape)
