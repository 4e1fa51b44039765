"""Launches the two O(N) kernels at production shape (config C2) a few times: the target of the
`ncu --set full` captures.  K1: 2368 frames of 640x480 (sbin 4); projection GEMM: 18944 x 500 x 19200
chained accumulation, pre-split masks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import kernels as K
dev = torch.device("cuda", 0)
H, W, sbin, T = 480, 640, 4, 2368
p = (H // sbin) * (W // sbin)
frames = torch.randint(0, 256, (T, H, W), dtype=torch.uint8, device=dev)
X = K.alloc_matrix(18944, p, dev); X.normal_()
avg = torch.rand(p, device=dev)
prev = torch.zeros(p, dtype=torch.int32, device=dev); last = torch.zeros_like(prev)
energy = torch.zeros(T, dtype=torch.int64, device=dev)
Ut = K.alloc_matrix(500, p, dev); Ut.normal_(); Ut_lo = K.split_lo(Ut)
V = torch.empty((18944, 500), device=dev)
for _ in range(3):
    K.bin_motion(frames, sbin, prev_sum=prev, avgmotion=avg, x_mot=X[:T], energy=energy, last_sum=last)
for _ in range(3):
    K.gemm_tn(X, Ut, V, B_lo=Ut_lo)
torch.cuda.synchronize()
print("ok")
