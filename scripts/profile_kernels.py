"""Launches the hot kernels of the default path at production shape (BASELINE configs[1]) a few times: the target of
the `ncu --set full` captures under profiles/.

  K1          fm_bin_motion_planes: 2368 frames of 640x480, sbin 4, integer energies as two byte planes
  projection  fm_project_i8: 18944 x 500 x 19200 (2 frame planes x 3 mask digits, kind::i8)
  Gram        fm_gemm_i8_tn: one digit-plane product of the merge Gram, 3750 x 3750 x 19200 (upper), and of a chunk
              Gram, 999 x 999 x 19200 (upper, split-K)
  eigensolve  fm_eigh_topk_batched: 15 x 999^2, top 250
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from facemap_b200 import kernels as K

dev = torch.device("cuda", 0)
H, W, sbin, T = 480, 640, 4, 2368
p = (H // sbin) * (W // sbin)
g = torch.Generator(device=dev).manual_seed(0)
frames = torch.randint(0, 256, (T, H, W), dtype=torch.uint8, device=dev, generator=g)
ldq = K.round_up(p, 16)
planes = torch.empty((2, 18944, ldq), dtype=torch.uint8, device=dev)
prev = torch.zeros(p, dtype=torch.int32, device=dev)
last = torch.zeros_like(prev)
energy = torch.zeros(T, dtype=torch.int64, device=dev)
for _ in range(3):
    K.bin_motion_planes(frames, sbin, prev_sum=prev, q_mot=(planes[0, :T, :p], planes[1, :T, :p]), energy=energy, last_sum=last)
planes[0].random_(0, 16, generator=g)
planes[1].random_(0, 256, generator=g)
Ut = K.alloc_matrix(500, p, dev)
Ut.copy_(torch.randn((500, p), device=dev, generator=g) / p ** 0.5)
Q0, Q1, Q2, exps = K.slice_rows_i8(Ut)
bias = K.row_dot(Ut, torch.rand(p, device=dev, generator=g, dtype=torch.float64))
V = torch.empty((18944, 500), dtype=torch.float32, device=dev)
for _ in range(3):
    K.project_i8((planes[0, :, :p], planes[1, :, :p]), (Q0, Q1, Q2), exps, 1.0 / 16, bias, V)
for n, ks in ((3750, 1), (999, 10)):
    A = torch.randint(0, 256, (n, ldq), dtype=torch.uint8, device=dev, generator=g)[:, :p]
    ldp = K.round_up(n, 4)
    P = torch.empty((ks, n, ldp), dtype=torch.int32, device=dev)
    for _ in range(3):
        if ks > 1:
            K.gemm_i8_tn(A, A, ksplit=ks, workspace=P, upper_only=True)
        else:
            K.gemm_i8_tn(A, A, P[0], upper_only=True)
X = torch.randn((15, 999, 3000), device=dev, dtype=torch.float64, generator=g)
G0 = X @ X.transpose(1, 2)
s = K.Solver()
for _ in range(2):
    s.eigh_topk_batched(G0.clone(), 250)
torch.cuda.synchronize()
print("ok")
