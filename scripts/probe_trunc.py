"""GPU probe: (1) does kind::tf32 ignore the low 13 operand bits (impl 5 == impl 0 bit for bit)?
(2) which way does the fp32 accumulator round (sign of the error for all-positive vs all-negative sums)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import kernels as K
dev = torch.device("cuda", 0)
torch.manual_seed(1)
M, N, Kd = 256, 512, 4096
A = K.alloc_matrix(M, Kd, dev); B = K.alloc_matrix(N, Kd, dev)
A.copy_(torch.randn(M, Kd)); B.copy_(torch.randn(N, Kd))
C0 = K.alloc_matrix(M, N, dev); C5 = K.alloc_matrix(M, N, dev)
K.gemm_tn(A, B, C0, impl=0); K.gemm_tn(A, B, C5, impl=5); torch.cuda.synchronize()
print("impl0 == impl5 bit for bit:", bool(torch.equal(C0, C5)), " max |diff|", float((C0 - C5).abs().max()))
ref = A.double() @ B.double().T
print("impl0 rel err", float((C0.double() - ref).abs().max() / ref.abs().max()), "impl5 rel err", float((C5.double() - ref).abs().max() / ref.abs().max()))
for sign in (1.0, -1.0):
    A.copy_(torch.rand(M, Kd) * sign + 0.5 * sign); B.copy_(torch.rand(N, Kd) + 0.5)
    K.gemm_tn(A, B, C0, impl=0); torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    rel = (C0.double() - ref) / ref.abs()
    print(f"sums of sign {sign:+.0f}: mean (C-ref)/|ref| = {float(rel.mean()):+.3e}  min {float(rel.min()):+.3e} max {float(rel.max()):+.3e}")
