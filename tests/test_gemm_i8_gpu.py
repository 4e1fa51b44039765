"""fm_gemm_i8_tn (tcgen05.mma kind::i8, int32 accumulation) against exact integer arithmetic.

WRITTEN WITHOUT A GPU (round 1 had spent its GPU budget): opt-in with FM_EXPERIMENTAL=1 until it has run on a B200.
Exact equality everywhere — integer work; ragged shapes, both signednesses, split-K partials, upper-triangle skipping,
a K range long enough to approach the int32 limit."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FM_EXPERIMENTAL") != "1", reason="opt-in until validated on a B200")]


def _operand(rs, rows, K, signed, device, extreme=False):
    lo, hi = (-128, 128) if signed else (0, 256)
    a = rs.randint(lo, hi, size=(rows, K))
    if extreme:
        a[:] = hi - 1 if not signed else lo
    ld = -(-K // 16) * 16
    buf = torch.zeros((rows, ld), dtype=torch.int8 if signed else torch.uint8, device=device)
    buf[:, :K] = torch.from_numpy(a.astype(np.int8 if signed else np.uint8)).to(device)
    return a.astype(np.int64), buf[:, :K]


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 1000), (999, 999, 19200), (130, 257, 77), (1, 129, 33)])
@pytest.mark.parametrize("a_signed,b_signed", [(False, False), (False, True), (True, True)])
def test_gemm_i8_exact(cuda_device, M, N, K, a_signed, b_signed):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(M + 7 * N + K)
    a, A = _operand(rs, M, K, a_signed, cuda_device)
    b, B = _operand(rs, N, K, b_signed, cuda_device)
    want = a @ b.T
    ldc = -(-N // 4) * 4
    C = torch.full((M, ldc), -77, dtype=torch.int32, device=cuda_device)
    K_.gemm_i8_tn(A, B, C)
    assert np.array_equal(C[:, :N].cpu().numpy().astype(np.int64), want)
    ks = 3
    if K >= 64 * ks:
        ws = torch.full((ks, M, ldc), -77, dtype=torch.int32, device=cuda_device)
        K_.gemm_i8_tn(A, B, ksplit=ks, workspace=ws)
        assert np.array_equal(ws[:, :, :N].cpu().numpy().astype(np.int64).sum(axis=0), want)


def test_gemm_i8_upper_only_and_int32_range(cuda_device):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(0)
    n, K = 700, 32768                                       # 255^2 * 32768 = 2 130 739 200 < 2^31
    a, A = _operand(rs, n, K, False, cuda_device)
    a[5], a[6] = 255, 255
    A[5], A[6] = 255, 255
    C = torch.full((n, 704), -1, dtype=torch.int32, device=cuda_device)
    K_.gemm_i8_tn(A, A, C, upper_only=True)
    got = C[:, :n].cpu().numpy().astype(np.int64)
    want = a @ a.T
    iu = np.triu_indices(n)
    assert np.array_equal(got[iu], want[iu]) and got[5, 6] == 255 * 255 * K
    assert (got[256:, :256] == -1).all()                    # the tile below the diagonal was skipped
    with pytest.raises(Exception):
        K_.gemm_i8_tn(torch.zeros((128, 40000), dtype=torch.uint8, device=cuda_device),
                      torch.zeros((256, 40000), dtype=torch.uint8, device=cuda_device), C)   # would overflow int32
