"""3xTF32 tcgen05 GEMM (fm_gemm_tn) against a float64 torch matmul and against the library's own
plain-FP32 reference kernel.  Tolerance: 3xTF32 keeps ~22 mantissa bits per product, so the result
must sit within (2e-6 + 1e-8 K) * sum|a||b| of the exact value: tcgen05 truncates the fp32
accumulator at every MMA, which adds a bias that grows with the length of the K chain (the pipeline
cuts long contractions into short chains and reduces them in float64, see svd.KCHAIN_*)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(rows, cols, dev, seed, ld_pad=0):
    from facemap_b200 import kernels as K

    g = torch.Generator(device="cpu").manual_seed(seed)
    m = K.alloc_matrix(rows, cols, dev)
    m.copy_(torch.randn((rows, cols), generator=g) * (1 + torch.rand((rows, 1), generator=g) * 5))
    return m


SHAPES = [
    # M, N, K
    (128, 128, 32), (128, 256, 64), (100, 90, 50), (250, 700, 999), (999, 999, 1280), (300, 500, 4100),
    (64, 24, 8), (513, 129, 333), (129, 3750, 257),
]


@pytest.mark.parametrize("impl", [0, 2, 3, 4, 5, 7, 8, 9])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp64(cuda_device, M, N, K, impl):
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    A, B = _rand(M, K, dev, 1), _rand(N, K, dev, 2)
    C = Kn.alloc_matrix(M, N, dev, zero=True)
    Kn.gemm_tn(A, B, C, impl=impl)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    # the tensor cores truncate the fp32 accumulator on every MMA, so the bound grows with the chain
    bound = (A.double().abs() @ B.double().abs().T) * (2e-6 + (0.0 if (impl == 0 and N > 128) else 1e-8 * K)) + 1e-30
    err = (C.double() - ref).abs()
    assert bool((err <= bound).all()), f"max err/bound {float((err / bound).max()):.3g}"
    # and against the FP32 FFMA reference kernel of the library
    C2 = Kn.alloc_matrix(M, N, dev, zero=True)
    Kn.gemm_tn(A, B, C2, impl=1)
    torch.cuda.synchronize()
    assert bool(((C2.double() - ref).abs() <= bound * 4).all())


def test_tf32_operand_truncation_makes_hi_rewrite_redundant(cuda_device):
    """kind::tf32 ignores the low 13 mantissa bits of an operand word: feeding the raw fp32 tile as the
    hi operand (impl 9, single-accumulator kernel) is bit-identical to masking it first (impl 5 / 6)."""
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    for (M, N, K, a, b) in [(300, 700, 2049, 9, 5), (200, 100, 777, 9, 6)]:
        A, B = _rand(M, K, dev, 11), _rand(N, K, dev, 12)
        C0, C1 = Kn.alloc_matrix(M, N, dev), Kn.alloc_matrix(M, N, dev)
        Kn.gemm_tn(A, B, C0, impl=a)
        Kn.gemm_tn(A, B, C1, impl=b)
        torch.cuda.synchronize()
        assert torch.equal(C0, C1)


@pytest.mark.parametrize("M,N,K,ksplit", [(300, 500, 2049, 1), (1000, 100, 640, 1), (400, 500, 4100, 3)])
def test_presplit_b_operand_is_bit_identical(cuda_device, M, N, K, ksplit):
    """B_lo from fm_split_lo staged by TMA == the in-kernel split of B."""
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    A, B = _rand(M, K, dev, 21), _rand(N, K, dev, 22)
    B_lo = Kn.split_lo(B)
    tr = (B.view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(B_lo, B - tr)
    if ksplit == 1:
        C0, C1 = Kn.alloc_matrix(M, N, dev), Kn.alloc_matrix(M, N, dev)
        Kn.gemm_tn(A, B, C0)
        Kn.gemm_tn(A, B, C1, B_lo=B_lo)
    else:
        C0 = torch.zeros((ksplit, M, Kn.round_up(N, 32)), device=dev)
        C1 = torch.zeros_like(C0)
        Kn.gemm_tn(A, B, ksplit=ksplit, workspace=C0)
        Kn.gemm_tn(A, B, ksplit=ksplit, workspace=C1, B_lo=B_lo)
    torch.cuda.synchronize()
    assert torch.equal(C0, C1)


def test_fully_presplit_gram_is_bit_identical(cuda_device):
    """A_lo = B_lo = split_lo(X) staged by TMA (no splitter warps) == in-kernel split, for the Gram shape."""
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    n, K, ksplit = 999, 4100, 5
    X = _rand(n, K, dev, 31)
    X_lo = Kn.split_lo(X)
    ld = Kn.round_up(n, 32)
    W0 = torch.zeros((ksplit, n, ld), device=dev)
    W1 = torch.zeros_like(W0)
    Kn.gemm_tn(X, X, ksplit=ksplit, workspace=W0, upper_only=True)
    Kn.gemm_tn(X, X, ksplit=ksplit, workspace=W1, upper_only=True, B_lo=X_lo, A_lo=X_lo)
    torch.cuda.synchronize()
    assert torch.equal(W0, W1)
    A, B = _rand(300, 777, dev, 32), _rand(200, 777, dev, 33)
    C0, C1 = Kn.alloc_matrix(300, 200, dev), Kn.alloc_matrix(300, 200, dev)
    Kn.gemm_tn(A, B, C0)
    Kn.gemm_tn(A, B, C1, B_lo=Kn.split_lo(B), A_lo=Kn.split_lo(A))
    torch.cuda.synchronize()
    assert torch.equal(C0, C1)


@pytest.mark.parametrize("mode", ["split", "b_lo", "both"])
def test_chained_accumulation_removes_truncation_bias(cuda_device, mode):
    """The product kernel (impl 0) bounds every TMEM chain to 128 K and sums chains in float32 registers with
    round-to-nearest: on an all-positive 19200-long contraction (the worst case for the truncating accumulator)
    the relative error must stay below 5e-6, where one plain chain (impl 9) is off by more than 1e-4."""
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    M, N, K = 256, 512, 19200
    g = torch.Generator().manual_seed(5)
    A, B = Kn.alloc_matrix(M, K, dev), Kn.alloc_matrix(N, K, dev)
    A.copy_(torch.rand((M, K), generator=g) + 0.5)
    B.copy_(torch.rand((N, K), generator=g) + 0.5)
    kw = {}
    if mode in ("b_lo", "both"):
        kw["B_lo"] = Kn.split_lo(B)
    if mode == "both":
        kw["A_lo"] = Kn.split_lo(A)
    C9, C0 = Kn.alloc_matrix(M, N, dev), Kn.alloc_matrix(M, N, dev)
    Kn.gemm_tn(A, B, C9, impl=0, **kw)        # chained accumulation (product)
    Kn.gemm_tn(A, B, C0, impl=9, **kw)        # one TMEM chain
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    e9 = float(((C9.double() - ref).abs() / ref).max())
    e0 = float(((C0.double() - ref).abs() / ref).max())
    assert e9 < 5e-6, e9
    assert e0 > 1e-4, e0          # documents why the chains exist
    # ragged shape, scale / bias epilogue, split over blockIdx.z
    A, B = _rand(333, 2500, dev, 41), _rand(700, 2500, dev, 42)
    rs, rb = torch.rand(333, device=dev) + 0.5, torch.randn(333, device=dev)
    C = Kn.alloc_matrix(333, 700, dev)
    Kn.gemm_tn(A, B, C, rscale=rs, rbias=rb, impl=0)
    ws = torch.full((3, 333, 704), float("nan"), device=dev)
    Kn.gemm_tn(A, B, ksplit=3, workspace=ws, impl=0)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    bound = (A.double().abs() @ B.double().abs().T) * 3e-6
    assert bool(((C.double() - (ref * rs.double()[:, None] + rb.double()[:, None])).abs() <= bound * 2 + 1e-6).all())
    assert bool(((ws[:, :, :700].double().sum(0) - ref).abs() <= bound).all())


def test_gemm_scale_bias_epilogue(cuda_device):
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    M, N, K = 250, 1280, 999
    A, B = _rand(M, K, dev, 3), _rand(N, K, dev, 4)
    rs = torch.rand(M, device=dev) + 0.5
    rb = torch.randn(M, device=dev)
    C = Kn.alloc_matrix(M, N, dev)
    Kn.gemm_tn(A, B, C, rscale=rs, rbias=rb)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T) * rs.double()[:, None] + rb.double()[:, None]
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 3e-6     # chained accumulation


@pytest.mark.parametrize("n,K,ksplit", [(999, 4096, 7), (300, 2000, 3), (1000, 640, 20), (3750, 1100, 2)])
def test_gram_splitk_upper_assemble(cuda_device, n, K, ksplit):
    """split-K partials + fm_gram_assemble == centred Gram in float64 (the svdecon core)."""
    from facemap_b200 import kernels as Kn

    dev = cuda_device
    X = _rand(n, K, dev, 5)
    ld = Kn.round_up(n, 32)
    ws = torch.full((ksplit, n, ld), float("nan"), dtype=torch.float32, device=dev)
    Kn.gemm_tn(X, X, ksplit=ksplit, workspace=ws, upper_only=True)
    mean = Kn.row_means(X)
    G = Kn.gram_assemble(ws, min(ksplit, (K + 31) // 32), n, mean, K, True)
    torch.cuda.synchronize()
    Xd = X.double()
    md = Xd.mean(dim=1)
    assert float((mean - md).abs().max()) < 1e-12 * float(Xd.abs().max()) * K
    ref = Xd @ Xd.T - K * torch.outer(md, md)
    scale = float(torch.diagonal(Xd @ Xd.T).max())
    assert bool(torch.isfinite(G).all())
    # in-kernel chains of 128 K: <= 48 truncating accumulations x 5.4e-8 = 2.6e-6 of an entry
    assert float((G - ref).abs().max()) / scale < (5e-6 if n > 128 else 3e-5)
    assert torch.equal(G, G.T)
