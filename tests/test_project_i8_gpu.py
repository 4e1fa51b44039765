"""fm_project_i8 and fm_bin_motion_planes on the GPU against their NumPy restatement (tests/i8_model.py).

First run on a B200 at the start of round 2 (17 passed for either kernel variant, profiles/r02a_first_gpu_run.md); part of the
regular `-m gpu` run.
Integer stages (planes, digits, class sums) are compared exactly; the projection itself bit for bit, because the model
repeats the kernel's float64 scaling and its one rounding to float32; the whole path in FM_PROJ_MODE=i8 against the
default path and the oracle within the north-star tolerance for projected traces (1e-3 relative; observed ~1e-6)."""
import os

import numpy as np
import pytest
import torch

from . import i8_model as M8

pytestmark = pytest.mark.gpu


# FM_TEST_I8_IMPLS=0 or =1 runs one kernel variant per process (a trapped kernel poisons the CUDA context for every test
# after it, so the first GPU run keeps the two apart)
IMPLS = [int(x) for x in os.environ.get("FM_TEST_I8_IMPLS", "0,1").split(",")]


def _pitched(a, device, pitch=16):
    rows, cols = a.shape
    buf = torch.zeros((max(rows, 1), -(-cols // pitch) * pitch), dtype=torch.from_numpy(a[:0]).dtype, device=device)
    buf[:rows, :cols] = torch.from_numpy(a).to(device)
    return buf[:rows, :cols]


def _masks(rs, k, p):
    U = rs.standard_normal((k, p)).astype(np.float32) * np.exp(rs.uniform(-3, 0, size=(1, p))).astype(np.float32)
    return (U / np.linalg.norm(U, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("M,N,K,sbin", [(128, 128, 64, 4), (300, 500, 19200, 4), (77, 130, 1000, 1), (1000, 500, 40000, 4),
                                        (5, 7, 33, 16), (260, 250, 16384 + 64, 2)])
@pytest.mark.parametrize("impl", IMPLS)
def test_project_i8_bit_exact_against_model(cuda_device, M, N, K, sbin, impl):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(M + N + K)
    D = rs.randint(0, 255 * sbin * sbin + 1, size=(M, K))
    D[0] = 255 * sbin * sbin
    U = _masks(rs, N, K)
    U[0] = np.float32(0.03)                                  # a constant row: every digit plane saturated alike
    avg = (D.mean(axis=0) / sbin ** 2).astype(np.float32)
    alpha = 1.0 / sbin ** 2
    hi, lo = M8.planes_of(D)
    Dg = (None if sbin == 1 else _pitched(hi, cuda_device), _pitched(lo, cuda_device))
    Ug = K_.alloc_matrix(N, K, cuda_device)
    Ug.copy_(torch.from_numpy(U))
    Q0, Q1, Q2, exps = K_.slice_rows_i8(Ug)
    m0, m1, m2, mex = M8.mask_digits(U)
    assert np.array_equal(Q0.cpu().numpy(), m0) and np.array_equal(Q1.cpu().numpy(), m1)
    assert np.array_equal(Q2.cpu().numpy(), m2) and np.array_equal(exps.cpu().numpy(), mex)
    bias = K_.row_dot(Ug, torch.from_numpy(avg).to(cuda_device).double())
    # float64 sums in another order: compare against the sum of magnitudes, not against a possibly cancelled value
    assert (np.abs(bias.cpu().numpy() - M8.bias_of(U, avg)) <= 1e-13 * (np.abs(U).astype(np.float64) @ np.abs(avg))).all()
    V = torch.full((M, -(-N // 4) * 4), float("nan"), dtype=torch.float32, device=cuda_device)
    K_.project_i8(Dg, (Q0, Q1, Q2), exps, alpha, bias, V, impl=impl)
    want = M8.project_i8_model(D, U, alpha, bias=bias.cpu().numpy())
    got = V[:, :N].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and without a bias, into an unaligned output
    V2 = torch.full((M, N + 1), float("nan"), dtype=torch.float32, device=cuda_device)
    K_.project_i8(Dg, (Q0, Q1, Q2), exps, alpha, None, V2[:, 1:], impl=impl)
    assert np.array_equal(V2[:, 1:].cpu().numpy().view(np.uint32), M8.project_i8_model(D, U, alpha).view(np.uint32))
    assert torch.isnan(V2[:, 0]).all()


@pytest.mark.parametrize("M,N,K,sbin,kb", [(300, 500, 10000, 4, 4096), (130, 70, 4096, 1, 4096), (260, 250, 16384 + 640, 2, 1024),
                                           (77, 130, 1000, 1, 64)])
@pytest.mark.parametrize("impl", IMPLS)
def test_blocked_projection_equals_plain(cuda_device, M, N, K, sbin, kb, impl):
    """fm_retile_u8 + fm_project_i8_blocked (K-blocked operands for wide views) give the bits of fm_project_i8 on the
    plain planes, whatever garbage the plain planes carry in their pitch padding (the blocked copy zero-fills its tail)"""
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(M + K + kb)
    D = rs.randint(0, 255 * sbin * sbin + 1, size=(M, K))
    hi, lo = M8.planes_of(D)
    pitch = -(-K // 16) * 16 + 32
    planes = torch.randint(0, 256, (2, M, pitch), dtype=torch.uint8, device=cuda_device)      # garbage in the padding
    planes[0, :, :K] = torch.from_numpy(hi).to(cuda_device)
    planes[1, :, :K] = torch.from_numpy(lo).to(cuda_device)
    Dg = (None if sbin == 1 else planes[0, :, :K], planes[1, :, :K])
    Ug = K_.alloc_matrix(N, K, cuda_device)
    Ug.copy_(torch.from_numpy(_masks(rs, N, K)))
    Q0, Q1, Q2, exps = K_.slice_rows_i8(Ug)
    bias = K_.row_dot(Ug, torch.rand(K, device=cuda_device, dtype=torch.float64))
    V = torch.full((M, -(-N // 4) * 4), float("nan"), dtype=torch.float32, device=cuda_device)
    Vb = torch.full_like(V, float("nan"))
    K_.project_i8(Dg, (Q0, Q1, Q2), exps, 1.0 / sbin ** 2, bias, V, impl=impl)
    Db = tuple(None if pl is None else K_.retile_u8(pl, kb) for pl in Dg)
    Qb = tuple(K_.retile_u8(q, kb) for q in (Q0, Q1, Q2))
    nblk = -(-K // kb)
    assert Db[1].shape == (nblk, M, kb) and Qb[0].dtype == torch.int8
    # the blocked copy holds the plane and zeros behind it
    flat = Db[1].permute(1, 0, 2).reshape(M, nblk * kb)
    assert torch.equal(flat[:, :K], planes[1, :, :K]) and not flat[:, K:].any()
    K_.project_i8_blocked(Db, Qb, exps, 1.0 / sbin ** 2, bias, Vb, K, impl=impl)
    torch.cuda.synchronize()
    assert torch.equal(V[:, :N], Vb[:, :N]) and not torch.isnan(Vb[:, :N]).any()


def test_pipeline_with_blocked_projection(cuda_device, monkeypatch):
    """Stage3 takes the blocked route for wide views (plane pitch > PROJ_BLOCKED_MIN_PITCH); forced on a small video, the
    whole path gives the same bits as the plain route"""
    from facemap_b200 import process, svd
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    v = SynthVideo(48, 64, 2300, seed=13).all_frames()
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device)])
    want = process.process_source(src, sbin=1, motSVD=True, movSVD=True)
    monkeypatch.setattr(svd, "PROJ_BLOCKED_MIN_PITCH", 1024)
    monkeypatch.setattr(svd, "PROJ_KB", 512)
    got = process.process_source(src, sbin=1, motSVD=True, movSVD=True)
    for key in ("motSVD", "movSVD", "motMask", "movMask", "motion"):
        for a, b in zip(got[key], want[key]):
            assert np.array_equal(a, b), key


@pytest.mark.parametrize("impl", IMPLS)
def test_project_i8_extreme_digits_fill_the_int32_chain(cuda_device, impl):
    from facemap_b200 import kernels as K_

    M, N, K = 130, 129, 3 * 16384                            # three full chains at the largest digits
    hi = np.full((M, K), 255, np.uint8)
    Dg = (_pitched(hi, cuda_device), _pitched(hi, cuda_device))
    Q0 = _pitched(np.full((N, K), -128, np.int8), cuda_device)
    Q1 = _pitched(np.full((N, K), 255, np.uint8), cuda_device)
    exps = torch.zeros(N, dtype=torch.int32, device=cuda_device)
    V = torch.zeros((M, 132), dtype=torch.float32, device=cuda_device)
    K_.project_i8(Dg, (Q0, Q1, Q1), exps, 1.0, None, V, impl=impl)
    acc = K * 65535 * ((-128 << 16) + (255 << 8) + 255)
    want = np.float32(np.float64(acc) * 2.0 ** -23)
    assert (V[:, :N].cpu().numpy() == want).all()


@pytest.mark.parametrize("sbin,H,W", [(1, 40, 64), (2, 48, 64), (4, 64, 96), (8, 64, 128), (16, 64, 128), (3, 45, 60),
                                      (5, 50, 80), (4, 37, 53)])
def test_bin_motion_planes_exact(cuda_device, sbin, H, W):
    from facemap_b200 import kernels as K_

    rs = np.random.RandomState(sbin * 1000 + W)
    T = 70
    fr = rs.randint(0, 256, size=(T, H, W)).astype(np.uint8)
    fr[5], fr[6] = 255, 0                                     # the largest possible energy
    S = M8.block_sums(fr, sbin)
    npix, col0 = S.shape[1], 16
    frames = torch.from_numpy(fr).to(cuda_device)
    ldq = -(-(col0 + npix) // 16) * 16
    for carry in (False, True):
        prev = torch.from_numpy(S[0].astype(np.int32)).to(cuda_device) if carry else None
        src = frames[1:] if carry else frames
        rows = T - 1
        planes = torch.full((4, rows, ldq), 0xEE, dtype=torch.uint8, device=cuda_device)
        energy = torch.zeros(rows, dtype=torch.int64, device=cuda_device)
        last = torch.zeros(npix, dtype=torch.int32, device=cuda_device)
        hi_ok = sbin > 1
        got_rows = K_.bin_motion_planes(src, sbin, prev_sum=prev, q_mot=(planes[0] if hi_ok else None, planes[1]),
                                        q_mov=(planes[2] if hi_ok else None, planes[3]), col0=col0, energy=energy,
                                        last_sum=last)
        assert got_rows == rows
        d = np.abs(np.diff(S, axis=0))
        P = planes.cpu().numpy().astype(np.int64)
        mot = P[1][:, col0:col0 + npix] + (256 * P[0][:, col0:col0 + npix] if hi_ok else 0)
        mov = P[3][:, col0:col0 + npix] + (256 * P[2][:, col0:col0 + npix] if hi_ok else 0)
        assert np.array_equal(mot, d) and np.array_equal(mov, S[1:])
        assert (P[1][:, :col0] == 0xEE).all() and (P[1][:, col0 + npix:] == 0xEE).all()      # nothing outside the view
        assert np.array_equal(energy.cpu().numpy(), d.sum(axis=1))
        assert np.array_equal(last.cpu().numpy(), S[-1])
    # painted frames (pupil / blink ROIs write 255 into the frames stage 3 bins, facemap/process.py:543, 567)
    mask = np.where(rs.rand(H, W) < 0.2, 255, 0).astype(np.uint8)
    Sp = M8.block_sums(fr | mask[None], sbin)
    planes = torch.zeros((2, T - 1, ldq), dtype=torch.uint8, device=cuda_device)
    K_.bin_motion_planes(frames, sbin, q_mot=(planes[0] if sbin > 1 else None, planes[1]), col0=col0,
                         ormask=torch.from_numpy(mask).to(cuda_device))
    P = planes.cpu().numpy().astype(np.int64)
    assert np.array_equal(P[1][:, col0:col0 + npix] + 256 * P[0][:, col0:col0 + npix], np.abs(np.diff(Sp, axis=0)))


@pytest.mark.parametrize("impl", IMPLS)
def test_pipeline_in_i8_projection_mode(cuda_device, monkeypatch, impl):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    H, W, N, sbin = 96, 128, 2600, 2
    v = SynthVideo(H, W, N, seed=5).all_frames()
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device)])
    p0 = process.process_source(src, sbin=sbin, movSVD=True)
    monkeypatch.setenv("FM_PROJ_MODE", "i8")
    monkeypatch.setenv("FM_PROJ_I8_IMPL", str(impl))
    p1 = process.process_source(src, sbin=sbin, movSVD=True)
    assert np.array_equal(p0["motion"][0], p1["motion"][0])
    for key, sk in (("motSVD", "motSv"), ("movSVD", "movSv")):
        assert np.array_equal(p0[sk], p1[sk])                 # stage 2 is untouched
        a, b = p0[key][0].astype(np.float64), p1[key][0].astype(np.float64)
        rel = np.abs(a - b).max(axis=0) / np.abs(a).max(axis=0)
        print(key, "max relative trace difference, first 16 / all:", rel[:16].max(), rel.max())
        assert rel[:16].max() < 1e-4 and rel.max() < 1e-3


def test_pipeline_in_i8_projection_mode_with_motion_rois(cuda_device, monkeypatch):
    """two cameras, motion + movie SVD, two motion ROIs (tests/helpers.py:multi_roi_s4): the ROI projections come from
    gathered byte planes (fm_gather_roi_u8) and must agree with the default path like the full-frame ones"""
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    from .helpers import CASES, case_rois, case_videos

    name = "multi_roi_s4"
    vids = case_videos(name)
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device) for v in vids])
    sbin = CASES[name][2]
    p0 = process.process_source(src, sbin=sbin, motSVD=True, movSVD=True, rois=case_rois(name))
    monkeypatch.setenv("FM_PROJ_MODE", "i8")
    p1 = process.process_source(src, sbin=sbin, motSVD=True, movSVD=True, rois=case_rois(name))
    for key in ("motSVD", "movSVD"):
        assert len(p0[key]) == len(p1[key]) == 3
        for a, b in zip(p0[key], p1[key]):
            a, b = a.astype(np.float64), b.astype(np.float64)
            rel = np.abs(a - b).max(axis=0) / np.abs(a).max(axis=0)
            print(key, a.shape, "max relative trace difference, first 16 / all:", rel[:16].max(), rel.max())
            assert rel[:16].max() < 1e-4 and rel.max() < 1e-3
    for a, b in zip(p0["motion"], p1["motion"]):
        assert np.array_equal(a, b)
