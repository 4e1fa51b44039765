"""Edge cases of the frame count / geometry against the live oracle (exact svdecon): videos shorter than
one SVD chunk, exactly on and just past the 500 / 1000 / 2000-frame boundaries of the reference's chunking
(process.py:134-141, 394-395), frame sizes that do not divide by sbin, sbin == 1, three views."""
import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc

pytestmark = pytest.mark.gpu


def _video(T, H, W, seed):
    from facemap_b200.synth import SynthVideo

    return SynthVideo(H, W, T, seed=seed, nmodes=12).all_frames()


def _compare(p, o, motSVD=True, movSVD=False, s_floor=0.0):
    for key in ("avgframe", "avgmotion", "motion"):
        assert len(p[key]) == len(o[key]), key
        for a, b in zip(p[key], o[key]):
            assert a.shape == b.shape and np.array_equal(a, b), key
    for on, mk, vk, sk in ((motSVD, "motMask", "motSVD", "motSv"), (movSVD, "movMask", "movSVD", "movSv")):
        assert len(p[mk]) == len(o[mk]) and len(p[vk]) == len(o[vk])
        assert np.asarray(p[sk]).shape == np.asarray(o[sk]).shape
        if not on:
            continue
        for U, Ur, V, Vr in zip(p[mk], o[mk], p[vk], o[vk]):
            assert U.shape == Ur.shape and V.shape == Vr.shape, (mk, U.shape, Ur.shape, V.shape, Vr.shape)
            nrm, nrm_r = np.linalg.norm(U.astype(np.float64), axis=0), np.linalg.norm(Ur.astype(np.float64), axis=0)
            k = min(3, U.shape[1])
            cos = np.abs((U[:, :k].astype(np.float64) * Ur[:, :k]).sum(axis=0)) / (nrm[:k] * nrm_r[:k])
            assert np.abs(cos - 1).max() < 1e-5, (mk, cos)
            err = np.abs(np.abs(V[:, :k]) - np.abs(Vr[:, :k])).max()
            assert err <= 1e-3 * np.abs(Vr[:, :k]).max(), (vk, err)
        S, Sr = np.asarray(p[sk], np.float64), np.asarray(o[sk], np.float64)
        if np.any(Sr):
            # every singular value to 1e-4 relative; numerically zero ones (rank-deficient inputs) come out of a
            # float64 Gram matrix as sqrt(n * 2^-53) ~ 1e-6 of the largest and are only held to that
            big = Sr >= s_floor * Sr[0]
            assert np.all(np.abs(S - Sr)[big] <= np.maximum(1e-4 * Sr[big], 2e-6 * Sr[0])), \
                (sk, float((np.abs(S - Sr)[big] / np.maximum(Sr[big], 1e-300)).max()))


@pytest.mark.parametrize("N", [120, 499, 500, 501, 999, 1000, 1001, 1999, 2000, 2001, 3217])
def test_frame_count_boundaries(cuda_device, N):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    v = _video(N, 32, 48, seed=N % 7)
    o = orc.run_oracle([v], sbin=2, motSVD=True, movSVD=True, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=2, motSVD=True, movSVD=True)
    _compare(p, o, True, True)
    assert p["motion"][0].shape == (N,) and p["motSVD"][0].shape[0] == N
    if N > 1:
        assert np.array_equal(p["motSVD"][0][0], p["motSVD"][0][1])


@pytest.mark.parametrize("H,W,sbin", [(33, 47, 2), (50, 70, 3), (41, 64, 4), (24, 32, 1), (64, 48, 8), (37, 53, 5)])
def test_odd_geometry_and_bin_factors(cuda_device, H, W, sbin):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    N = 2100
    v = _video(N, H, W, seed=H)
    o = orc.run_oracle([v], sbin=sbin, motSVD=True, movSVD=False, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=sbin, motSVD=True, movSVD=False)
    if sbin & (sbin - 1) == 0:
        _compare(p, o)
    else:   # non power-of-two bins: the reference's float64 means of means are replayed (fm_mean_ref, fm_motion_ref)
        for key in ("avgframe", "avgmotion", "motion"):
            assert np.array_equal(p[key][0], o[key][0]), key
        S, Sr = p["motSv"].astype(np.float64), o["motSv"].astype(np.float64)
        # the float operand of a non-power-of-two bin is the reference's to <= 1 ulp, not bit for bit: singular values
        # move by up to that relative to the LARGEST one
        assert np.all(np.abs(S - Sr) <= np.maximum(1e-4 * Sr, 4e-6 * Sr[0]))


def test_three_views_canvas(cuda_device):
    """3 simultaneous views of different sizes: concatenated pixel axis + 1x2 canvas rule (utils.py:703-748)."""
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    N = 2050
    vids = [_video(N, 32, 48, 1), _video(N, 40, 40, 2), _video(N, 24, 64, 3)]
    o = orc.run_oracle(vids, sbin=2, motSVD=True, movSVD=False, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device) for v in vids]), sbin=2)
    _compare(p, o)
    for key in ("sybin", "sxbin", "Lybin", "Lxbin"):
        assert np.array_equal(np.asarray(p[key]), np.asarray(o[key])), key
    assert int(p["LYbin"]) == int(o["LYbin"]) and int(p["LXbin"]) == int(o["LXbin"])
    assert p["motMask_reshape"][0].shape == o["motMask_reshape"][0].shape
    assert np.array_equal(p["avgframe_reshape"], o["avgframe_reshape"])
    # the canvas of the first mask is the scatter of its rows
    k0 = np.abs(p["motMask_reshape"][0][..., 0]).sum()
    assert np.isclose(k0, np.abs(p["motMask"][0][:, 0]).sum(), rtol=1e-6)
    # laid out on the device, it is exactly what utils.multivideo_reshape (utils.py:622-630) makes of the mask matrix
    from facemap_b200.utils import binned_inds, multivideo_reshape

    _, _, iinds = binned_inds([32, 40, 24], [48, 40, 64], 2)
    want = multivideo_reshape(p["motMask"][0], p["LYbin"], p["LXbin"], p["sybin"], p["sxbin"], p["Lybin"], p["Lxbin"], iinds)
    assert np.array_equal(p["motMask_reshape"][0], want)


def test_fullsvd_off_with_roi(cuda_device):
    """fullSVD=False + one motion ROI: placeholders of process.py:153-160, 336-339 and the ROI outputs."""
    import copy

    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from tests.helpers import motion_roi

    N = 2100
    v = _video(N, 48, 64, 5)
    rois = [motion_roi(0, 4, 40, 6, 50)]
    o = orc.run_oracle([v], sbin=2, motSVD=True, movSVD=False, rois=copy.deepcopy(rois), fullSVD=False, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=2, motSVD=True,
                               movSVD=False, rois=copy.deepcopy(rois), fullSVD=False)
    assert p["motMask"][0].shape == (0, 1) and p["motSVD"][0].shape == (0, 1) and p["motion"][0].shape == (0,)
    assert p["motMask"][1].shape == o["motMask"][1].shape and p["motSVD"][1].shape == o["motSVD"][1].shape
    assert np.array_equal(p["motion"][1], o["motion"][1])
    S, Sr = p["motSv"].astype(np.float64), o["motSv"].astype(np.float64)
    nz = Sr > 0
    assert S.shape == Sr.shape and (np.abs(S - Sr)[nz] / Sr[nz]).max() <= 1e-4         # every singular value
    U, Ur = p["motMask"][1][:, :3].astype(np.float64), o["motMask"][1][:, :3].astype(np.float64)
    assert np.abs(np.abs((U * Ur).sum(axis=0)) - 1).max() < 1e-5


@pytest.mark.parametrize("with_rois", [False, True])
def test_movie_svd_without_motion_svd(cuda_device, with_rois):
    """motSVD off, movSVD on: the reference's merge loop runs over len(U_mot) (process.py:264), so the movie targets
    beyond it are never merged — all of them without ROIs, the last ROI with ROIs (quirk Q8, pinned against the live
    reference in tests/test_oracle_vs_live_reference.py)"""
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    from .helpers import motion_roi

    v = _video(2100, 36, 44, seed=5)
    rois = [motion_roi(0, 4, 16, 4, 20), motion_roi(0, 20, 34, 10, 40)] if with_rois else None
    o = orc.run_oracle([v], sbin=2, motSVD=False, movSVD=True, rois=rois, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=2, motSVD=False,
                               movSVD=True, rois=rois)
    _compare(p, o, False, True)
    shapes = [tuple(u.shape) for u in p["movMask"]]
    # full frame merged (395 = npix - 1), first ROI merged (40 pixels -> 39), last ROI left as 2 x 90 stacked chunk components
    assert shapes == ([(396, 500)] if not with_rois else [(396, 395), (40, 39), (90, 180)]), shapes
    assert [tuple(u.shape) for u in p["motMask"]] == [tuple(u.shape) for u in o["motMask"]]
