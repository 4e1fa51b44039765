"""Full-resolution ROI processors on the GPU (fm_paint / fm_blink / fm_running / fm_pupil and their host
driver facemap_b200.fullres) against oracle/roi_oracle.py and the fixtures the unmodified reference produced
(tests/golden/fullres_*.npz, scripts/make_golden_fullres.py).

Gates: paint, running (integer argmax) and integer-valued blink sums bit-exact; float blink sums 1e-12
relative (the order of a float64 sum of exactly representable terms); pupil outputs 1e-9 (relative for
com / area / axlen, absolute for the unit vectors axdir) with the NaN pattern identical; motion energy of
the painted frames bit-exact; projected traces as in test_pipeline_gpu (1e-3 relative)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import roi_oracle
from tests.helpers import FULLRES_CASES, ellipse_mask, eye_video, fullres_case, sign_align

pytestmark = pytest.mark.gpu

PUPIL_TOL = 1e-9
BLINK_RTOL = 1e-12


def _dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def _pupil_close(got, com, area, axdir, axlen, where=""):
    """got: [T, 9] from the kernel; reference arrays from the oracle / the reference fixture."""
    ref = np.concatenate((com, area[:, None], axdir.reshape(-1, 4), axlen), axis=1)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{where}: NaN pattern differs"
    ok = ~np.isnan(ref).any(axis=1)
    g, r = got[ok], ref[ok]
    rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-300)
    cols_rel = [0, 1, 2, 7, 8]
    assert rel[:, cols_rel].max(initial=0) <= PUPIL_TOL, (where, rel[:, cols_rel].max(), np.argmax(rel[:, cols_rel].max(axis=1)))
    assert np.abs(g[:, 3:7] - r[:, 3:7]).max(initial=0) <= PUPIL_TOL, (where, np.abs(g[:, 3:7] - r[:, 3:7]).max())


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(5, 48, 64), (3, 37, 53), (1, 16, 16), (70, 30, 40)])
def test_paint(cuda_device, shape):
    from facemap_b200 import kernels as K

    rs = np.random.RandomState(0)
    fr = rs.randint(0, 256, size=shape).astype(np.uint8)
    m = np.where(rs.rand(*shape[1:]) < 0.3, 255, 0).astype(np.uint8)
    out = torch.empty(shape, dtype=torch.uint8, device=cuda_device)
    got = K.paint(_dev(fr, cuda_device), _dev(m, cuda_device), out).cpu().numpy()
    assert np.array_equal(got, fr | m)


@pytest.mark.parametrize("saturation", [140.25, 100, 0, 255, 254.9, 17.5])
def test_blink_kernel(cuda_device, saturation):
    from facemap_b200 import fullres
    from facemap_b200 import kernels as K

    rs = np.random.RandomState(1)
    T, H, W = 33, 50, 71
    fr = rs.randint(0, 256, size=(T, H, W)).astype(np.uint8)
    rect = (7, 44, 13, 60)
    y0, y1, x0, x1 = rect
    mask = (~ellipse_mask(y1 - y0, x1 - x0)).astype(np.uint8)
    out = torch.zeros(T, dtype=torch.float64, device=cuda_device)
    K.blink(_dev(fr, cuda_device), rect, _dev(mask, cuda_device), _dev(fullres.blink_terms(saturation), cuda_device), out)
    crop = fr[:, y0:y1, x0:x1].copy()
    crop[:, mask > 0] = 255
    ref = roi_oracle.blink_chunk(crop, saturation)
    got = out.cpu().numpy()
    if isinstance(saturation, int):
        assert np.array_equal(got, ref)
    else:
        assert np.allclose(got, ref, rtol=BLINK_RTOL, atol=0)


@pytest.mark.parametrize("geom", [(40, 60, (3, 37, 5, 58)), (96, 128, (10, 95, 3, 100)), (30, 200, (0, 30, 0, 200)),
                                  (300, 40, (2, 290, 1, 39))])
def test_running_kernel_with_carry(cuda_device, geom):
    """two batches with the carried frame in between == one oracle call over all frames"""
    from facemap_b200 import fullres
    from facemap_b200 import kernels as K

    H, W, rect = geom
    y0, y1, x0, x1 = rect
    rs = np.random.RandomState(2)
    T = 23
    base = rs.randint(0, 256, size=(H, W + 3 * T)).astype(np.uint8)
    fr = np.stack([base[:, 2 * t:2 * t + W] for t in range(T)])          # a drifting texture
    fr[5] = 77                                                            # constant frame: mean hits the pixels exactly
    fr[9] = fr[8]
    om = np.where(rs.rand(y1 - y0, x1 - x0) < (0.0 if H == 40 else 0.1), 255, 0).astype(np.uint8)   # H == 40: no paint, so frame 5 ties
    g2 = _dev(fullres.gaussian_fourier_sq(y1 - y0, x1 - x0), cuda_device)
    d_fr, d_om = _dev(fr, cuda_device), _dev(om, cuda_device)
    out = torch.zeros((T, 2), dtype=torch.int32, device=cuda_device)
    means = torch.empty(T, dtype=torch.float32, device=cuda_device)
    cut = 11
    K.running(d_fr[:cut], rect, d_om, g2, means, out[:cut])
    prev_crop = torch.bitwise_or(d_fr[cut - 1, y0:y1, x0:x1], d_om).contiguous()
    prev_mean = means[cut - 1:cut].clone()
    K.running(d_fr[cut:], rect, d_om, g2, means, out[cut:], prev_crop=prev_crop, prev_mean=prev_mean)
    crop = fr[:, y0:y1, x0:x1] | om
    dy, dx = roi_oracle.running_chunk(crop)
    got = out.cpu().numpy()
    assert np.array_equal(got[0], [0, 0])
    assert np.array_equal(got[1:, 0], dy) and np.array_equal(got[1:, 1], dx)
    # the float32 means follow NumPy's rounding exactly
    X = crop.astype(np.float32)
    assert np.array_equal(means.cpu().numpy()[:T - cut], X.mean(axis=-1).mean(axis=-1)[cut:])


@pytest.mark.parametrize("case", [("eye", 90.27, 2.0, False), ("eye", 60.0, 2.5, True), ("noise", 128.5, 2.0, True),
                                  ("eye", 254.0, 4.0, False), ("big", 100.45, 2.0, True)])
def test_pupil_kernel(cuda_device, case):
    from facemap_b200 import fullres
    from facemap_b200 import kernels as K
    from tests.helpers import pupil_roi, reflector_roi

    kind, sat, sigma, with_reflector = case
    if kind == "eye":
        fr = eye_video(72, 96, 640, seed=5)[400:640]
        rect = (8, 44, 16, 64)
    elif kind == "big":          # > 128-element leaves and several tree levels in the float32 pairwise sum
        fr = eye_video(200, 260, 24, seed=6)
        rect = (5, 190, 10, 250)
    else:
        fr = np.random.RandomState(3).randint(0, 256, size=(40, 50, 60)).astype(np.uint8)
        rect = (4, 40, 6, 51)
    y0, y1, x0, x1 = rect
    refl = [reflector_roi(8, 16, 14, 25), reflector_roi(-3, 4, 30, 44)] if with_reflector else None
    r = pupil_roi(0, y0, y1, x0, x1, sat, sigma=sigma, reflector=refl)
    mask = (~r["ellipse"]).astype(np.uint8)
    T = fr.shape[0]
    out = torch.zeros((T, 9), dtype=torch.float64, device=cuda_device)
    blocks = 7                    # fewer blocks than frames: every block loops
    scratch = K.pupil_scratch(rect, blocks, cuda_device)
    rm = fullres.reflector_map(r) if with_reflector else None
    K.pupil(_dev(fr, cuda_device), rect, _dev(mask, cuda_device), np.float32(255.0 - sat), sigma, scratch, blocks, out,
            reflector=None if rm is None else _dev(rm.astype(np.uint8), cuda_device))
    crop = fr[:, y0:y1, x0:x1].copy()
    crop[:, mask > 0] = 255
    com, area, axdir, axlen = roi_oracle.pupil_chunk(crop, sat, sigma, roi_oracle.reflector_pixels(r) if with_reflector else None)
    _pupil_close(out.cpu().numpy(), com, area, axdir, axlen, where=str(case))


# ---------------------------------------------------------------------------------------------
# whole path
# ---------------------------------------------------------------------------------------------
def _run(name, device, **kw):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource

    views, sbin, mot, mov, rois, fullSVD = fullres_case(name)
    src = DeviceArraySource([torch.from_numpy(v).to(device) for v in views])
    return process.process_source(src, sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), fullSVD=fullSVD, **kw)


@pytest.mark.parametrize("name", FULLRES_CASES)
def test_fullres_matches_reference_golden(cuda_device, golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"fullres_{name}.npz"))
    p = _run(name, cuda_device)
    assert len(p["pupil"]) == int(z["n_pupil"]) and len(p["blink"]) == int(z["n_blink"])
    assert len(p["running"]) == int(z["n_running"])
    for k, pu in enumerate(p["pupil"]):
        got = np.concatenate((pu["com"], pu["area"][:, None], pu["axdir"].reshape(-1, 4), pu["axlen"]), axis=1)
        _pupil_close(got, z[f"pupil{k}_com"], z[f"pupil{k}_area"], z[f"pupil{k}_axdir"], z[f"pupil{k}_axlen"], f"{name} pupil {k}")
        # the smoothed traces pick samples of the raw / median traces: same tolerance
        for key in ("area_smooth", "com_smooth"):
            assert np.allclose(pu[key], z[f"pupil{k}_{key}"], rtol=1e-8, atol=1e-12, equal_nan=True), (name, k, key)
    for k, b in enumerate(p["blink"]):
        assert b.dtype == np.float64 and np.allclose(b, z[f"blink{k}"], rtol=BLINK_RTOL, atol=0), (name, k)
    for k, r in enumerate(p["running"]):
        assert r.dtype == np.float64 and np.array_equal(r, z[f"running{k}"]), (name, k)
    for k, m in enumerate(p["motion"]):          # R1: the painted frames feed stage 3
        assert np.array_equal(m, z[f"motion{k}"]), (name, "motion", k)
    for key in ("motSVD", "movSVD"):
        assert len(p[key]) == sum(1 for f in z.files if f.startswith(key))
        for k, v in enumerate(p[key]):
            ref = z[f"{key}{k}"]
            if ref.shape[0] == 0:
                assert v.shape == (0, 1)
                continue
            va, _ = sign_align(v[:, :4], ref[:, :4])
            err = np.linalg.norm(va - ref[:, :4], axis=0) / np.linalg.norm(ref[:, :4], axis=0)
            assert err.max() <= 1e-3, (name, key, k, err)


def test_fullres_independent_of_batching(cuda_device):
    """small fetches (many carries) and the side-stream overlap give the same bits as one big fetch"""
    from facemap_b200 import svd

    a = _run("eye_all", cuda_device)
    old = svd.Stage3.__init__.__defaults__
    svd.Stage3.__init__.__defaults__ = tuple(97 if d == 2368 else d for d in old)
    try:
        b = _run("eye_all", cuda_device, overlap_stage3=False)
    finally:
        svd.Stage3.__init__.__defaults__ = old
    for k in range(len(a["pupil"])):
        for key in ("com", "area", "axdir", "axlen"):
            assert np.array_equal(a["pupil"][k][key], b["pupil"][k][key], equal_nan=True)
    assert np.array_equal(a["blink"][0], b["blink"][0])
    assert np.array_equal(a["running"][0], b["running"][0])
    assert np.array_equal(a["motion"][0], b["motion"][0])


def test_fullres_host_source_and_file(cuda_device, tmp_path):
    """through run(): host arrays in, _proc.npy out with the pupil / blink / running entries filled"""
    from facemap_b200 import process

    views, sbin, mot, mov, rois, fullSVD = fullres_case("eye_only")
    cfg = {"sbin": sbin, "fullSVD": fullSVD, "save_mat": False, "rois": copy.deepcopy(rois), "sy": 0, "sx": 0,
           "savepath": str(tmp_path)}
    name = process.run(views, motSVD=mot, movSVD=mov, proc=cfg)
    p = np.load(name, allow_pickle=True).item()
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "fullres_eye_only.npz"))
    assert np.array_equal(p["running"][0], z["running0"])
    assert set(p["pupil"][0]) == {"area", "com", "axdir", "axlen", "area_smooth", "com_smooth"}
    assert np.allclose(p["pupil"][0]["area"], z["pupil0_area"], rtol=PUPIL_TOL, equal_nan=True)
    assert p["motion"][0].shape == (0,) and p["motSVD"][0].shape == (0, 1) and p["motMask"] == []


def test_fullres_frame_range_shards_concatenate(cuda_device):
    """what a rank of a multi-GPU run does: its own frame range with the one-frame halo seeding running's
    predecessor; the shards put together are the single-range result bit for bit"""
    from facemap_b200 import fullres
    from facemap_b200.frames import DeviceArraySource

    views, sbin, mot, mov, rois, fullSVD = fullres_case("eye_all")
    N = views[0].shape[0]
    src = DeviceArraySource([torch.from_numpy(v).to(cuda_device) for v in views])
    Ly, Lx = [v.shape[1] for v in views], [v.shape[2] for v in views]
    whole = fullres.FullRes(rois, Ly, Lx, N, cuda_device)
    whole.sweep(src, fetch_frames=300)
    parts = []
    for rng in ((0, 500), (500, 501), (501, N)):
        fr = fullres.FullRes(rois, Ly, Lx, N, cuda_device, frame_range=rng)
        fr.sweep(src, fetch_frames=128)
        parts.append(fr.device_outputs())
    for kind in range(3):
        for k, ref in enumerate(whole.device_outputs()[kind]):
            got = torch.cat([p[kind][k] for p in parts], dim=0)
            assert torch.equal(torch.nan_to_num(got.double(), nan=-1.0), torch.nan_to_num(ref.double(), nan=-1.0)), (kind, k)
