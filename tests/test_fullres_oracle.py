"""oracle/roi_oracle.py (pupil / blink / running + the paint quirk R1) replayed against the fixtures
the unmodified reference produced (scripts/make_golden_fullres.py).  Bit-for-bit: the oracle is NumPy
calling the same library routines in the same order as facemap/pupil.py, running.py, process.py."""
import copy
import os

import numpy as np
import pytest

from oracle import facemap_oracle as orc
from oracle import roi_oracle
from tests.helpers import FULLRES_CASES, fullres_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", FULLRES_CASES)
def test_fullres_oracle_matches_reference(name):
    z = np.load(os.path.join(GOLD, f"fullres_{name}.npz"))
    views, sbin, mot, mov, rois, fullSVD = fullres_case(name)
    o = orc.run_oracle(views, sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), fullSVD=fullSVD, svd="exact")
    assert len(o["pupil"]) == int(z["n_pupil"]) and len(o["blink"]) == int(z["n_blink"])
    assert len(o["running"]) == int(z["n_running"])
    for k, p in enumerate(o["pupil"]):
        for key in ("com", "area", "axdir", "axlen", "area_smooth", "com_smooth"):
            assert np.array_equal(p[key], z[f"pupil{k}_{key}"], equal_nan=True), (k, key)
    for k, b in enumerate(o["blink"]):
        assert np.array_equal(b, z[f"blink{k}"])
    for k, r in enumerate(o["running"]):
        assert np.array_equal(r, z[f"running{k}"])
    for k, m in enumerate(o["motion"]):           # painted frames feed the motion energy (R1)
        assert np.array_equal(m, z[f"motion{k}"])
    for key in ("motSVD", "movSVD"):
        for k, v in enumerate(o[key]):
            assert np.array_equal(v[:, :8], z[f"{key}{k}"])


def test_paint_changes_stage3_only():
    """R1: painting reaches stage 3's motion energy but not the stage-1 averages."""
    views, sbin, mot, mov, rois, fullSVD = fullres_case("eye_only")
    a = orc.run_oracle(views, sbin=sbin, rois=copy.deepcopy(rois), fullSVD=True, svd="gram")
    b = orc.run_oracle(views, sbin=sbin, rois=None, fullSVD=True, svd="gram")
    assert np.array_equal(a["avgframe"][0], b["avgframe"][0])
    assert np.array_equal(a["motMask"][0], b["motMask"][0])
    assert not np.array_equal(a["motion"][0], b["motion"][0])


def test_running_is_pixelwise_sign_agreement():
    """R2: with no FFT executed, (dy, dx) is the first argmax of sign agreement weighted by fhg^2."""
    rs = np.random.RandomState(0)
    im = rs.randint(0, 256, size=(6, 21, 30)).astype(np.uint8)
    dy, dx = roi_oracle.running_chunk(im)
    g2 = roi_oracle.running_filter(21, 30).astype(np.float32) ** 2
    l = min(int(21 * 0.4), int(30 * 0.4))
    X = im.astype(np.float32)
    X -= X.mean(axis=-1).mean(axis=-1)[:, None, None]
    s = np.sign(X)
    for t in range(5):
        cc = s[t + 1] * s[t] * g2
        ys = np.r_[21 - l:21, 0:l + 1]
        xs = np.r_[30 - l:30, 0:l + 1]
        k = cc[np.ix_(ys, xs)].argmax()
        assert (dy[t], dx[t]) == (k // (2 * l + 1) - l, k % (2 * l + 1) - l)


def test_blink_saturation_types():
    """R3: Python float -> clamped float64 sum; Python int -> uint8 wrap-around."""
    crop = np.array([[[10, 200], [255, 0]]], np.uint8)
    assert roi_oracle.blink_chunk(crop, 100.5)[0] == (90.5 + 0 + 0 + 100.5)
    assert roi_oracle.blink_chunk(crop, 100)[0] == float((90 + (100 - 200) % 256 + (100 - 255) % 256 + 100))


def test_smooth_trace_fills_nan_and_outliers():
    x = np.sin(np.arange(200) / 10.0) + 5
    x[50] = np.nan
    x[120] = 50.0
    y = roi_oracle.smooth_trace(x)
    assert not np.isnan(y).any() and abs(y[120] - 5) < 2 and abs(y[50] - x[49]) < 0.5
