"""Host geometry helpers against the reference's own functions on randomised inputs (1-6 views of mixed sizes, every
sbin from 1 to 12): `binned_inds` (facemap/process.py:16-26), `video_placement` (facemap/utils.py:703-748),
`multivideo_reshape` (facemap/utils.py:622-630) and the binned ROI ranges of `run` (facemap/process.py:713-724, with
its divide-after-floor quirk).  Runs where /root/reference exists (the build container)."""
import numpy as np
import pytest

from facemap_b200 import utils as mine
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference (build container)")


def _cases(n=150):
    rs = np.random.RandomState(0)
    for _ in range(n):
        nv = int(rs.randint(1, 7))
        Ly = [int(rs.randint(12, 200)) for _ in range(nv)]
        Lx = [int(rs.randint(12, 260)) for _ in range(nv)]
        if rs.rand() < 0.3:                                   # equal areas exercise argmax's tie rule
            Ly[-1], Lx[-1] = Ly[0], Lx[0]
        yield Ly, Lx, int(rs.randint(1, 13)), rs


def test_binned_inds_placement_and_reshape_match_reference():
    ref_process, ref_utils = ref_harness._import_reference()
    for Ly, Lx, sbin, rs in _cases():
        Lyb_r, Lxb_r, ir_r = ref_process.binned_inds(Ly, Lx, sbin)
        Lyb, Lxb, ir = mine.binned_inds(Ly, Lx, sbin)
        assert np.array_equal(Lyb, Lyb_r) and np.array_equal(Lxb, Lxb_r) and len(ir) == len(ir_r)
        assert all(np.array_equal(a, b) for a, b in zip(ir, ir_r))
        want = ref_utils.video_placement(np.array(Lyb_r), np.array(Lxb_r))
        got = mine.video_placement(Lyb, Lxb)
        assert int(got[0]) == int(want[0]) and int(got[1]) == int(want[1]), (Ly, Lx, sbin)
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]), (Ly, Lx, sbin)
        X = rs.rand(int(np.sum(Lyb_r * Lxb_r)), 3).astype(np.float32)
        A = ref_utils.multivideo_reshape(X, want[0], want[1], want[2], want[3], Lyb_r, Lxb_r, ir_r)
        B = mine.multivideo_reshape(X, got[0], got[1], got[2], got[3], Lyb, Lxb, ir)
        assert A.dtype == B.dtype and np.array_equal(A, B)


def test_roi_bin_ranges_match_the_expression_in_run():
    """the reference computes these inline in run (process.py:717-723); restated here character for character"""
    rs = np.random.RandomState(1)
    for _ in range(300):
        sbin = int(rs.randint(1, 13))
        y0, x0 = int(rs.randint(0, 100)), int(rs.randint(0, 100))
        r = {"yrange": np.arange(y0, y0 + int(rs.randint(2, 150))), "xrange": np.arange(x0, x0 + int(rs.randint(2, 150)))}
        yb = np.arange(np.floor(r["yrange"][0] / sbin), np.floor(r["yrange"][-1] / sbin)).astype(int)
        xb = np.arange(np.floor(r["xrange"][0] / sbin), np.floor(r["xrange"][-1]) / sbin).astype(int)
        got = mine.roi_bin_ranges(r, sbin)
        assert np.array_equal(got[0], yb) and np.array_equal(got[1], xb), (sbin, y0, x0)
