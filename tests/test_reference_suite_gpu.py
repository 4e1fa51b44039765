"""The reference's own SVD integration tests (facemap/tests/test_svd_output.py:14-66), restated for the
drop-in: the same two calls — `process.run([[cam1.avi]], sbin=7, motSVD=True, movSVD=True)` and
`process.run([[cam1.avi, cam2.avi]], sbin=12, ...)` — on lossless synthetic videos, and the same assertions
(`check_params` exact, `check_frames` allclose(rtol 1e-2, atol 1e-1), `check_motion` exact ==, `check_U`
per-component allclose up to sign + reshape shapes; `check_V` is disabled upstream and enforced here), with
the expected output produced by the reference's algorithm (oracle, stock randomized svdecon = what upstream's
golden files were made with).  The reference's goldens themselves are not reachable offline."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from facemap_b200.synth import SynthVideo  # noqa: E402
from oracle import facemap_oracle as orc  # noqa: E402
from tests.test_video_files import write_avi  # noqa: E402

pytestmark = pytest.mark.gpu

R_TOL, A_TOL = 1e-2, 1e-1          # test_svd_output.py:11


def params_match(out, exp):        # test_svd_output.py:69-81
    return (out["Ly"] == exp["Ly"] and out["Lx"] == exp["Lx"] and out["sbin"] == exp["sbin"]
            and out["Lybin"][0] == exp["Lybin"][0] and out["Lxbin"][0] == exp["Lxbin"][0]
            and out["sybin"][0] == exp["sybin"][0] and out["sxbin"][0] == exp["sxbin"][0]
            and out["LYbin"] == exp["LYbin"] and out["LXbin"] == exp["LXbin"])


def frames_match(out, exp):        # test_svd_output.py:84-114
    return all(np.allclose(out[k][0], exp[k][0], rtol=R_TOL, atol=A_TOL)
               for k in ("avgframe", "avgmotion", "avgframe_reshape", "avgmotion_reshape"))


def motion_match(out, exp):        # test_svd_output.py:155-156
    return bool((out["motion"][0] == exp["motion"][0]).all())


def masks_match(out, exp, n=150):  # test_svd_output.py:117-142
    """Upstream loops over all components; its tolerance only bites on frames as small as these (mask entries
    reach 0.2-0.5 on a few hundred pixels, atol is 0.1), and there the stock randomized solver has not converged
    on near-degenerate and noise-floor components — the reference run with an exact SVD fails the same check
    against the stock reference from component 175 (single video) / 238 (two cameras) on.  The first 150 are
    compared."""
    assert out["motSVD"][0].shape[1] == exp["motSVD"][0].shape[1]
    ok = True
    for key in ("motMask", "movMask"):
        a, b = out[key][0], exp[key][0]
        ok &= all(np.allclose(a[:, i], b[:, i], rtol=R_TOL, atol=A_TOL) or
                  np.allclose(a[:, i], -b[:, i], rtol=R_TOL, atol=A_TOL) for i in range(n))
        ok &= out[key + "_reshape"][0].shape == exp[key + "_reshape"][0].shape
    return bool(ok)


def traces_match(out, exp, k=20):
    """upstream's check_V (test_svd_output.py:145-152, disabled there) on the leading components, up to sign and
    relative to the trace scale — its absolute tolerance is meaningless for traces of magnitude 1e4..1e6"""
    ok = True
    for key in ("motSVD", "movSVD"):
        a, b = out[key][0][:, :k].astype(np.float64), exp[key][0][:, :k].astype(np.float64)
        sg = np.sign((a * b).sum(axis=0))
        ok &= bool(np.all(np.linalg.norm(a * sg - b, axis=0) <= R_TOL * np.linalg.norm(b, axis=0)))
    return ok


def run_and_load(filenames, sbin, savepath):
    from facemap_b200 import process

    name = process.run(filenames, sbin=sbin, motSVD=True, movSVD=True, savepath=savepath)
    stem = os.path.splitext(os.path.basename(filenames[0][0]))[0]
    assert name == os.path.join(savepath, stem + "_proc.npy")
    return np.load(name, allow_pickle=True).item()


def test_output_single_video(cuda_device, tmp_path):
    cam1 = SynthVideo(140, 196, 2300, seed=31).all_frames()      # even width: FFV1 drops an odd last column
    path = str(tmp_path / "cam1_test.avi")
    write_avi(path, cam1)
    out = run_and_load([[path]], 7, str(tmp_path))
    exp = orc.run_oracle([cam1], sbin=7, motSVD=True, movSVD=True, svd="stock")
    exp["sbin"] = 7
    assert params_match(out, exp) and frames_match(out, exp)
    assert motion_match(out, exp) and masks_match(out, exp) and traces_match(out, exp)


def test_output_multivideo(cuda_device, tmp_path):
    cam1 = SynthVideo(144, 192, 2100, seed=32).all_frames()
    cam2 = SynthVideo(120, 168, 2100, seed=33).all_frames()
    p1, p2 = str(tmp_path / "cam1_test.avi"), str(tmp_path / "cam2_test.avi")
    write_avi(p1, cam1)
    write_avi(p2, cam2)
    out = run_and_load([[p1, p2]], 12, str(tmp_path))
    exp = orc.run_oracle([cam1, cam2], sbin=12, motSVD=True, movSVD=True, svd="stock")
    exp["sbin"] = 12
    assert params_match(out, exp) and frames_match(out, exp)
    assert motion_match(out, exp) and masks_match(out, exp) and traces_match(out, exp)
