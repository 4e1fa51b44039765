"""The merge eigensolve by block subspace iteration (svd.subspace_topk, the default from 3 000 stacked components on)
against the direct cuSOLVER solve on the same video: motion and movie SVD of 16 000 frames of 640x480 at sbin 4
(15 chunks, 3 750 stacked components per modality — facemap/process.py:264-295).  The comparison with the reference
itself at this size is tests/test_large_golden_gpu.py[c2_20k_s4]; scripts/probe_merge_modes.py is the same run as a
probe (profiles/r02_logs/probe_merge_modes.json)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_subspace_merge_equals_direct_merge(cuda_device, monkeypatch):
    from facemap_b200 import process, svd
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    frames = SynthVideo(480, 640, 16000, seed=3).all_frames_torch(cuda_device)
    taken = []
    original = svd.subspace_topk

    def counting(G, k, **kw):
        res = original(G, k, **kw)
        taken.append((int(G.shape[0]), int(k), res is not None))
        return res

    monkeypatch.setattr(svd, "subspace_topk", counting)
    outs = {}
    for mode in ("cusolver", "subspace"):
        monkeypatch.setenv("FM_MERGE_EIG", mode)
        outs[mode] = process.process_source(DeviceArraySource([frames]), sbin=4, motSVD=True, movSVD=True)
    # both modalities' merges (3 750 columns, 500 components) were given to the subspace solver, and the motion one
    # accepted (either may decline: the direct solve then takes over and the comparison below still has to hold)
    assert [t[:2] for t in taken] == [(3750, 500), (3750, 500)] and taken[0][2], taken
    ref, sub = outs["cusolver"], outs["subspace"]
    for key in ("motSv", "movSv"):
        a, b = ref[key].astype(np.float64), sub[key].astype(np.float64)
        assert a.shape == (500,) and np.max(np.abs(a - b) / a) <= 1e-5, (key, float(np.max(np.abs(a - b) / a)))
    for key in ("motMask", "movMask"):
        U, V = ref[key][0].astype(np.float64), sub[key][0].astype(np.float64)
        assert U.shape == V.shape == (19200, 500)
        c = np.abs(np.sum(U[:, :40] * V[:, :40], axis=0)) / (np.linalg.norm(U[:, :40], axis=0) * np.linalg.norm(V[:, :40], axis=0))
        assert np.sqrt(np.max(1 - np.minimum(c, 1.0) ** 2)) <= 1e-5, key          # the leading, well-separated masks
        assert np.abs(V.T @ V - np.eye(500)).max() <= 1e-3, key                    # float32 masks, orthonormal
    for key in ("motSVD", "movSVD"):
        a, b = ref[key][0][:, :40].astype(np.float64), sub[key][0][:, :40].astype(np.float64)
        assert np.max(np.abs(np.abs(a) - np.abs(b))) <= 1e-4 * np.max(np.abs(a)), key
    for key in ("avgframe", "avgmotion", "motion"):
        assert np.array_equal(ref[key][0], sub[key][0]), key


def test_small_merges_stay_on_the_direct_solver(cuda_device, monkeypatch):
    """below 3 000 stacked components (here 2 chunks x 250) subspace_topk is never asked"""
    from facemap_b200 import process, svd
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    asked = []
    monkeypatch.setattr(svd, "subspace_topk", lambda *a, **k: asked.append(1))
    frames = SynthVideo(64, 96, 2300, seed=0, nmodes=12).all_frames_torch(cuda_device)
    out = process.process_source(DeviceArraySource([frames]), sbin=2, motSVD=True, movSVD=False)
    assert not asked and out["motMask"][0].shape[1] > 0
