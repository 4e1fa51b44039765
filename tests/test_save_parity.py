"""`facemap_b200.process.save` against the reference's own `process.save` (facemap/process.py:605-635) on the same proc
dict: the pickled `.npy`, the `.mat` flattening rule (lists of arrays -> `key_i`, `rois=None -> 0`, `save_path` filled
in), the returned path and the in-place side effects.  Runs where /root/reference exists (the build container); the
GPU box has no reference and skips.  SURVEY §8(f) row 3 (proc writer)."""
import copy
import os

import numpy as np
import pytest

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference (build container)")


def _proc(folder, save_mat, rois):
    rs = np.random.RandomState(0)
    return {
        "filenames": [[os.path.join(folder, "cam0.avi"), os.path.join(folder, "cam1.avi")]],
        "save_path": None, "Ly": [24, 16], "Lx": [36, 20], "sbin": 2, "fullSVD": True, "save_mat": save_mat,
        "Lybin": np.array([12, 8]), "Lxbin": np.array([18, 10]), "sybin": np.array([0, 0]), "sxbin": np.array([0, 18]),
        "LYbin": 12, "LXbin": 28,
        "avgframe": [rs.rand(216).astype(np.float32), rs.rand(80).astype(np.float32)],
        "avgmotion": [rs.rand(216).astype(np.float32), rs.rand(80).astype(np.float32)],
        "avgframe_reshape": rs.rand(12, 28).astype(np.float32), "avgmotion_reshape": rs.rand(12, 28).astype(np.float32),
        "motion": [rs.rand(50).astype(np.float32), np.zeros(0, np.float32)],
        "motSv": rs.rand(7).astype(np.float32), "movSv": np.zeros(0, np.float32),
        "motMask": [rs.rand(296, 7).astype(np.float32)], "movMask": [],
        "motMask_reshape": [rs.rand(12, 28, 7).astype(np.float32)], "movMask_reshape": [],
        "motSVD": [rs.rand(50, 7).astype(np.float32)], "movSVD": [],
        "pupil": [], "running": [], "blink": [], "rois": rois, "sy": np.array([0, 0]), "sx": np.array([0, 36]),
    }


def _same(a, b, path=""):
    assert type(a) is type(b), (path, type(a), type(b))
    if isinstance(a, dict):
        assert list(a.keys()) == list(b.keys()), path
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif isinstance(a, np.ndarray):
        assert a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, b.dtype, a.shape, b.shape)
        if a.dtype.names:                                   # MATLAB struct arrays (the ROI dicts)
            for name in a.dtype.names:
                _same(a[name], b[name], f"{path}.{name}")
        elif a.dtype == object:
            for i, (x, y) in enumerate(zip(a.ravel(), b.ravel())):
                _same(x, y, f"{path}<{i}>")
        else:
            assert np.array_equal(a, b, equal_nan=a.dtype.kind in "fc"), path
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("save_mat", [False, True])
@pytest.mark.parametrize("with_rois", [False, True])
@pytest.mark.parametrize("use_savepath", [False, True])
def test_save_matches_reference(tmp_path, save_mat, with_rois, use_savepath):
    from scipy import io

    from facemap_b200 import process as mine

    ref_process, _ = ref_harness._import_reference()
    outs = []
    for tag, save in (("ref", ref_process.save), ("mine", mine.save)):
        src = tmp_path / tag / "videos"
        dst = tmp_path / tag / "out"
        src.mkdir(parents=True)
        dst.mkdir(parents=True)
        rois = [{"rind": 1, "ivid": 0, "yrange": np.arange(2, 9), "xrange": np.arange(3, 11), "saturation": 0}] if with_rois else None
        proc = _proc(str(src), save_mat, copy.deepcopy(rois))
        name = save(proc, savepath=str(dst) if use_savepath else None)
        folder = dst if use_savepath else src
        assert name == os.path.join(str(folder), "cam0_proc.npy") and os.path.exists(name)
        loaded = np.load(name, allow_pickle=True).item()
        mat = io.loadmat(os.path.join(str(folder), "cam0_proc.mat")) if save_mat else None
        assert os.path.exists(os.path.join(str(folder), "cam0_proc.mat")) == save_mat
        outs.append((proc, loaded, mat, str(tmp_path / tag)))
    (p_ref, l_ref, m_ref, root_ref), (p_me, l_me, m_me, root_me) = outs

    def unroot(d, root):     # the two runs write into sibling folders: compare paths relative to them
        d = dict(d)
        d["filenames"] = [[f.replace(root, "") for f in blk] for blk in d["filenames"]]
        if isinstance(d.get("save_path"), str):
            d["save_path"] = d["save_path"].replace(root, "")
        return d

    _same(unroot(l_ref, root_ref), unroot(l_me, root_me), "npy")
    _same(unroot(p_ref, root_ref), unroot(p_me, root_me), "in-place side effects")
    if save_mat:
        assert sorted(m_ref.keys()) == sorted(m_me.keys())
        for k in m_ref:
            if k.startswith("__") or k in ("filenames", "save_path"):
                continue
            _same(m_ref[k], m_me[k], f"mat/{k}")
        assert "motSVD_0" in m_me and "motMask_reshape_0" in m_me and "avgframe_1" in m_me
        if not with_rois:
            assert int(np.squeeze(m_me["rois"])) == 0
