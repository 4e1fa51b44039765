"""`facemap_b200.process.run`'s host logic on the CPU with the GPU stages stubbed out: argument precedence
(parent > proc > keyword arguments, facemap/process.py:676-708), the default settings, what reaches the three stages,
the proc dict's keys in the reference's order (facemap/process.py:859-892) and the save location.  The stages
themselves are covered by the GPU tests; a stub stands in for `process_source` here."""
import os
import re

import numpy as np
import pytest

from facemap_b200 import process

REF_KEYS = ["filenames", "save_path", "Ly", "Lx", "sbin", "fullSVD", "save_mat", "Lybin", "Lxbin", "sybin", "sxbin",
            "LYbin", "LXbin", "avgframe", "avgmotion", "avgframe_reshape", "avgmotion_reshape", "motion", "motSv", "movSv",
            "motMask", "movMask", "motMask_reshape", "movMask_reshape", "motSVD", "movSVD", "pupil", "running", "blink",
            "rois", "sy", "sx"]


class _Source:
    closed = False

    def close(self):
        self.closed = True


@pytest.fixture
def stub(monkeypatch):
    calls = {}

    def fake_process_source(source, **kw):
        calls["kw"] = kw
        z = np.zeros(3, np.float32)
        return {"Ly": [8], "Lx": [8], "Lybin": np.array([4]), "Lxbin": np.array([4]), "sybin": np.array([0]),
                "sxbin": np.array([0]), "LYbin": 4, "LXbin": 4, "avgframe": [z], "avgmotion": [z],
                "avgframe_reshape": z, "avgmotion_reshape": z, "motion": [z], "motSv": z, "movSv": z, "motMask": [z],
                "movMask": [], "motMask_reshape": [z], "movMask_reshape": [], "motSVD": [z], "movSVD": [], "pupil": [],
                "running": [], "blink": []}

    src = _Source()
    monkeypatch.setattr(process, "process_source", fake_process_source)
    monkeypatch.setattr(process, "as_source", lambda filenames: src)
    calls["source"] = src
    return calls


def test_defaults_and_key_order(stub, tmp_path):
    files = [[str(tmp_path / "a.avi")]]
    name = process.run(files, sbin=3, motSVD=True, movSVD=True)
    assert name == str(tmp_path / "a_proc.npy") and stub["source"].closed
    early = stub["kw"].pop("stream_out")              # the writer that streams the file during the pass; never begun here
    assert isinstance(early, process._EarlyFile) and early.final_path == name
    assert stub["kw"] == {"sbin": 3, "motSVD": True, "movSVD": True, "rois": None, "fullSVD": True}
    assert process.LAST_RUN["written_early"] is False and sorted(os.listdir(tmp_path)) == ["a_proc.npy"]
    proc = np.load(name, allow_pickle=True).item()
    assert list(proc.keys()) == REF_KEYS
    assert proc["fullSVD"] is True and proc["save_mat"] is False and proc["rois"] is None
    assert proc["sy"] == 0 and proc["sx"] == 0 and proc["save_path"] is None and proc["filenames"] == files
    assert not os.path.exists(str(tmp_path / "a_proc.mat"))


def test_proc_settings_override_keyword_arguments(stub, tmp_path):
    out = tmp_path / "out"
    out2 = tmp_path / "out2"
    out.mkdir()
    out2.mkdir()
    rois = [{"rind": 1, "ivid": 0, "yrange": np.arange(0, 6), "xrange": np.arange(0, 6), "saturation": 0}]
    settings = {"sbin": 2, "fullSVD": False, "save_mat": True, "rois": rois, "sy": np.array([0]), "sx": np.array([0]),
                "savepath": str(out)}
    name = process.run([[str(tmp_path / "v.avi")]], sbin=7, motSVD=True, movSVD=False, proc=settings)
    assert name == str(out / "v_proc.npy")                   # proc["savepath"] when no savepath is passed
    assert stub["kw"]["sbin"] == 2 and stub["kw"]["fullSVD"] is False and stub["kw"]["rois"] is rois
    proc = np.load(name, allow_pickle=True).item()
    assert proc["sbin"] == 2 and proc["save_mat"] is True and proc["save_path"] == str(out)
    assert os.path.exists(str(out / "v_proc.mat"))
    name2 = process.run([[str(tmp_path / "v.avi")]], proc=settings, savepath=str(out2))
    assert name2 == str(out2 / "v_proc.npy")                 # an explicit savepath wins over proc["savepath"]


def test_key_order_is_the_references(stub):
    """REF_KEYS restates facemap/process.py:859-892; where the reference is present, read it from there"""
    path = "/root/reference/facemap/process.py"
    if not os.path.exists(path):
        pytest.skip("needs /root/reference (build container)")
    text = open(path).read()
    block = text[text.index("    proc = {\n", text.index("# Save output")):]
    block = block[:block.index("    }\n")]
    assert re.findall(r'^\s+"(\w+)":', block, flags=re.M) == REF_KEYS
