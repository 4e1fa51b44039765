"""Generate tests/golden/fullres_*.npz: the UNMODIFIED reference (/root/reference) run in-process on
the pupil / blink / running cases of tests/helpers.py, checked against oracle/roi_oracle.py.

Build-container only.   python scripts/make_golden_fullres.py [case ...]
"""
from __future__ import annotations

import copy
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import facemap_oracle as orc  # noqa: E402
from oracle.ref_harness import run_reference  # noqa: E402
from tests.helpers import FULLRES_CASES, fullres_case  # noqa: E402


def compare(a, b, name, report):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        report.append(f"{name}: shape {a.shape} vs {b.shape}")
        return
    same_nan = np.array_equal(np.isnan(a), np.isnan(b))
    exact = np.array_equal(a, b, equal_nan=True)
    with np.errstate(all="ignore"):
        err = np.nanmax(np.abs(a - b) / np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b)))) if a.size else 0.0
    report.append(f"{name}: {'bit-exact' if exact else f'max rel err {err:.3g}'}{'' if same_nan else '  NaN PATTERN DIFFERS'}")


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    only = sys.argv[1:]
    for name in FULLRES_CASES:
        if only and name not in only:
            continue
        views, sbin, mot, mov, rois, fullSVD = fullres_case(name)
        t = time.time()
        p = run_reference([views], sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), fullSVD=fullSVD, exact=True)
        tr = time.time() - t
        t = time.time()
        o = orc.run_oracle(views, sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), fullSVD=fullSVD, svd="exact")
        to = time.time() - t
        rep = []
        out = {"case": np.array(name)}
        for k, (pr, po) in enumerate(zip(p["pupil"], o["pupil"])):
            for key in ("com", "area", "axdir", "axlen", "area_smooth", "com_smooth"):
                compare(pr[key], po[key], f"pupil{k}.{key}", rep)
                out[f"pupil{k}_{key}"] = np.asarray(pr[key])
        for k, (br, bo) in enumerate(zip(p["blink"], o["blink"])):
            compare(br, bo, f"blink{k}", rep)
            out[f"blink{k}"] = np.asarray(br)
        for k, (rr, ro) in enumerate(zip(p["running"], o["running"])):
            compare(rr, ro, f"running{k}", rep)
            out[f"running{k}"] = np.asarray(rr)
        for k, (mr, mo) in enumerate(zip(p["motion"], o["motion"])):
            compare(mr, mo, f"motion{k}", rep)
            out[f"motion{k}"] = np.asarray(mr)
        for key in ("motSVD", "movSVD"):
            for k, (vr, vo) in enumerate(zip(p[key], o[key])):
                compare(vr[:, :8], vo[:, :8], f"{key}{k}[:, :8]", rep)
                out[f"{key}{k}"] = np.asarray(vr[:, :8])
        out["n_pupil"], out["n_blink"], out["n_running"] = len(p["pupil"]), len(p["blink"]), len(p["running"])
        print(f"{name}: reference {tr:.1f}s oracle {to:.1f}s")
        for line in rep:
            print("   ", line)
        np.savez_compressed(os.path.join(outdir, "fullres_" + name + ".npz"), **out)
        print("  wrote fullres_" + name + ".npz", os.path.getsize(os.path.join(outdir, "fullres_" + name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
