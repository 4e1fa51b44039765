"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in-process.

Build-container only (needs /root/reference).  For every case it runs the reference twice
(O1 = stock, O2 = `utils.svdecon` swapped for an exact float64 SVD), checks the
oracle restatement (`oracle/facemap_oracle.py`) against it, and stores a reduced set of outputs:
geometry, avgframe/avgmotion, motion, singular values, the leading KEEP mask components and the
leading KEEP projected traces on a window of frames.

    python scripts/make_golden.py            # writes tests/golden/<case>.npz
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facemap_b200.synth import SynthVideo  # noqa: E402
from oracle import facemap_oracle as orc  # noqa: E402
from oracle.ref_harness import run_reference  # noqa: E402

KEEP = 16


def motion_roi(ivid, y0, y1, x0, x1):
    return {"rind": 1, "rtype": "motion SVD", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
            "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": 0}


CASES = {
    # name: (views [(H, W, seed)], N, sbin, motSVD, movSVD, rois)
    "single_s2": ([(64, 96, 0)], 2300, 2, True, True, None),
    "multi_roi_s4": ([(96, 128, 0), (80, 112, 1)], 2200, 4, True, True,
                     [motion_roi(0, 10, 70, 17, 90), motion_roi(1, 8, 60, 20, 100)]),
    "onechunk_s1": ([(24, 32, 2)], 1300, 1, True, False, None),
    "crop_s3": ([(50, 70, 3)], 2100, 3, True, False, None),
    "short_s2": ([(40, 56, 4)], 700, 2, True, True, None),
    # fewer binned pixels (384) than chunk components (500) in the single-chunk regime: the reference leaves the
    # unfilled columns of its preallocated mask buffer zero (process.py:146-151, 246-259; never cut: :264-267)
    "tiny_onechunk_s2": ([(32, 48, 5)], 1200, 2, True, True, None),
}


# BASELINE configs[0] at full size and configs[1] at the 20 000-frame size SURVEY.md §8(d) names for parity runs (the
# reference needs ~12 min for the 100k-frame video and a prefix is not equivalent: chunk positions depend on N).
# Generated only when named on the command line; replayed by opt-in tests (tests/test_large_golden_gpu.py).
LARGE_CASES = {
    "c1_s4": ([(256, 256, 0)], 2000, 4, True, False, None),
    "c2_20k_s4": ([(480, 640, 0)], 20000, 4, True, False, None),
    # configs[2] at its 4 000-frame parity size: two cameras, motion + movie SVD, the three motion ROIs of SURVEY §8(d)
    "c3_4k_s4": ([(480, 640, 0), (480, 640, 1)], 4000, 4, True, True,
                 [motion_roi(0, 40, 200, 80, 320), motion_roi(1, 32, 160, 64, 300), motion_roi(0, 240, 440, 400, 600)]),
    # configs[3] on a 256-pixel-wide crop: full resolution (sbin 1, float32 path of the reference), 262 144 pixels
    "c4_crop_s1": ([(1024, 256, 0)], 2000, 1, True, False, None),
}
LARGE_KEEP = {"c3_4k_s4": 8, "c4_crop_s1": 2}          # leading components stored (fixture size)


def case_videos(name):
    views, N, *_ = {**CASES, **LARGE_CASES}[name]
    return [SynthVideo(H, W, N, seed=seed).all_frames() for (H, W, seed) in views]


def frame_window(N):
    """Frames kept for the trace fixtures: start (Q1/Q2 + first seam) and the tail."""
    idx = np.concatenate((np.arange(0, min(N, 520)), np.arange(max(0, N - 60), N)))
    return np.unique(idx)


def reduce_proc(p, N, prefix, out, keep=KEEP):
    w = frame_window(N)
    out[prefix + "frame_window"] = w
    for key in ("Lybin", "Lxbin", "sybin", "sxbin"):
        out[prefix + key] = np.asarray(p[key])
    out[prefix + "LYbin"] = np.int64(p["LYbin"])
    out[prefix + "LXbin"] = np.int64(p["LXbin"])
    for key in ("avgframe", "avgmotion"):
        for i, a in enumerate(p[key]):
            out[f"{prefix}{key}_{i}"] = np.asarray(a)
    out[prefix + "avgframe_reshape"] = p["avgframe_reshape"]
    out[prefix + "avgmotion_reshape"] = p["avgmotion_reshape"]
    for i, a in enumerate(p["motion"]):
        out[f"{prefix}motion_{i}"] = a
    out[prefix + "motSv"] = np.asarray(p["motSv"])
    out[prefix + "movSv"] = np.asarray(p["movSv"])
    for mk, vk in (("motMask", "motSVD"), ("movMask", "movSVD")):
        out[f"{prefix}{mk}_n"] = np.int64(len(p[mk]))
        for i, a in enumerate(p[mk]):
            out[f"{prefix}{mk}_{i}_shape"] = np.array(a.shape)
            out[f"{prefix}{mk}_{i}"] = a[:, :keep].copy()
        for i, a in enumerate(p[vk]):
            out[f"{prefix}{vk}_{i}_shape"] = np.array(a.shape)
            out[f"{prefix}{vk}_{i}"] = a[w][:, :keep].copy() if a.shape[0] == N else a
    if p["rois"] is not None:
        for i, r in enumerate(p["rois"]):
            if r["rind"] == 1:
                out[f"{prefix}roi{i}_yrange_bin"] = r["yrange_bin"]
                out[f"{prefix}roi{i}_xrange_bin"] = r["xrange_bin"]


def slim(out):
    """large fixtures: the integer-stage outputs are identical in both modes (checked above against the oracle) and the
    canvases repeat the vectors — keep the stock copy of the vectors only, and of the stock SVD only the spectrum"""
    import re

    drop = [k for k in out if k.endswith("_reshape") or (k.startswith("exact/") and k.split("/")[1].split("_")[0] in
                                                          ("avgframe", "avgmotion", "motion"))]
    # masks and traces of the stock run: only its singular values are compared at these sizes
    drop += [k for k in out if re.fullmatch(r"stock/(motMask|movMask|motSVD|movSVD)_\d+", k)]
    return {k: v for k, v in out.items() if k not in drop}


def same(a, b):
    if isinstance(a, list):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b)


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = sys.argv[1:]
    todo = dict(CASES) if not only else {k: v for k, v in {**CASES, **LARGE_CASES}.items() if k in only}
    for name, (views, N, sbin, mot, mov, rois) in todo.items():
        vids = case_videos(name)
        keep = LARGE_KEEP.get(name, KEEP)
        out = {"case": np.array(name), "N": np.int64(N), "sbin": np.int64(sbin),
               "motSVD": np.bool_(mot), "movSVD": np.bool_(mov), "keep": np.int64(keep),
               "views": np.array(views)}
        for mode, exact in (("stock", False), ("exact", True)):
            t = time.time()
            p = run_reference([vids], sbin=sbin, motSVD=mot, movSVD=mov, rois=rois, exact=exact)
            tr = time.time() - t
            t = time.time()
            o = orc.run_oracle(vids, sbin=sbin, motSVD=mot, movSVD=mov, rois=rois, svd=mode)
            to = time.time() - t
            bad = [k for k in ("Lybin", "Lxbin", "sybin", "sxbin", "avgframe", "avgmotion",
                               "avgframe_reshape", "avgmotion_reshape", "motion", "motSv", "movSv",
                               "motMask", "movMask", "motMask_reshape", "movMask_reshape",
                               "motSVD", "movSVD") if not same(p[k], o[k])]
            print(f"{name:14s} {mode:5s} reference {tr:6.1f}s oracle {to:6.1f}s  "
                  f"oracle==reference bit-for-bit: {'YES' if not bad else 'NO ' + str(bad)}")
            reduce_proc(p, N, mode + "/", out, keep)
        if name in LARGE_KEEP:
            out = slim(out)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print("  wrote", name + ".npz", os.path.getsize(os.path.join(outdir, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
