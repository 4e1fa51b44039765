"""Shared test helpers: synthetic cases (same definitions as scripts/make_golden.py) and the
comparison metrics of SURVEY.md §8(c)."""
import numpy as np

from facemap_b200.synth import SynthVideo

KEEP = 16


def motion_roi(ivid, y0, y1, x0, x1):
    return {"rind": 1, "rtype": "motion SVD", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
            "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": 0}


# name: (views [(H, W, seed)], N, sbin, motSVD, movSVD, rois) — must match scripts/make_golden.py
CASES = {
    "single_s2": ([(64, 96, 0)], 2300, 2, True, True, None),
    "multi_roi_s4": ([(96, 128, 0), (80, 112, 1)], 2200, 4, True, True,
                     [motion_roi(0, 10, 70, 17, 90), motion_roi(1, 8, 60, 20, 100)]),
    "onechunk_s1": ([(24, 32, 2)], 1300, 1, True, False, None),
    "crop_s3": ([(50, 70, 3)], 2100, 3, True, False, None),
    "short_s2": ([(40, 56, 4)], 700, 2, True, True, None),
}


def case_videos(name):
    views, N = CASES[name][0], CASES[name][1]
    return [SynthVideo(H, W, N, seed=seed).all_frames() for (H, W, seed) in views]


def case_rois(name):
    import copy

    return copy.deepcopy(CASES[name][5])


def frame_window(N):
    idx = np.concatenate((np.arange(0, min(N, 520)), np.arange(max(0, N - 60), N)))
    return np.unique(idx)


def ulp_distance(a, b):
    """max distance in float32 ulps between two float32 arrays."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return int(np.abs(a - b).max()) if a.size else 0


def subspace_sin(A, B):
    """sin of the largest principal angle between span(A) and span(B) (columns)."""
    Qa, _ = np.linalg.qr(np.asarray(A, np.float64))
    Qb, _ = np.linalg.qr(np.asarray(B, np.float64))
    s = np.linalg.svd(Qa.T @ Qb, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - min(1.0, s.min()) ** 2)))


def resolved_count(S, gap_rel=1e-3, floor_rel=1e-3):
    """Number of leading components that are spectrally resolved: singular value above
    floor_rel * S[0] and separated from both neighbours by a relative gap > gap_rel."""
    S = np.asarray(S, np.float64)
    k = 0
    for i in range(len(S) - 1):
        if S[i] < floor_rel * S[0]:
            break
        gap_next = (S[i] - S[i + 1]) / S[i]
        if gap_next < gap_rel:
            break
        k = i + 1
    return k


def sign_align(A, B):
    """flip columns of A to match the sign of B's columns (largest-|.| entry rule may pick a
    different pixel when two entries are nearly tied)."""
    sg = np.sign(np.sum(np.asarray(A, np.float64) * np.asarray(B, np.float64), axis=0))
    sg[sg == 0] = 1.0
    return A * sg[None, :], sg
