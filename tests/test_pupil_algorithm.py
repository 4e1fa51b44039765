"""The algorithm of `pupil_kernel` (facemap_b200/csrc/roi_procs.cu), restated step by step in NumPy, against the
oracle's line-by-line restatement of pupil.py — a CPU pin of the formulas the kernel uses where they differ in
form from the reference:

  * compaction order (dark non-reflector pixels row-major, then the reflector pixels), weights recomputed from the
    pixel value, a zero-weight flag instead of a stored weight;
  * first and second moments from ONE pass about the ROI centre (the reference makes two passes about the mean);
  * the reflector median from a histogram of pixel values (zeros first, then pixel value 255 .. 0);
  * NaN bookkeeping: a frame the reference abandons (numpy.linalg.eig raising) comes out all-NaN.

The pieces that decide bits (pairwise float32 sum, 2x2 inverse / eigen conventions, median) are pinned separately
in tests/test_numpy_sum.py on the very code the kernel compiles."""
import math
import warnings

import numpy as np
import pytest

from facemap_b200 import fullres
from oracle import roi_oracle
from tests.helpers import fullres_case

f32 = np.float32


def _moments_one_pass(val, ys, xs, n, S, cy, cx):
    with np.errstate(all="ignore"):
        lam = (val[:n] / f32(S)).astype(f32)
        s2 = np.sqrt(lam).astype(np.float64) ** 2
        y, x = ys[:n], xs[:n]
        yc, xc = y - cy, x - cx
        mu0, mu1 = float((lam.astype(np.float64) * y).sum()), float((lam.astype(np.float64) * x).sum())
        q0, qy, qx = float(s2.sum()), float((s2 * yc).sum()), float((s2 * xc).sum())
        qyy, qxy, qxx = float((s2 * yc * yc).sum()), float((s2 * yc * xc).sum()), float((s2 * xc * xc).sum())
        e0, e1 = mu0 - cy, mu1 - cx
        return (mu0, mu1, qyy - 2 * e0 * qy + e0 * e0 * q0 + 0.01, qxy - e0 * qx - e1 * qy + e0 * e1 * q0,
                qxx - 2 * e1 * qx + e1 * e1 * q0 + 0.01)


def kernel_algorithm(crop, mask, miss, c32, sigma):
    ny, nx = crop.shape
    pix = np.where(mask > 0, 255, crop).astype(np.int64)

    def darkness(p):
        return np.maximum(f32(0), (255 - p).astype(f32) - c32).astype(f32)

    missing = miss.astype(bool) if miss is not None else np.zeros_like(mask, bool)
    take = (darkness(pix) > 0) & ~missing
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    ys = np.concatenate((yy[take], yy[missing])).astype(np.float64)
    xs = np.concatenate((xx[take], xx[missing])).astype(np.float64)
    px = np.concatenate((pix[take], pix[missing]))
    n0, nm = int(take.sum()), int(missing.sum()) if miss is not None else 0
    dead = np.zeros(n0 + nm, bool)
    fill = f32(0)

    def weights():
        w = np.where(np.arange(n0 + nm) < n0, darkness(px), fill).astype(f32)
        w[dead] = 0
        return w

    thr2, thr1, thrm = 2.0 * sigma * sigma, sigma * sigma, 1.15 * sigma * sigma
    cy, cx = 0.5 * ny, 0.5 * nx
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w = weights()
        m = _moments_one_pass(w, ys, xs, n0, w[:n0].sum() if n0 else f32(0), cy, cx)
        ntot = n0 + nm
        for _ in range(5):
            inv = np.linalg.inv(np.array([[m[2], m[3]], [m[3], m[4]]]))
            dy, dx = ys[:n0] - m[0], xs[:n0] - m[1]
            dd = (dy * inv[0, 0] + dx * inv[1, 0]) * dy + (dy * inv[0, 1] + dx * inv[1, 1]) * dx
            dead[:n0] |= ~(dd <= thr2)
            if miss is not None:
                hist = np.zeros(257, np.int64)
                sel = dd <= thr1
                np.add.at(hist, np.where(dead[:n0][sel], 256, px[:n0][sel]), 1)
                vals = [f32(0)] * int(hist[256])
                for h in range(255, -1, -1):
                    vals += [darkness(np.array(h))] * int(hist[h])
                cnt = len(vals)
                med = f32(np.nan) if cnt == 0 else (vals[cnt // 2] if cnt % 2 else f32(f32(vals[cnt // 2 - 1] + vals[cnt // 2]) / f32(2)))
                fill = f32(med * f32(1.1))
                dy, dx = ys[n0:] - m[0], xs[n0:] - m[1]
                ddm = (dy * inv[0, 0] + dx * inv[1, 0]) * dy + (dy * inv[0, 1] + dx * inv[1, 1]) * dx
                dead[n0:] = ~(ddm <= thrm)
            w = weights()
            m = _moments_one_pass(w, ys, xs, ntot, w[:ntot].sum() if ntot else f32(0), cy, cx)
    a, b, c = sigma * sigma * m[2], sigma * sigma * m[3], sigma * sigma * m[4]
    if not all(map(math.isfinite, (a, b, c, m[0], m[1]))):
        return np.full(9, np.nan)
    ev, u = np.linalg.eig(np.array([[a, b], [b, c]]))
    ev = ev.min() * np.minimum(4, ev / ev.min())
    return np.array([m[0], m[1], math.pi * math.sqrt(ev[1] * ev[0]), u[0, 1], u[0, 0], u[1, 1], u[1, 0], ev[1], ev[0]])


@pytest.mark.parametrize("name", ["eye_all", "eye_reflector", "eye_only"])
def test_kernel_algorithm_matches_reference_restatement(name):
    views, sbin, mot, mov, rois, full = fullres_case(name)
    Ly, Lx = [v.shape[1] for v in views], [v.shape[2] for v in views]
    painters, _, _ = fullres.paint_plan(rois, Ly, Lx)
    worst = 0.0
    for pl in painters:
        if pl["rind"] != fullres.PUPIL:
            continue
        r = rois[pl["index"]]
        y0, y1, x0, x1 = pl["rect"]
        c32, sigma = f32(255.0 - r["saturation"]), float(r["pupil_sigma"])
        refl = fullres.reflector_map(r) if "reflector" in r else None
        missing = roi_oracle.reflector_pixels(r) if "reflector" in r else None
        for t in list(range(0, 12)) + list(range(503, 510)):           # ordinary frames + the degenerate ones
            crop = views[pl["view"]][t, y0:y1, x0:x1]
            painted = np.where(pl["mask"] > 0, 255, crop).astype(np.uint8)
            com, area, axdir, axlen = roi_oracle.pupil_chunk(painted[None], r["saturation"], sigma, missing)
            ref = np.concatenate((com[0], [area[0]], axdir[0].ravel(), axlen[0]))
            got = kernel_algorithm(crop, pl["mask"], refl, c32, sigma)
            assert np.array_equal(np.isnan(ref), np.isnan(got)), (name, pl["index"], t)
            if np.isnan(ref).any():
                continue
            rel = np.abs(ref - got) / np.maximum(np.abs(ref), 1e-300)
            worst = max(worst, rel[[0, 1, 2, 7, 8]].max(), np.abs(ref[3:7] - got[3:7]).max())
    assert worst <= 1e-9, worst
