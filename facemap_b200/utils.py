"""Host-side geometry of the motion/movie-SVD path (no array arithmetic on frames happens here).

Mirrors, by name and meaning, the helpers the reference keeps in facemap/process.py and
facemap/utils.py so that callers of the reference find the same vocabulary:

  binned_inds          facemap/process.py:16-26
  video_placement      facemap/utils.py:703-748
  multivideo_reshape   facemap/utils.py:622-630
  roi bin ranges       facemap/process.py:716-723
"""
from __future__ import annotations

import numpy as np


def binned_inds(Ly, Lx, sbin):
    """Binned sizes per view and each view's index range on the concatenated pixel axis."""
    Lyb = np.array([int(h) // int(sbin) for h in Ly], dtype=np.int32)
    Lxb = np.array([int(w) // int(sbin) for w in Lx], dtype=np.int32)
    start, ir = 0, []
    for h, w in zip(Lyb, Lxb):
        n = int(h) * int(w)
        ir.append(np.arange(start, start + n, 1, int))
        start += n
    return Lyb, Lxb, ir


def video_placement(Ly, Lx):
    """Where each view sits on the shared display canvas: views are taken largest-area first and
    stacked into columns of `gridy` views; a column is as wide as its widest view.
    Returns (LY, LX, sy, sx)."""
    Ly = np.asarray(Ly)
    Lx = np.asarray(Lx)
    nv = Ly.size
    gridy = 1 if nv in (2, 3) else int(np.round(nv ** 0.5 * 0.75))
    order = []
    area = (Ly * Lx).astype(np.int64)
    left = np.ones(nv, bool)
    while left.any():
        a = np.where(left, area, 0)
        i = int(np.argmax(a))          # first maximum, like the reference (ties / zero-area views)
        order.append(i)
        left[i] = False
    sy = np.zeros(Ly.shape, int)
    sx = np.zeros(Lx.shape, int)
    LY = 0
    x_off = 0
    for c0 in range(0, nv, max(gridy, 1)):
        column = order[c0:c0 + max(gridy, 1)]
        y_off = 0
        for i in column:
            sy[i], sx[i] = y_off, x_off
            y_off += int(Ly[i])
        LY = max(LY, y_off)
        x_off += max(int(Lx[i]) for i in column)
    return LY, x_off, sy, sx


def multivideo_reshape(X, LY, LX, sy, sx, Ly, Lx, iinds):
    """(sum npix, n) rows -> (LY, LX, n) float32 canvas with every view at its placement."""
    X = np.asarray(X)
    out = np.zeros((LY, LX, X.shape[-1]), np.float32)
    for i in range(len(Ly)):
        out[sy[i]:sy[i] + Ly[i], sx[i]:sx[i] + Lx[i]] = X[iinds[i]].reshape(Ly[i], Lx[i], -1)
    return out


def roi_bin_ranges(roi, sbin):
    """Binned index ranges of a motion ROI.  The x stop is divided AFTER the floor in the
    reference (process.py:721-723), so a fractional stop rounds up; reproduced on purpose."""
    y_lo, y_hi = np.floor(roi["yrange"][0] / sbin), np.floor(roi["yrange"][-1] / sbin)
    x_lo, x_hi = np.floor(roi["xrange"][0] / sbin), np.floor(roi["xrange"][-1]) / sbin
    return np.arange(y_lo, y_hi).astype(int), np.arange(x_lo, x_hi).astype(int)
