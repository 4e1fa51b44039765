"""Full-resolution ROI processors on the stage-3 frame sweep: pupil (rind 0), blink (rind 2), running
(rind 3) — facemap/process.py:354-385, 410-425, 533-601, pupil.py, running.py.

One `FullRes` object holds the device-side ROI tables and the output arrays; stage 3 hands it every fetched
batch of raw frames (`chunk`) before binning, then bins the painted copy (`painted`).  Behaviour that is
reproduced on purpose (oracle/roi_oracle.py R1-R4 has the long form):

  R1  pupil and blink ROIs paint the part of their rectangle outside their ellipse with 255 in the chunk
      buffer (process.py:543, 567).  Processing order is pupils, blinks, running, so ROI j reads the paint of
      ROIs <= j on its view, running and stage 3's binning read all of it.  Stages 1 and 2 see none.
  R2  running.py never applies its FFTs (results of fft2/ifft2 are dropped): (dy, dx) is the first argmax of
      the pixel-wise sign agreement of consecutive mean-subtracted frames, weighted by the squared
      Fourier-domain Gaussian, over the corner window.  Row 0 stays zero.
  R3  blink evaluates `maximum(0, 255 - v - (255 - saturation))` in whatever dtype NumPy promotes to
      (float64 clamp for a Python float, uint8 wrap-around for a Python int); tabulated here for v = 0..255
      by literally evaluating that expression.
  R4  pupil weights are float32 (NumPy pairwise sum, float32 division and square root), moments float64.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from ._lib import FacemapB200Error

PUPIL, BLINK, RUNNING = 0, 2, 3
NOT_PAINTED = 255
PUPIL_BLOCKS = 148 * 4          # persistent blocks of the pupil fit (one frame per block at a time)
PUPIL_SCRATCH_MAX = 1 << 30     # bytes of per-warp compaction scratch; fewer blocks for very large ROIs


def pupil_blocks(rect):
    y0, y1, x0, x1 = rect
    per_block = K.pupil_scratch_bytes(rect, 1)
    return int(max(8, min(PUPIL_BLOCKS, PUPIL_SCRATCH_MAX // per_block)))


def roi_rect(r):
    """[y0, y1) x [x0, x1): process.py:538-542 slices yrange[0] .. yrange[-1] inclusive."""
    return int(r["yrange"][0]), int(r["yrange"][-1]) + 1, int(r["xrange"][0]), int(r["xrange"][-1]) + 1


def gaussian_fourier_sq(ny, nx):
    """float32 [ny, nx]: square (in float32) of the float32 cast of running.py:29-42's Fourier-domain
    Gaussian (std 2 px) — the only factor of the reference's `phase_norm` output that is not a sign."""
    ax = np.abs(np.arange(nx) - np.arange(nx).mean())
    ay = np.abs(np.arange(ny) - np.arange(ny).mean())
    xx, yy = np.meshgrid(ax, ay)
    h = np.exp(-np.square(yy / 2) / 2) * np.exp(-np.square(xx / 2) / 2)
    h /= h.sum()
    g = np.real(np.fft.fft2(np.fft.ifftshift(h))).astype(np.float32)
    return g * g


def blink_terms(saturation):
    """float64 [256]: the reference's blink expression evaluated for every pixel value (R3)."""
    v = np.arange(256).astype(np.uint8)
    return np.asarray(np.maximum(0, 255 - v - (255 - saturation))).astype(np.float64)


def reflector_map(r):
    """bool [ny, nx] of reflector pixels inside a pupil crop (utils.py:666-700, dict branch)."""
    ny, nx = r["yrange"].size, r["xrange"].size
    m = np.zeros((ny, nx), bool)
    for q in r.get("reflector") or []:
        ry, rx = np.asarray(q["yrange"]), np.asarray(q["xrange"])
        ky, kx = (ry >= 0) & (ry < ny), (rx >= 0) & (rx < nx)
        m[np.ix_(ry[ky], rx[kx])] |= np.asarray(q["ellipse"], bool)[ky][:, kx]
    return m


def smooth_trace(x, win=30):
    """pupil.py:135-166: running nan-median over `win` zero-padded shifts, gaps interpolated from it,
    samples further than std/2 from it replaced.  Host post-processing of one trace (returns a copy)."""
    x = np.array(x, dtype=np.float64)
    nt, half = x.size, win // 2
    stack = np.zeros((win, nt))
    for row, k in enumerate(range(-half, half)):
        if k < 0:
            stack[row, :k] = x[-k:]
        elif k > 0:
            stack[row, k:] = x[:-k]
        else:
            stack[row] = x
    stack.sort(axis=0)                                    # NaNs go last
    cnt = win - np.isnan(stack).sum(axis=0)
    cols = np.arange(nt)
    with np.errstate(invalid="ignore"):
        med = np.where(cnt > 0, (stack[np.maximum(cnt - 1, 0) // 2, cols] + stack[cnt // 2 % win, cols]) / 2, np.nan)
    bad = (np.isnan(x) | np.isnan(med)).nonzero()[0]
    good = (~np.isnan(x) & ~np.isnan(med)).nonzero()[0]
    x[bad] = np.interp(bad, good, med[good]) if good.size > 0 else 0
    med[bad] = x[bad]
    swap = np.abs(x - med) > x.std() / 2
    x[swap] = med[swap]
    return x


def paint_plan(rois, Ly, Lx):
    """Host-side ROI tables (NumPy only).  Returns (painters, runners, ormask):
      painters  pupil then blink ROIs in processing order: dict(index, rind, view, rect, mask) where mask
                (uint8 [ny, nx]) marks the crop pixels that read as 255 for THIS roi (R1: its own paint and
                that of the ROIs processed before it on the same view)
      runners   running ROIs: dict(index, view, rect, ormask) with ormask 0x00 / 0xFF for all paint
      ormask    per view uint8 [H, W] 0x00 / 0xFF of all paint (what stage 3 bins), None for a clean view"""
    rois = rois or []
    order = ([i for i, r in enumerate(rois) if r["rind"] == PUPIL]
             + [i for i, r in enumerate(rois) if r["rind"] == BLINK])
    if len(order) >= NOT_PAINTED:
        raise FacemapB200Error("too many pupil / blink ROIs")
    paint = [np.full((int(h), int(w)), NOT_PAINTED, np.uint8) for h, w in zip(Ly, Lx)]

    def geometry(i, r):
        v = int(r["ivid"])
        y0, y1, x0, x1 = roi_rect(r)
        if not (0 <= v < len(Ly)) or y0 < 0 or x0 < 0 or y1 > Ly[v] or x1 > Lx[v] or y1 <= y0 or x1 <= x0:
            raise FacemapB200Error(f"ROI {i}: rectangle [{y0},{y1})x[{x0},{x1}) does not fit view {v}")
        return v, (y0, y1, x0, x1)

    painters, runners = [], []
    for j, i in enumerate(order):
        r = rois[i]
        v, rect = geometry(i, r)
        y0, y1, x0, x1 = rect
        ell = np.asarray(r["ellipse"], bool)
        if ell.shape != (y1 - y0, x1 - x0):
            raise FacemapB200Error(f"ROI {i}: ellipse {ell.shape} does not match its rectangle {(y1 - y0, x1 - x0)}")
        sub = paint[v][y0:y1, x0:x1]
        sub[(~ell) & (sub == NOT_PAINTED)] = j                                       # first painter wins
        painters.append({"index": i, "rind": int(r["rind"]), "view": v, "rect": rect,
                         "mask": np.ascontiguousarray((sub <= j).astype(np.uint8))})
    for i, r in enumerate(rois):
        if r["rind"] != RUNNING:
            continue
        v, rect = geometry(i, r)
        y0, y1, x0, x1 = rect
        if min(int(np.floor((y1 - y0) * 0.4)), int(np.floor((x1 - x0) * 0.4))) < 1:
            raise FacemapB200Error(f"running ROI {i} ({y1 - y0}x{x1 - x0}) is too small")
        runners.append({"index": i, "view": v, "rect": rect,
                        "ormask": np.ascontiguousarray(np.where(paint[v][y0:y1, x0:x1] != NOT_PAINTED, 255, 0).astype(np.uint8))})
    ormask = [None if (p == NOT_PAINTED).all() else np.where(p != NOT_PAINTED, 255, 0).astype(np.uint8) for p in paint]
    return painters, runners, ormask


class FullRes:
    """rois: the reference's list of ROI dicts (all kinds; motion ROIs are ignored here).
    Frames [f0, f1) of an N-frame video are processed; outputs are indexed by frame - f0."""

    def __init__(self, rois, Ly, Lx, N, device, frame_range=None):
        self.device, self.N = device, int(N)
        self.f0, self.f1 = (0, self.N) if frame_range is None else frame_range
        n_out = self.f1 - self.f0
        painters, runners, ormask = paint_plan(rois, Ly, Lx)
        self.pupils, self.blinks, self.runs = [], [], []
        for pl in painters:
            r = rois[pl["index"]]
            entry = {"view": pl["view"], "rect": pl["rect"], "mask": torch.from_numpy(pl["mask"]).to(device)}
            if pl["rind"] == PUPIL:
                entry["dark_offset"] = np.float32(255.0 - r["saturation"])
                entry["sigma"] = float(r["pupil_sigma"])
                refl = reflector_map(r) if "reflector" in r else None
                entry["reflector"] = None if refl is None or not refl.any() else \
                    torch.from_numpy(np.ascontiguousarray(refl.astype(np.uint8))).to(device)
                entry["blocks"] = pupil_blocks(pl["rect"])
                entry["scratch"] = K.pupil_scratch(pl["rect"], entry["blocks"], device)
                entry["out"] = torch.zeros((n_out, 9), dtype=torch.float64, device=device)
                self.pupils.append(entry)
            else:
                entry["term"] = torch.from_numpy(blink_terms(r["saturation"])).to(device)
                entry["out"] = torch.zeros(n_out, dtype=torch.float64, device=device)
                self.blinks.append(entry)
        for rn in runners:
            y0, y1, x0, x1 = rn["rect"]
            self.runs.append({
                "view": rn["view"], "rect": rn["rect"], "ormask": torch.from_numpy(rn["ormask"]).to(device),
                "g2": torch.from_numpy(np.ascontiguousarray(gaussian_fourier_sq(y1 - y0, x1 - x0))).to(device),
                "prev_crop": None, "prev_mean": None,
                "out": torch.zeros((n_out, 2), dtype=torch.int32, device=device)})
        self.ormask = [None if m is None else torch.from_numpy(m).to(device) for m in ormask]
        self._painted = [None] * len(Ly)
        self._means = None

    @property
    def any(self):
        return bool(self.pupils or self.blinks or self.runs)

    @property
    def paints(self):
        return any(m is not None for m in self.ormask)

    # ---- per fetched batch -----------------------------------------------------------------------
    def _carry(self, e, frames, means):
        """keep the last frame's painted crop and mean for the next batch (process.py:589-590)."""
        y0, y1, x0, x1 = e["rect"]
        e["prev_crop"] = torch.bitwise_or(frames[-1, y0:y1, x0:x1], e["ormask"]).contiguous()
        e["prev_mean"] = means[frames.shape[0] - 1:frames.shape[0]].clone()

    def seed(self, halo):
        """halo: per-view [1, H, W] frames f0 - 1 of a shard that starts mid-video (running's predecessor)."""
        for e in self.runs:
            fr = halo[e["view"]]
            means = torch.empty(1, dtype=torch.float32, device=self.device)
            scratch_out = torch.empty((1, 2), dtype=torch.int32, device=self.device)
            K.running(fr, e["rect"], e["ormask"], e["g2"], means, scratch_out)
            self._carry(e, fr, means)

    def chunk(self, frames, t):
        """frames: per-view raw uint8 [n, H, W] device tensors holding frames t .. t+n-1."""
        o = t - self.f0
        for e in self.pupils:
            fr = frames[e["view"]]
            K.pupil(fr, e["rect"], e["mask"], e["dark_offset"], e["sigma"], e["scratch"], e["blocks"],
                    e["out"][o:o + fr.shape[0]], reflector=e["reflector"])
        for e in self.blinks:
            fr = frames[e["view"]]
            K.blink(fr, e["rect"], e["mask"], e["term"], e["out"][o:o + fr.shape[0]])
        for e in self.runs:
            fr = frames[e["view"]]
            n = fr.shape[0]
            if self._means is None or self._means.numel() < n:
                self._means = torch.empty(n, dtype=torch.float32, device=self.device)
            K.running(fr, e["rect"], e["ormask"], e["g2"], self._means, e["out"][o:o + n],
                      prev_crop=e["prev_crop"], prev_mean=e["prev_mean"])
            self._carry(e, fr, self._means)

    def painted(self, frames):
        """The frames stage 3 bins: painted copies for the views that carry paint (R1), else the input."""
        out = list(frames)
        for v, fr in enumerate(frames):
            if self.ormask[v] is None or fr.shape[0] == 0:
                continue
            buf = self._painted[v]
            if buf is None or buf.shape[0] < fr.shape[0]:
                buf = self._painted[v] = torch.empty((fr.shape[0],) + tuple(fr.shape[1:]), dtype=torch.uint8,
                                                     device=self.device)
            out[v] = K.paint(fr, self.ormask[v], buf)
        return out

    def sweep(self, source, fetch_frames=2368):
        """Stand-alone pass for runs without an SVD stage 3 (fullSVD off and no motion ROI)."""
        source.begin_sweep()
        if self.f0 > 0:
            self.seed(source.fetch(self.f0 - 1, self.f0, self.device))
        for t in range(self.f0, self.f1, fetch_frames):
            self.chunk(source.fetch(t, min(self.f1, t + fetch_frames), self.device), t)

    # ---- results ---------------------------------------------------------------------------------
    def device_outputs(self):
        """(pupil [n, 9] float64, blink [n] float64, running [n, 2] int32) lists of device tensors."""
        return [e["out"] for e in self.pupils], [e["out"] for e in self.blinks], [e["out"] for e in self.runs]


def assemble(pupil_raw, blink_raw, running_raw, smooth=True):
    """Host arrays in the reference's proc layout (process.py:358-385, 843-851) from the raw outputs."""
    pups = []
    for a in pupil_raw:
        a = np.asarray(a)
        p = {"area": a[:, 2].copy(), "com": a[:, 0:2].copy(), "axdir": a[:, 3:7].reshape(-1, 2, 2).copy(),
             "axlen": a[:, 7:9].copy()}
        if smooth:
            p["area_smooth"] = smooth_trace(p["area"])
            p["com_smooth"] = p["com"].copy()
            p["com_smooth"][:, 0] = smooth_trace(p["com_smooth"][:, 0])
            p["com_smooth"][:, 1] = smooth_trace(p["com_smooth"][:, 1])
        pups.append(p)
    blinks = [np.asarray(b, dtype=np.float64).copy() for b in blink_raw]
    runs = [np.asarray(r).astype(np.float64) for r in running_raw]
    return pups, blinks, runs
