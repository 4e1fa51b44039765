"""TEST INFRASTRUCTURE — CPU (NumPy) restatement of the reference's full-resolution ROI
processors that ride on stage 3's frame loop: pupil (rind 0), blink (rind 2), running (rind 3).

Checker only (tests/, smoke); the product never imports it.  Pinned against the unmodified
reference by `tests/golden/fullres_*.npz` (scripts/make_golden_fullres.py runs
`/root/reference/facemap/process.py` in-process on the same bytes).

Reference behaviour restated here, with the quirks that the CUDA path has to reproduce:

  R1  process.py:543, 567 paint the pixels of the ROI rectangle that lie outside the ROI's ellipse
      with 255 *in the chunk buffer itself* (the ROI crop is a view).  Later ROIs on the same view
      (order: all pupils, all blinks, running) and stage 3's binning / motion energy / projections
      therefore see the painted frames.  Stages 1 and 2 read fresh frames and do not.
  R2  running.py:118-127 calls `fft2(X[t])` / `ifft2(Xout[t])` and drops the results, so no
      transform happens: the "correlation" is sign(x_t - mean_t) * sign(x_{t-1} - mean_{t-1}) *
      fhg32^2 evaluated pixel-wise on the four corner blocks, and (dy, dx) is its first argmax.
  R3  blink: `255 - roi - (255 - saturation)` is evaluated in the dtype NumPy promotes to — float64
      when saturation is a Python float (what the GUI stores, guiparts.py:97), uint8 WRAPPING when
      it is a Python int in [0, 255].
  R4  pupil.py:8-92 mixes precisions: weights are float32 (sum, normalisation, square root),
      moments float64.
"""
from __future__ import annotations

import numpy as np

PUPIL, MOTION, BLINK, RUNNING = 0, 1, 2, 3


# ------------------------------------------------------------------------------------------
# ROI bookkeeping
# ------------------------------------------------------------------------------------------
def roi_rect(r):
    """process.py:538-542 — the crop is [yrange[0], yrange[-1]] x [xrange[0], xrange[-1]] inclusive."""
    return int(r["yrange"][0]), int(r["yrange"][-1]) + 1, int(r["xrange"][0]), int(r["xrange"][-1]) + 1


def processing_order(rois):
    """process.py:354-385, 413-425 — pupils first, then blinks, then running, each in list order."""
    rois = rois or []
    return ([i for i, r in enumerate(rois) if r["rind"] == PUPIL],
            [i for i, r in enumerate(rois) if r["rind"] == BLINK],
            [i for i, r in enumerate(rois) if r["rind"] == RUNNING])


# ------------------------------------------------------------------------------------------
# blink  (process.py:557-572)
# ------------------------------------------------------------------------------------------
def blink_chunk(crop, saturation):
    """crop: painted uint8 (n, ny, nx) -> float64 (n,) (stored into a float64 array, process.py:382)."""
    val = np.maximum(0, 255 - crop - (255 - saturation))
    return val.sum(axis=(-2, -1)).astype(np.float64)


# ------------------------------------------------------------------------------------------
# running  (running.py:91-133 as executed, see R2)
# ------------------------------------------------------------------------------------------
def running_filter(Ly, Lx):
    """running.py:29-42 — Gaussian (std 2 px) in the Fourier domain, float64."""
    ax = np.abs(np.arange(Lx) - np.arange(Lx).mean())
    ay = np.abs(np.arange(Ly) - np.arange(Ly).mean())
    xx, yy = np.meshgrid(ax, ay)
    h = np.exp(-np.square(yy / 2) / 2) * np.exp(-np.square(xx / 2) / 2)
    h /= h.sum()
    return np.real(np.fft.fft2(np.fft.ifftshift(h)))


def running_taper(Ly, Lx):
    """running.py:14-26 with sig = 0.01*sqrt(Ly*Lx) (running.py:100); strictly positive."""
    sig = (Ly * Lx) ** 0.5 * 0.01
    ax = np.abs(np.arange(Lx) - np.arange(Lx).mean())
    ay = np.abs(np.arange(Ly) - np.arange(Ly).mean())
    xx, yy = np.meshgrid(ax, ay)
    return (1.0 / (1.0 + np.exp((yy - (ay.max() - 2 * sig)) / sig))) * \
           (1.0 / (1.0 + np.exp((xx - (ax.max() - 2 * sig)) / sig)))


def running_chunk(imr):
    """imr: uint8 (nt, Ly, Lx) including the carried previous frame -> (dy, dx) int32 (nt-1,)."""
    nt, Ly, Lx = imr.shape
    lcorr = min(int(np.floor(Ly * 0.4)), int(np.floor(Lx * 0.4)))
    taper = running_taper(Ly, Lx).astype(np.complex64)
    g = running_filter(Ly, Lx).astype(np.complex64)
    X = imr.astype(np.float32)
    X -= X.mean(axis=-1).mean(axis=-1)[:, None, None]
    X = X.astype(np.complex64) * taper                      # multiplytype, running.py:45-51
    mag = np.abs(X)                                          # float32
    X = ((X.astype(np.complex128) / (1e-20 + mag.astype(np.float64))) * g).astype(np.complex64)  # phase_norm
    out = X[1:] * np.conj(X[:-1])                            # no (i)fft2 happens (R2)
    a, b = out[:, :lcorr + 1, :lcorr + 1], out[:, :lcorr + 1, -lcorr:]
    c, d = out[:, -lcorr:, :lcorr + 1], out[:, -lcorr:, -lcorr:]
    cc = np.real(np.block([[d, c], [b, a]]))
    w = 2 * lcorr + 1
    flat = cc.reshape(nt - 1, -1).argmax(axis=1)
    return (flat // w - lcorr).astype(np.int32), (flat % w - lcorr).astype(np.int32)


# ------------------------------------------------------------------------------------------
# pupil  (pupil.py:8-132)
# ------------------------------------------------------------------------------------------
def reflector_pixels(r):
    """utils.py:666-700 (rdict branch) — (rows, cols) of reflector pixels inside the pupil crop."""
    ny, nx = r["yrange"].size, r["xrange"].size
    m = np.zeros((ny, nx), bool)
    for q in r.get("reflector", []) or []:
        e, ry, rx = q["ellipse"].copy(), q["yrange"].copy(), q["xrange"].copy()
        kx = (rx >= 0) & (rx < nx)
        ky = (ry >= 0) & (ry < ny)
        e = e[:, kx][ky, :]
        m[np.ix_(ry[ky], rx[kx])] |= e
    return m.nonzero()


def _moments(lam, iy, ix):
    """weights float32, coordinates int -> mean (float64 pair), offsets, 2x2 scatter + 0.01 I (pupil.py:26-32)."""
    mu = [(lam * iy).sum(), (lam * ix).sum()]
    d = np.concatenate(((iy - mu[0])[:, None], (ix - mu[1])[:, None]), axis=1)
    w = d * lam[:, None] ** 0.5
    return mu, d, w.T @ w + 0.01 * np.eye(2)


def pupil_frame(im, sigma, missing=None):
    """One frame of pupil.py:8-92.  im: float32 (ny, nx) darkness-above-threshold image.
    Returns com (2,), axlen (2,), axdir (2, 2)."""
    iy, ix = im.nonzero()
    have_miss = missing is not None and len(missing) > 0 and len(missing[0]) > 0
    if have_miss:
        my, mx = missing
        keep = ~np.isin(iy * im.shape[1] + ix, my * im.shape[1] + mx)
        iy, ix = iy[keep], ix[keep]
    lam0 = im[iy, ix].copy()
    n0 = iy.size
    iy0, ix0 = iy, ix
    lam = lam0 / lam0.sum()
    mu, d, sig = _moments(lam, iy, ix)
    for _ in range(5):
        dd = ((d @ np.linalg.inv(sig)) * d).sum(axis=1)[:n0]
        lam0[~(dd <= 2 * sigma ** 2)] = 0
        med = np.median(lam0[dd <= sigma ** 2])
        lam = lam0.copy()
        if have_miss:
            dm = np.concatenate(((my - mu[0])[:, None], (mx - mu[1])[:, None]), axis=1)
            ddm = ((dm @ np.linalg.inv(sig)) * dm).sum(axis=1)
            inside = ddm <= 1.15 * sigma ** 2
            lamm = np.zeros(my.size, np.float32)
            lamm[inside] = med * 1.1
            iy, ix = np.concatenate((iy0, my)), np.concatenate((ix0, mx))
            lam = np.concatenate((lam0, lamm))
        lam /= lam.sum()
        mu, d, sig = _moments(lam, iy, ix)
    sv, u = np.linalg.eig(sigma ** 2 * sig)
    sv = np.real(sv)
    sv = sv.min() * np.minimum(4, sv / sv.min())
    return np.array(mu), sv[::-1], u[:, ::-1]


def pupil_chunk(crop, saturation, sigma, missing=None):
    """pupil.py:95-132 — crop: painted uint8 (n, ny, nx).  Any numerical failure -> NaNs for the frame."""
    n = crop.shape[0]
    com = np.full((n, 2), np.nan)
    area = np.full(n, np.nan)
    axdir = np.full((n, 2, 2), np.nan)
    axlen = np.full((n, 2), np.nan)
    f = crop.astype(np.float32)
    with np.errstate(all="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for t in range(n):
                try:
                    im = np.maximum(0, (255.0 - f[t]) - (255.0 - saturation))
                    mu, sv, u = pupil_frame(im, sigma, missing)
                except Exception:
                    continue
                com[t], axlen[t], axdir[t] = mu, sv, u
                area[t] = np.pi * (sv[0] * sv[1]) ** 0.5
    return com, area, axdir, axlen


def smooth_trace(x, win=30):
    """pupil.py:135-166 — running nan-median (zero padded), gaps interpolated, outliers (> std/2
    from the median trace) replaced.  Returns the smoothed copy."""
    x = np.array(x, dtype=np.float64)
    nt = x.size
    half = win // 2
    stack = np.zeros((win, nt))
    for row, k in enumerate(range(-half, half)):
        if k < 0:
            stack[row, :k] = x[-k:]
        elif k > 0:
            stack[row, k:] = x[:-k]
        else:
            stack[row] = x
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        med = np.nanmedian(stack, axis=0)
    bad = (np.isnan(x) | np.isnan(med)).nonzero()[0]
    good = (~np.isnan(x) & ~np.isnan(med)).nonzero()[0]
    x[bad] = np.interp(bad, good, med[good]) if good.size > 0 else 0
    med[bad] = x[bad]
    swap = np.abs(x - med) > x.std() / 2
    x[swap] = med[swap]
    return x


# ------------------------------------------------------------------------------------------
# the per-chunk driver that stage 3 calls (process.py:410-425)
# ------------------------------------------------------------------------------------------
class FullResRois:
    """Holds the output arrays and the running carry; `chunk(t, ims)` processes one stage-3 chunk
    and PAINTS `ims` in place (R1), so call it before binning."""

    def __init__(self, rois, N):
        self.rois = rois or []
        self.pup_i, self.bl_i, self.run_i = processing_order(self.rois)
        self.N = N
        self.pupil = [{"area": np.zeros(N), "com": np.zeros((N, 2)), "axdir": np.zeros((N, 2, 2)),
                       "axlen": np.zeros((N, 2))} for _ in self.pup_i]
        self.blink = [np.zeros(N) for _ in self.bl_i]
        self.running = [np.zeros((N, 2)) for _ in self.run_i]
        self.carry = [None] * len(self.run_i)
        self.missing = [reflector_pixels(self.rois[i]) if "reflector" in self.rois[i] else None
                        for i in self.pup_i]

    @property
    def any(self):
        return bool(self.pup_i or self.bl_i or self.run_i)

    def chunk(self, t, ims, first):
        n = ims[0].shape[0]
        for k, i in enumerate(self.pup_i):
            r = self.rois[i]
            y0, y1, x0, x1 = roi_rect(r)
            crop = ims[r["ivid"]][:, y0:y1, x0:x1]
            crop[:, ~r["ellipse"]] = 255
            com, area, axdir, axlen = pupil_chunk(crop, r["saturation"], r["pupil_sigma"], self.missing[k])
            self.pupil[k]["com"][t:t + n] = com
            self.pupil[k]["area"][t:t + n] = area
            self.pupil[k]["axdir"][t:t + n] = axdir
            self.pupil[k]["axlen"][t:t + n] = axlen
        for k, i in enumerate(self.bl_i):
            r = self.rois[i]
            y0, y1, x0, x1 = roi_rect(r)
            crop = ims[r["ivid"]][:, y0:y1, x0:x1]
            crop[:, ~r["ellipse"]] = 255
            self.blink[k][t:t + n] = blink_chunk(crop, r["saturation"])
        for k, i in enumerate(self.run_i):
            r = self.rois[i]
            y0, y1, x0, x1 = roi_rect(r)
            crop = ims[r["ivid"]][:, y0:y1, x0:x1]
            if not first:
                crop = np.concatenate((self.carry[k][None], crop), axis=0)
            self.carry[k] = crop[-1].copy()
            dy, dx = running_chunk(crop)
            a = t if not first else t + 1
            self.running[k][a:a + dy.size, 0] = dy
            self.running[k][a:a + dy.size, 1] = dx

    def finish(self):
        """process.py:843-851 — smoothed pupil traces (the blink smoothing there has no effect)."""
        for p in self.pupil:
            p["area_smooth"] = smooth_trace(p["area"])
            p["com_smooth"] = p["com"].copy()
            p["com_smooth"][:, 0] = smooth_trace(p["com_smooth"][:, 0])
            p["com_smooth"][:, 1] = smooth_trace(p["com_smooth"][:, 1])
        return self.pupil, self.blink, self.running
