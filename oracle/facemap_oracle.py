"""TEST INFRASTRUCTURE — CPU (NumPy/SciPy) restatement of the reference's motion/movie-SVD path.

This file is the *checker* for the CUDA path (tests/, __graft_entry__.smoke(), and the
cpu_baseline / --impl reference legs of bench.py).  The product (`facemap_b200/`) never imports
it.  It is pinned against the unmodified reference by `tests/golden/*.npz`, which were produced
by `scripts/make_golden.py` running `/root/reference/facemap/process.py` in-process (see
`oracle/ref_harness.py`); `tests/test_oracle_golden.py` replays them.

Every function cites the reference lines it restates (paths relative to /root/reference, or
`sklearn/...` for scikit-learn 1.9.0, the un-vendored dependency the SVD arithmetic lives in).

SVD back-ends (`svd=`):
  "stock"  the reference's live code path: sklearn PCA(svd_solver="randomized",
           random_state=RandomState(0))._fit  (facemap/utils.py:773-775)            -> oracle O1
  "exact"  float64 LAPACK SVD of the column-centred matrix + same sign rule         -> oracle O2
  "gram"   float64 Gram + eigh of the smaller side (the algorithm the CUDA path runs;
           facemap/utils.py:753-772 dead code / matlab/utils/svdecon.m:16-43)
"""
from __future__ import annotations

import numpy as np

NCOMPS = 500          # facemap/process.py:764
CHUNK_SVD = 1000      # facemap/process.py:134
BUDGET_SVD = 15000    # facemap/process.py:135
NC_CHUNK = 250        # facemap/process.py:136
CHUNK_PROJ = 500      # facemap/process.py:394


# ------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------
def binned_geometry(Ly, Lx, sbin):
    """facemap/process.py:16-26 — binned sizes and per-view column ranges in the concatenated
    pixel axis."""
    Lyb = np.array([int(np.floor(h / sbin)) for h in Ly], np.int32)
    Lxb = np.array([int(np.floor(w / sbin)) for w in Lx], np.int32)
    offs = np.concatenate(([0], np.cumsum(Lyb.astype(np.int64) * Lxb)))
    cols = [np.arange(offs[i], offs[i + 1], 1, int) for i in range(len(Ly))]
    return Lyb, Lxb, cols


def canvas_layout(Lyb, Lxb):
    """facemap/utils.py:703-748 — tile the views on one canvas: biggest first, fill a column
    of `gridy` views top to bottom, then move right by the widest view of that column."""
    Lyb = np.asarray(Lyb)
    Lxb = np.asarray(Lxb)
    nv = Lyb.size
    area = Lyb * Lxb
    if nv in (2, 3):
        gridy = 1
    else:
        gridy = int(np.round(nv ** 0.5 * 0.75))
    sy = np.zeros(Lyb.shape, int)
    sx = np.zeros(Lxb.shape, int)
    done = np.zeros(nv, bool)
    LY = 0
    ycur = xcur = 0
    colw = 0
    row_in_col = 0
    col = 0
    while not done.all():
        a = area.copy()
        a[done] = 0
        i = int(np.argmax(a))
        done[i] = True
        if row_in_col == 0:
            ycur, colw = 0, 0
        if col == 0:
            xcur = 0
        sy[i], sx[i] = ycur, xcur
        ycur += Lyb[i]
        colw = max(colw, Lxb[i])
        if row_in_col == gridy - 1 or done.all():
            xcur += colw
        LY = max(LY, ycur)
        row_in_col += 1
        if row_in_col >= gridy:
            row_in_col = 0
            col += 1
    return LY, xcur, sy, sx


def scatter_canvas(X, LY, LX, sy, sx, Lyb, Lxb, cols):
    """facemap/utils.py:622-630 — (sum npix, n) rows -> (LY, LX, n) canvas, float32."""
    out = np.zeros((LY, LX, X.shape[-1]), np.float32)
    for i in range(len(Lyb)):
        out[sy[i]:sy[i] + Lyb[i], sx[i]:sx[i] + Lxb[i]] = X[cols[i]].reshape(Lyb[i], Lxb[i], -1)
    return out


def roi_bin_ranges(roi, sbin):
    """facemap/process.py:716-723 incl. quirk Q5 (x stop divided *after* the floor)."""
    yb = np.arange(np.floor(roi["yrange"][0] / sbin), np.floor(roi["yrange"][-1] / sbin)).astype(int)
    xb = np.arange(np.floor(roi["xrange"][0] / sbin), np.floor(roi["xrange"][-1]) / sbin).astype(int)
    return yb, xb


# ------------------------------------------------------------------------------------------
# binning
# ------------------------------------------------------------------------------------------
def bin_frames(im, sbin, Lyb, Lxb):
    """facemap/process.py:34-43 — crop, block-mean (float64 for sbin>1, float32 for sbin==1),
    flatten row-major."""
    if sbin > 1:
        v = im[:, :Lyb * sbin, :Lxb * sbin].reshape(-1, Lyb, sbin, Lxb, sbin)
        b = v.mean(axis=-1).mean(axis=-2)
    else:
        b = im.astype(np.float32)
    return b.reshape(-1, Lyb * Lxb)


def bin_sums_int(im, sbin, Lyb, Lxb):
    """Integer block sums Σ_t[y,x]: the bit-exact form the CUDA kernel is held
    to.  `bin_frames == bin_sums_int / sbin²` exactly when sbin is a power of two."""
    v = im[:, :Lyb * sbin, :Lxb * sbin].reshape(-1, Lyb, sbin, Lxb, sbin)
    return v.sum(axis=(2, 4), dtype=np.int64).reshape(-1, Lyb * Lxb)


# ------------------------------------------------------------------------------------------
# frame access (in-memory stand-in for utils.get_frames, facemap/utils.py:522-553)
# ------------------------------------------------------------------------------------------
class ArrayVideo:
    """views[v] = uint8 (N, H_v, W_v); sequential blocks already concatenated along time.
    `block_frames` (frames per sequential block) only matters for stage 1's nt0."""

    def __init__(self, views, block_frames=None):
        self.views = list(views)
        self.N = int(self.views[0].shape[0])
        self.Ly = [int(v.shape[1]) for v in self.views]
        self.Lx = [int(v.shape[2]) for v in self.views]
        self.block_frames = list(block_frames) if block_frames is not None else [self.N]

    def get(self, t0, n):
        """Frames t0 .. t0+n-1 clamped to the video (utils.py:523-525, 551-553)."""
        a = max(0, min(self.N - 1, t0))
        b = max(0, min(self.N - 1, t0 + n - 1))
        return [v[a:b + 1] for v in self.views]


def read_frames(containers, cumframes, shapes, t0, n):
    """utils.get_frames (facemap/utils.py:522-553) on duck-typed captures `containers[block][view]`
    (`.get(POS_FRAMES)`, `.set`, `.read`), into freshly zeroed chunk buffers like imall_init (process.py:46-50).

    Restated behaviours: the request is clamped to the video and re-expanded to a contiguous range (:523-525);
    sequential files are stitched in order (:527-535); a frame that fails to decode takes the previous slot of the
    buffer — slot -1, i.e. the still-zero LAST slot, when it is the very first of the request — and reading of that
    file stops there, so the rest of its range keeps the buffer's zeros (:541-548); the buffers are cut to the
    frames that exist (:551-553).  Returns one uint8 (T', H, W) array per view."""
    import cv2

    cumframes = np.asarray(cumframes).astype(int)
    N = int(cumframes[-1])
    a = max(0, min(N - 1, int(t0)))
    b = max(0, min(N - 1, int(t0) + int(n) - 1))
    cframes = np.arange(a, b + 1).astype(int)
    ivids = (cframes[np.newaxis, :] >= cumframes[1:, np.newaxis]).sum(axis=0)
    bufs = [np.zeros((int(n), h, w), np.uint8) for (h, w) in shapes]
    nk = 0
    for ii in range(len(shapes)):
        nk = 0
        for blk in np.unique(ivids):
            cfr = cframes[ivids == blk]
            start, count = int(cfr[0] - cumframes[blk]), int(cfr.size)
            cap = containers[blk][ii]
            if int(cap.get(cv2.CAP_PROP_POS_FRAMES)) != start:
                cap.set(cv2.CAP_PROP_POS_FRAMES, start)
            for fc in range(count):
                ok, frame = cap.read()
                if not ok:
                    bufs[ii][nk + fc] = bufs[ii][nk + fc - 1]
                    break
                bufs[ii][nk + fc] = frame if frame.ndim == 2 else cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
            nk += count
    return [buf[:nk].copy() for buf in bufs]


# ------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------
def stage1_sample_times(N, block_frames):
    """facemap/process.py:61-67."""
    nf = min(1000, N)
    nt0 = int(min(100, min(block_frames)))
    nsegs = int(np.floor(nf / nt0))
    tf = np.floor(np.linspace(0, N - nt0, nsegs)).astype(int)
    return nt0, tf


def stage1_means(video, sbin):
    """facemap/process.py:53-102 — avgframe / avgmotion, float32 accumulators updated with
    float64 addends, divided by the segment count in float32."""
    Lyb, Lxb, cols = binned_geometry(video.Ly, video.Lx, sbin)
    nt0, tf = stage1_sample_times(video.N, video.block_frames)
    p = int((Lyb.astype(np.int64) * Lxb).sum())
    avgframe = np.zeros(p, np.float32)
    avgmotion = np.zeros(p, np.float32)
    for t in tf:
        ims = video.get(int(t), nt0)
        for v, im in enumerate(ims):
            b = bin_frames(im, sbin, Lyb[v], Lxb[v])
            avgframe[cols[v]] += b.mean(axis=0)
            avgmotion[cols[v]] += np.abs(np.diff(b, axis=0)).mean(axis=0)
    avgframe /= float(len(tf))
    avgmotion /= float(len(tf))
    return [avgframe[c] for c in cols], [avgmotion[c] for c in cols]


# ------------------------------------------------------------------------------------------
# SVD back-ends
# ------------------------------------------------------------------------------------------
def _sign_fix(U, Vt):
    """sklearn/utils/extmath.py:924-982, u_based_decision=False: the largest-|.| entry of each
    right singular vector is made positive."""
    idx = np.argmax(np.abs(Vt), axis=1)
    sg = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    sg = np.where(sg == 0, 1.0, sg).astype(Vt.dtype)
    return U * sg[None, :], Vt * sg[:, None]


def svd_centered(X, k, svd="exact"):
    """facemap/utils.py:751-776 -> sklearn/decomposition/_pca.py:696-788: truncated SVD of
    X - mean(X, axis=0).  Returns (U[n,k], S[k], Vt[k,m]) in X's float dtype."""
    k = int(k)
    if svd == "stock":
        from sklearn.decomposition import PCA

        np.random.seed(0)
        return PCA(n_components=k, svd_solver="randomized",
                   random_state=np.random.RandomState(0))._fit(X)[:3]
    dt = X.dtype if X.dtype in (np.float32, np.float64) else np.float64
    X64 = np.asarray(X, np.float64)
    Xc = X64 - X64.mean(axis=0)
    if svd == "exact":
        import scipy.linalg

        U, S, Vt = scipy.linalg.svd(Xc, full_matrices=False, lapack_driver="gesdd")
        U, S, Vt = U[:, :k], S[:k], Vt[:k]
    elif svd == "gram":
        n, m = Xc.shape
        if m <= n:                      # Gram over columns: (m, m)
            lam, W = np.linalg.eigh(Xc.T @ Xc)
            lam, W = lam[::-1][:k], W[:, ::-1][:, :k]
            S = np.sqrt(np.maximum(lam, 0.0))
            U = (Xc @ W) / np.where(S > 0, S, 1.0)
            Vt = W.T
        else:                           # Gram over rows: (n, n)
            lam, W = np.linalg.eigh(Xc @ Xc.T)
            lam, W = lam[::-1][:k], W[:, ::-1][:, :k]
            S = np.sqrt(np.maximum(lam, 0.0))
            U = W
            Vt = (Xc.T @ W).T / np.where(S > 0, S, 1.0)[:, None]
    else:
        raise ValueError(svd)
    U, Vt = _sign_fix(U, Vt)
    return U.astype(dt), S.astype(dt), Vt.astype(dt)


# ------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------
def stage2_plan(N, ncomps=NCOMPS):
    """facemap/process.py:134-141 — chunk length, count, components kept per chunk, starts."""
    nt0 = min(CHUNK_SVD, N)
    nsegs = int(min(np.floor(BUDGET_SVD / nt0), np.floor(N / nt0)))
    nc = min(NC_CHUNK, nt0 - 1)
    if nsegs == 1:
        nc = min(ncomps, nt0 - 1)
    tf = np.floor(np.linspace(0, N - nt0 - 1, nsegs)).astype(int)
    return nt0, nsegs, nc, tf


def chunk_matrices(ims, sbin, Lyb, Lxb, cols, avgframe, avgmotion, want_mot, want_mov):
    """facemap/process.py:191-212 — per chunk: X_mot = |diff(bin)| - avgmotion,
    X_mov = bin[1:] - avgframe, float32 (rows = frames-1, cols = all views' pixels).
    Also returns the per-view un-cast arrays (float64 when sbin>1) for the ROI path."""
    T = ims[0].shape[0]
    p = int(cols[-1][-1]) + 1 if len(cols[-1]) else 0
    X_mot = np.zeros((T - 1, p), np.float32) if want_mot else None
    X_mov = np.zeros((T - 1, p), np.float32) if want_mov else None
    per_view_mot, per_view_mov = [], []
    for v, im in enumerate(ims):
        b = bin_frames(im, sbin, Lyb[v], Lxb[v])
        if want_mot:
            d = np.abs(np.diff(b, axis=0))
            d -= avgmotion[v].reshape(-1)
            X_mot[:, cols[v]] = d
            per_view_mot.append(d)
        if want_mov:
            f = b[1:, :].copy()
            f -= avgframe[v].reshape(-1)
            X_mov[:, cols[v]] = f
            per_view_mov.append(f)
    return X_mot, X_mov, per_view_mot, per_view_mov


def stage2_masks(video, sbin, avgframe, avgmotion, motSVD=True, movSVD=False, rois=None,
                 fullSVD=True, ncomps=NCOMPS, svd="exact"):
    """facemap/process.py:105-299 — two-level SVD.  Returns (U_mot, U_mov, S_mot, S_mov) lists
    laid out like the reference ([0] = full-frame, then one entry per motion ROI)."""
    N = video.N
    Lyb, Lxb, cols = binned_geometry(video.Ly, video.Lx, sbin)
    nt0, nsegs, nc, tf = stage2_plan(N, ncomps)
    mot_rois = [r for r in (rois or []) if r["rind"] == 1]
    mods = [m for m, on in (("mot", motSVD), ("mov", movSVD)) if on]
    blocks = {m: [[] for _ in range(1 + len(mot_rois))] for m in mods}

    for t in tf:
        ims = video.get(int(t), nt0)
        need = fullSVD or len(mot_rois) > 0
        if not need:
            continue
        X_mot, X_mov, pv_mot, pv_mov = chunk_matrices(
            ims, sbin, Lyb, Lxb, cols, avgframe, avgmotion, motSVD, movSVD)
        for m, X, pv in (("mot", X_mot, pv_mot), ("mov", X_mov, pv_mov)):
            if m not in mods:
                continue
            for j, r in enumerate(mot_rois):          # process.py:213-242
                v = r["ivid"]
                yb, xb = r["yrange_bin"], r["xrange_bin"]
                cube = pv[v].reshape(-1, Lyb[v], Lxb[v])
                sub = cube[:, yb[0]:yb[-1] + 1, xb[0]:xb[-1] + 1]
                sub = sub.reshape(sub.shape[0], -1)
                kk = min(nc, sub.shape[-1])
                u, s, _ = svd_centered(sub.T, kk, svd)
                blocks[m][1 + j].append((u * s).astype(np.float32))
            if fullSVD:                                # process.py:246-259
                kk = min(nc, X.shape[-1])
                u, s, _ = svd_centered(X.T, kk, svd)
                blocks[m][0].append((u * s).astype(np.float32))

    out_U = {"mot": [], "mov": []}
    out_S = {"mot": np.zeros(500, np.float32), "mov": np.zeros(500, np.float32)}
    for m in mods:
        for idx in range(1 + len(mot_rois)):
            if idx == 0 and not fullSVD:
                out_U[m].append(np.zeros((0, 1), np.float32))
                continue
            # C-ordered like the reference's preallocated U buffer (process.py:146-151):
            # BLAS results depend on the operand layout, and 'stock' is held bit-for-bit
            Y = np.ascontiguousarray(np.concatenate(blocks[m][idx], axis=1))
            # quirk Q8: the merge loop runs over range(len(U_mot)) (process.py:264) and indexes BOTH lists with it.
            # With the motion SVD switched off U_mot holds only the ROI placeholders (process.py:146-178), so the
            # movie targets from position len(U_mot) on — everything when there is no ROI, the last ROI otherwise —
            # are never merged and stay the stacked U S blocks of their chunks.
            merged = idx < (1 if motSVD else 0) + len(mot_rois)
            if nsegs > 1 and merged:                   # process.py:264-295
                u, s, _ = svd_centered(Y, min(ncomps, Y.shape[0] - 1), svd)
                out_U[m].append(u)
                out_S[m] = s                           # quirk Q3: last one wins
            else:
                # quirk Q6: a single chunk is never merged — U S of that chunk as is, singular values stay zero.
                # The full-frame buffer was preallocated nsegs * nc wide (process.py:146-151) and is only cut to
                # the filled columns inside the merge (process.py:267): with fewer binned pixels than nc the
                # trailing columns stay zero.  ROI buffers are allocated to size (process.py:163-178).
                if idx == 0 and Y.shape[1] < nsegs * nc:
                    Y = np.ascontiguousarray(np.pad(Y, ((0, 0), (0, nsegs * nc - Y.shape[1]))))
                out_U[m].append(Y)
    # quirk Q7: ROI placeholders exist for a modality that is off
    for m in ("mot", "mov"):
        if m not in mods:
            for r in mot_rois:
                nyb, nxb = r["yrange_bin"].size, r["xrange_bin"].size
                out_U[m].append(np.zeros((nyb * nxb, nsegs * min(nc, nyb * nxb)), np.float32))
    return out_U["mot"], out_U["mov"], out_S["mot"], out_S["mov"]


# ------------------------------------------------------------------------------------------
# stage 3
# ------------------------------------------------------------------------------------------
def stage3_project(video, sbin, avgframe, avgmotion, U_mot, U_mov, motSVD=True, movSVD=False,
                   rois=None, fullSVD=True, fullres=None):
    """facemap/process.py:302-530 (SVD/motion parts) — all frames in chunks of 500 with a
    one-frame carry; first projected row duplicated (Q2); motion[499] hole (Q1);
    avg arrays reshaped in place when a motion ROI sits on the view (Q4).
    `fullres` (oracle/roi_oracle.FullResRois) processes pupil/blink/running ROIs on each chunk
    first and paints the chunk in place (process.py:410-425, 543, 567), which the binning sees."""
    N = video.N
    Lyb, Lxb, cols = binned_geometry(video.Ly, video.Lx, sbin)
    p = int((Lyb.astype(np.int64) * Lxb).sum())
    mot_rois = [r for r in (rois or []) if r["rind"] == 1]
    if fullSVD:
        V_mot = [np.zeros((N, U_mot[0].shape[-1]), np.float32)] if motSVD else []
        V_mov = [np.zeros((N, U_mov[0].shape[-1]), np.float32)] if movSVD else []
        M = [np.zeros(N, np.float32)]
    else:
        V_mot = [np.zeros((0, 1), np.float32)] if motSVD else []
        V_mov = [np.zeros((0, 1), np.float32)] if movSVD else []
        M = [np.zeros((0,), np.float32)]
    for j, r in enumerate(mot_rois):
        if motSVD:
            V_mot.append(np.zeros((N, U_mot[1 + j].shape[1]), np.float32))
        if movSVD:
            V_mov.append(np.zeros((N, U_mov[1 + j].shape[1]), np.float32))
        M.append(np.zeros(N, np.float32))

    carry = [None] * len(video.views)
    t = 0
    nseg = int(np.ceil(N / CHUNK_PROJ))
    for n in range(nseg):
        ims = video.get(t, CHUNK_PROJ)
        nt1 = ims[0].shape[0]
        if fullres is not None and fullres.any:
            ims = [im.copy() for im in ims]
            fullres.chunk(t, ims, first=(n == 0))
        rows = nt1 if n > 0 else nt1 - 1
        if fullSVD:
            X_mot = np.zeros((rows, p), np.float32)
            X_mov = np.zeros((rows, p), np.float32)
        for v, im in enumerate(ims):
            on_view = [j for j, r in enumerate(mot_rois) if r["ivid"] == v]
            if not (fullSVD or on_view):
                continue
            b = bin_frames(im, sbin, Lyb[v], Lxb[v])
            if n > 0:
                b = np.concatenate((carry[v][None, :], b), axis=0)
            carry[v] = b[-1].copy()
            d = np.abs(np.diff(b, axis=0)) if motSVD else None
            f = b[1:, :] if movSVD else None
            if fullSVD:
                if motSVD:
                    M[0][t:t + d.shape[0]] += d.sum(axis=-1)
                    X_mot[:, cols[v]] = d - avgmotion[v].reshape(-1)
                if movSVD:
                    X_mov[:, cols[v]] = f - avgframe[v].reshape(-1)
            if on_view:
                if motSVD:
                    d3 = d.reshape(-1, Lyb[v], Lxb[v])
                    avgmotion[v] = avgmotion[v].reshape(Lyb[v], Lxb[v])      # Q4
                if movSVD:
                    f3 = f.reshape(-1, Lyb[v], Lxb[v])
                    avgframe[v] = avgframe[v].reshape(Lyb[v], Lxb[v])        # Q4
                for j in on_view:
                    r = mot_rois[j]
                    yb, xb = r["yrange_bin"], r["xrange_bin"]
                    ys, xs = slice(yb[0], yb[-1] + 1), slice(xb[0], xb[-1] + 1)
                    if motSVD:
                        sub = d3[:, ys, xs]
                        M[1 + j][t:t + sub.shape[0]] = sub.sum(axis=(-2, -1))
                        sub -= avgmotion[v][ys, xs]   # in-place on the view, like process.py:483
                        vp = sub.reshape(sub.shape[0], -1) @ U_mot[1 + j]
                        if n == 0:
                            vp = np.concatenate((vp[:1], vp), axis=0)
                        V_mot[1 + j][t:t + vp.shape[0]] = vp
                    if movSVD:
                        sub = f3[:, ys, xs]
                        sub -= avgframe[v][ys, xs]
                        vp = sub.reshape(sub.shape[0], -1) @ U_mov[1 + j]
                        if n == 0:
                            vp = np.concatenate((vp[:1], vp), axis=0)
                        V_mov[1 + j][t:t + vp.shape[0]] = vp
        if fullSVD:
            for on, X, U, V in ((motSVD, X_mot, U_mot, V_mot), (movSVD, X_mov, U_mov, V_mov)):
                if on:
                    vp = X @ U[0]
                    if n == 0:
                        vp = np.concatenate((vp[:1], vp), axis=0)
                    V[0][t:t + vp.shape[0]] = vp
        t += nt1
    return V_mot, V_mov, M


# ------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------
def run_oracle(views, sbin=1, motSVD=True, movSVD=False, rois=None, fullSVD=True,
               svd="exact", block_frames=None, stages=("mean", "svd", "proj")):
    """facemap/process.py:638-903 on in-memory views; returns the proc dict (§A.2 keys that
    belong to this path).  `views[v]` = uint8 (N, H_v, W_v)."""
    import copy

    video = ArrayVideo(views, block_frames)
    Ly, Lx = video.Ly, video.Lx
    Lyb, Lxb, cols = binned_geometry(Ly, Lx, sbin)
    LY, LX, sy, sx = canvas_layout(Lyb, Lxb)
    rois = copy.deepcopy(rois)
    nroi = 0
    if rois is not None:
        for r in rois:
            if r["rind"] == 1:
                r["yrange_bin"], r["xrange_bin"] = roi_bin_ranges(r, sbin)
                nroi += 1
    avgframe, avgmotion = stage1_means(video, sbin)
    proc = {
        "Ly": Ly, "Lx": Lx, "sbin": sbin, "fullSVD": fullSVD, "Lybin": Lyb, "Lxbin": Lxb,
        "sybin": sy, "sxbin": sx, "LYbin": LY, "LXbin": LX,
        "avgframe": avgframe, "avgmotion": avgmotion,
        "avgframe_reshape": np.squeeze(scatter_canvas(np.hstack(avgframe)[:, None], LY, LX, sy, sx, Lyb, Lxb, cols)),
        "avgmotion_reshape": np.squeeze(scatter_canvas(np.hstack(avgmotion)[:, None], LY, LX, sy, sx, Lyb, Lxb, cols)),
        "rois": rois,
    }
    if "svd" not in stages:
        return proc
    if fullSVD or nroi > 0:
        U_mot, U_mov, S_mot, S_mov = stage2_masks(video, sbin, avgframe, avgmotion, motSVD, movSVD,
                                                  rois, fullSVD, NCOMPS, svd)
        U_mot_r, U_mov_r = list(U_mot), list(U_mov)
        if fullSVD:
            if motSVD:
                U_mot_r[0] = scatter_canvas(U_mot[0], LY, LX, sy, sx, Lyb, Lxb, cols)
            if movSVD:
                U_mov_r[0] = scatter_canvas(U_mov[0], LY, LX, sy, sx, Lyb, Lxb, cols)
        k = 1
        for r in (rois or []):
            if r["rind"] == 1:
                ly, lx = r["yrange_bin"].size, r["xrange_bin"].size
                if motSVD:
                    U_mot_r[k] = U_mot[k].copy().reshape(ly, lx, -1)
                if movSVD:
                    U_mov_r[k] = U_mov[k].copy().reshape(ly, lx, -1)
                k += 1
    else:
        U_mot, U_mov, S_mot, S_mov = [], [], [], []
        U_mot_r, U_mov_r = [], []
    proc.update({"motMask": U_mot, "movMask": U_mov, "motSv": S_mot, "movSv": S_mov,
                 "motMask_reshape": U_mot_r, "movMask_reshape": U_mov_r})
    if "proj" not in stages:
        return proc
    from .roi_oracle import FullResRois

    fullres = FullResRois(rois, video.N)
    V_mot, V_mov, M = stage3_project(video, sbin, avgframe, avgmotion, U_mot, U_mov, motSVD,
                                     movSVD, rois, fullSVD, fullres)
    pups, blinks, runs = fullres.finish()
    proc.update({"motion": M, "motSVD": V_mot, "movSVD": V_mov,
                 "pupil": pups, "blink": blinks, "running": runs})
    return proc
