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
    # fewer binned pixels (384) than chunk components (500) in the single-chunk regime: the reference leaves the
    # unfilled columns of its preallocated mask buffer zero (process.py:146-151, 246-259; never cut: :264-267)
    "tiny_onechunk_s2": ([(32, 48, 5)], 1200, 2, True, True, None),
}


# BASELINE configs[0] at full size and configs[1] at SURVEY §8(d)'s 20 000-frame parity size: fixtures made by
# `python scripts/make_golden.py c1_s4 c2_20k_s4`; not part of the default parametrisations (minutes of oracle time)
LARGE_CASES = {
    "c1_s4": ([(256, 256, 0)], 2000, 4, True, False, None),
    "c2_20k_s4": ([(480, 640, 0)], 20000, 4, True, False, None),
    # configs[2] at its 4 000-frame parity size (two cameras, motion + movie SVD, SURVEY §8(d)'s three motion ROIs)
    "c3_4k_s4": ([(480, 640, 0), (480, 640, 1)], 4000, 4, True, True,
                 [motion_roi(0, 40, 200, 80, 320), motion_roi(1, 32, 160, 64, 300), motion_roi(0, 240, 440, 400, 600)]),
    # configs[3] on a 256-pixel-wide crop (sbin 1: the reference's float32 path, 262 144 pixels)
    "c4_crop_s1": ([(1024, 256, 0)], 2000, 1, True, False, None),
}


def case_videos(name):
    views, N = {**CASES, **LARGE_CASES}[name][0], {**CASES, **LARGE_CASES}[name][1]
    return [SynthVideo(H, W, N, seed=seed).all_frames() for (H, W, seed) in views]


def case_rois(name):
    import copy

    return copy.deepcopy({**CASES, **LARGE_CASES}[name][5])


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


def component_errors(U, U_ref, V, V_ref, S=None, S_ref=None):
    """EVERY component of one target against the reference's: U / U_ref masks [p, k] (orthonormal columns up to
    float32 rounding, possibly followed by all-zero ones), V / V_ref traces [frames, k].  Returns
      S_rel[j]          |S_j - Sref_j| / Sref_j                                 (None without S)
      prefix_sin[j]     sine of the largest principal angle between span(U[:, :j+1]) and span(U_ref[:, :j+1])
      one_minus_cos[j]  1 - |<u_j, uref_j>| / (|u_j| |uref_j|)
      trace_rel[j]      ||sg_j V_j - Vref_j|| / ||Vref_j||, sg_j the sign that aligns mask j
      k                 number of (non-zero) components compared
    The prefixes are orthonormalised through the Cholesky factors of the two Gram matrices, whose leading blocks
    are the factors of the leading blocks: one factorisation serves all k prefixes."""
    U64, Ur64 = np.asarray(U, np.float64), np.asarray(U_ref, np.float64)
    k = min(U64.shape[1], Ur64.shape[1])
    nz = np.flatnonzero(np.linalg.norm(Ur64[:, :k], axis=0) > 0)
    k = int(nz[-1]) + 1 if nz.size else 0
    U64, Ur64 = U64[:, :k], Ur64[:, :k]
    C = Ur64.T @ U64
    nu, nr = np.linalg.norm(U64, axis=0), np.linalg.norm(Ur64, axis=0)
    d = np.diag(C) / np.where(nu * nr > 0, nu * nr, 1.0)
    sg = np.where(d < 0, -1.0, 1.0)
    import scipy.linalg

    L = np.linalg.cholesky(U64.T @ U64)
    Lr = np.linalg.cholesky(Ur64.T @ Ur64)
    M = scipy.linalg.solve_triangular(Lr, C, lower=True)                       # Lr^-1 C
    M = scipy.linalg.solve_triangular(L, M.T, lower=True).T                    # Lr^-1 C L^-T
    prefix = np.zeros(k)
    for j in range(k):
        smin = np.linalg.svd(M[:j + 1, :j + 1], compute_uv=False)[-1]
        prefix[j] = np.sqrt(max(0.0, 1.0 - min(1.0, smin) ** 2))
    Va = np.asarray(V, np.float64)[:, :k] * sg[None, :]
    Vr = np.asarray(V_ref, np.float64)[:, :k]
    trace = np.linalg.norm(Va - Vr, axis=0) / (np.linalg.norm(Vr, axis=0) + 1e-300)
    out = {"k": k, "prefix_sin": prefix, "one_minus_cos": 1.0 - np.abs(d), "trace_rel": trace, "S_rel": None}
    if S is not None and S_ref is not None:
        Sr = np.asarray(S_ref, np.float64)[:k]
        out["S_rel"] = np.abs(np.asarray(S, np.float64)[:k] - Sr) / np.where(Sr > 0, Sr, 1.0)
    return out


# What "gap-resolved" means (north_star: components / subspaces must match "after sign alignment"; a component whose
# singular value is as good as tied with a neighbour's is not determined by the data): the chunk components U S the
# reference itself stores are float32 (facemap/process.py:146-151, 229-259), a relative perturbation of 2^-24 ~ 6e-8
# that moves a singular subspace by about  STORE_EPS * S_0 / gap.  A component (or a prefix of components) is held to
# the tolerance wherever that floor is below a quarter of it.  Measured on B200 (profiles/r02_accuracy.md): the
# trace error of every component stays below 3.3e-7 * S_0 / gap (motion) and 1.5e-6 * S_0 / gap (movie).
STORE_EPS = 1e-7


def gap_resolved(S_ref, tol):
    """(component mask, prefix mask) over the singular values S_ref: component j resolved if both gaps to its
    neighbours exceed 4 * STORE_EPS * S_0 / tol; prefix j (components 0..j) if the gap S_j - S_{j+1} does."""
    S = np.asarray(S_ref, np.float64)
    if S.size == 0 or S[0] <= 0:
        return np.zeros(S.size, bool), np.zeros(S.size, bool)
    need = 4.0 * STORE_EPS * S[0] / tol
    down = np.r_[S[:-1] - S[1:], S[-1]]            # gap to the next one (the last keeps its own size: the cut-off)
    up = np.r_[np.inf, S[:-1] - S[1:]]
    return (np.minimum(up, down) >= need) & (S > 0), (down >= need) & (S > 0)


def sign_align(A, B):
    """flip columns of A to match the sign of B's columns (largest-|.| entry rule may pick a
    different pixel when two entries are nearly tied)."""
    sg = np.sign(np.sum(np.asarray(A, np.float64) * np.asarray(B, np.float64), axis=0))
    sg[sg == 0] = 1.0
    return A * sg[None, :], sg


# ---------------------------------------------------------------------------------------------
# full-resolution ROI cases (pupil / blink / running) — shared with scripts/make_golden_fullres.py
# ---------------------------------------------------------------------------------------------
def ellipse_mask(ny, nx):
    """Boolean ellipse inscribed in an ny x nx rectangle (what facemap/roi.py:254-259 draws)."""
    y, x = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    return ((y - ny / 2) ** 2 / (ny / 2) ** 2 + (x - nx / 2) ** 2 / (nx / 2) ** 2) <= 1


def eye_video(H, W, N, seed=0, wheel=True):
    """uint8 (N, H, W): bright noisy background, a dark disc that drifts and pulses (pupil) with a
    small bright reflection on it, slow lid closures (blinks), and a drifting texture band at the
    bottom (running wheel).  Deterministic for a seed."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    t = np.arange(N)
    cy = H * 0.35 + 3.0 * np.sin(t / 37.0) + rs.randn(N).cumsum() * 0.05
    cx = W * 0.40 + 4.0 * np.cos(t / 53.0) + rs.randn(N).cumsum() * 0.05
    rad = min(H, W) * (0.10 + 0.03 * np.sin(t / 91.0))
    lid = np.clip(np.sin(t / 140.0) * 4 - 3, 0, 1)                       # 0 open .. 1 closed
    tex = rs.randint(0, 256, size=(H, 4 * W)).astype(np.int32)
    shift = np.cumsum(rs.randint(0, 4, size=N))
    out = np.empty((N, H, W), np.uint8)
    for i in range(N):
        f = 170 + 25 * np.sin(yy / 9.0 + i / 300.0) + rs.randint(-6, 7, size=(H, W))
        r2 = (yy - cy[i]) ** 2 + (xx - cx[i]) ** 2
        f = np.where(r2 <= rad[i] ** 2, 25 + rs.randint(0, 12, size=(H, W)), f)
        f = np.where((yy - cy[i] + 2) ** 2 + (xx - cx[i] - 3) ** 2 <= 4, 250, f)
        f = f * (1 - 0.6 * lid[i] * (yy < H * 0.6))
        if wheel:
            band = yy >= int(H * 0.7)
            s = int(shift[i]) % (3 * W)
            f = np.where(band, tex[:, s:s + W] // 2 + 60, f)
        out[i] = np.clip(f, 0, 255).astype(np.uint8)
    if N > 600:      # degenerate frames: nothing dark, everything dark, one dark pixel, two far dark pixels
        out[505] = 255
        out[506] = 0
        out[507] = 255
        out[507, int(H * 0.3), int(W * 0.4)] = 3
        out[508] = 255
        out[508, int(H * 0.2), int(W * 0.3)] = 0
        out[508, int(H * 0.4), int(W * 0.5)] = 10
    return out


def pupil_roi(ivid, y0, y1, x0, x1, saturation, sigma=2.0, reflector=None):
    r = {"rind": 0, "rtype": "Pupil", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
         "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": saturation,
         "pupil_sigma": sigma, "ellipse": ellipse_mask(y1 - y0, x1 - x0)}
    if reflector is not None:
        r["reflector"] = reflector
    return r


def reflector_roi(y0, y1, x0, x1):
    """Reflector rectangle in pupil-crop coordinates (facemap/utils.py:656-661)."""
    return {"yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "ellipse": ellipse_mask(y1 - y0, x1 - x0)}


def blink_roi(ivid, y0, y1, x0, x1, saturation):
    return {"rind": 2, "rtype": "blink", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
            "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": saturation,
            "ellipse": ellipse_mask(y1 - y0, x1 - x0)}


def running_roi(ivid, y0, y1, x0, x1):
    return {"rind": 3, "rtype": "running", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
            "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": 0}


def fullres_case(name):
    """-> (views, sbin, motSVD, movSVD, rois, fullSVD)"""
    if name == "eye_all":            # every processor on one view, overlapping rectangles (paint order)
        v = [eye_video(72, 96, 1150, seed=1)]
        rois = [pupil_roi(0, 8, 44, 16, 64, 90.27), blink_roi(0, 4, 40, 40, 90, 140.25),
                running_roi(0, 51, 72, 2, 94), motion_roi(0, 6, 60, 10, 80),
                pupil_roi(0, 10, 40, 20, 60, 60.0, sigma=2.5)]
        return v, 2, True, False, rois, True
    if name == "eye_reflector":      # reflector fill-in, int saturation for blink (uint8 wrap), two views
        v = [eye_video(64, 80, 700, seed=2, wheel=False), eye_video(48, 56, 700, seed=3)]
        rois = [pupil_roi(0, 6, 42, 10, 58, 80.325, reflector=[reflector_roi(10, 18, 18, 27), reflector_roi(20, 24, 30, 36)]),
                blink_roi(0, 2, 38, 6, 70, 200), running_roi(1, 34, 48, 0, 56), blink_roi(1, 0, 30, 3, 50, 120.7)]
        return v, 1, True, True, rois, True
    if name == "eye_only":           # no SVD at all: fullSVD off and no motion ROI
        v = [eye_video(60, 70, 620, seed=4)]
        rois = [pupil_roi(0, 5, 40, 8, 52, 100.47), running_roi(0, 42, 60, 0, 70)]
        return v, 2, True, False, rois, False
    raise KeyError(name)


FULLRES_CASES = ("eye_all", "eye_reflector", "eye_only")
