"""TEST INFRASTRUCTURE — runs the *unmodified* reference (`/root/reference/facemap/process.py`)
in-process on in-memory uint8 videos.  Only usable in the build container (the GPU box has no
/root/reference); used by `scripts/make_golden.py` to produce `tests/golden/*.npz` and by
opt-in CPU tests.  Nothing in the product imports this file.

How the reference is fed without a codec:
  * `facemap/utils.py:5` imports h5py (absent here, unused on this path) -> empty stub module.
  * `utils.get_frame_details` (`utils.py:567-591`) is replaced by one returning duck-typed
    captures over NumPy arrays (`.get/.set/.read/.release`, the only calls made by
    `utils.get_frames`, `utils.py:522-553`, and `utils.close_videos`, `:556-564`).
    Frames are handed out as 3-channel B=G=R so `cv2.cvtColor(BGR2GRAY)` is the identity.
  * Two oracles: O1 = stock; O2 = only `utils.svdecon` (`utils.py:751-776`)
    swapped for an exact float64 LAPACK SVD of the column-centred matrix + the same sign rule.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "facemap"))


def _import_reference():
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules["h5py"] = types.ModuleType("h5py")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from facemap import process, utils  # type: ignore

    return process, utils


class ArrayCapture:
    """Duck-typed cv2.VideoCapture over a (T, H, W) uint8 array."""

    def __init__(self, arr: np.ndarray):
        assert arr.dtype == np.uint8 and arr.ndim == 3
        self.arr, self.pos = arr, 0

    def get(self, prop):
        import cv2

        if prop == cv2.CAP_PROP_POS_FRAMES:
            return float(self.pos)
        if prop == cv2.CAP_PROP_FRAME_COUNT:
            return float(self.arr.shape[0])
        if prop == cv2.CAP_PROP_FRAME_HEIGHT:
            return float(self.arr.shape[1])
        if prop == cv2.CAP_PROP_FRAME_WIDTH:
            return float(self.arr.shape[2])
        return 0.0

    def set(self, prop, value):
        import cv2

        if prop == cv2.CAP_PROP_POS_FRAMES:
            self.pos = int(value)
        return True

    def read(self):
        if self.pos >= self.arr.shape[0]:
            return False, None
        f = self.arr[self.pos]
        self.pos += 1
        return True, np.repeat(f[:, :, None], 3, axis=2)

    def release(self):
        pass


def exact_svdecon(X, k=100):
    """O2 replacement for `utils.svdecon`: exact float64 SVD of the column-centred matrix and
    sklearn's `svd_flip(u_based_decision=False)` sign rule (`sklearn/utils/extmath.py:924-982`)."""
    import scipy.linalg

    X64 = np.asarray(X, dtype=np.float64)
    Xc = X64 - X64.mean(axis=0)
    U, S, Vt = scipy.linalg.svd(Xc, full_matrices=False, lapack_driver="gesdd")
    U, S, Vt = U[:, :k], S[:k], Vt[:k]
    idx = np.argmax(np.abs(Vt), axis=1)
    sg = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    sg[sg == 0] = 1.0
    U = U * sg[None, :]
    Vt = Vt * sg[:, None]
    dt = X.dtype if X.dtype in (np.float32, np.float64) else np.float64
    return U.astype(dt), S.astype(dt), Vt.astype(dt)


def run_reference(videos, sbin=1, motSVD=True, movSVD=False, rois=None, fullSVD=True,
                  exact=False, quiet=True):
    """Run reference `process.run` on `videos` = 2-D list [[view0_arr, view1_arr, ...] per block].

    Returns the proc dict loaded back from the `_proc.npy` the reference wrote."""
    import contextlib
    import io

    process, utils = _import_reference()
    nblocks = len(videos)
    filenames = [[f"mem_b{b}_v{v}.avi" for v in range(len(videos[b]))] for b in range(nblocks)]

    def fake_details(fnames):
        cum, containers = [0], []
        for b in range(nblocks):
            Ly, Lx, cs = [], [], []
            for arr in videos[b]:
                cs.append(ArrayCapture(arr))
                Ly.append(int(arr.shape[1]))
                Lx.append(int(arr.shape[2]))
            containers.append(cs)
            cum.append(cum[-1] + int(videos[b][0].shape[0]))
        return np.array(cum).astype(int), Ly, Lx, containers

    orig_details, orig_svdecon = utils.get_frame_details, utils.svdecon
    utils.get_frame_details = fake_details
    if exact:
        utils.svdecon = exact_svdecon
    tmp = tempfile.mkdtemp(prefix="refrun_")
    try:
        proc_in = None
        if rois is not None or not fullSVD:
            proc_in = {"sbin": sbin, "fullSVD": fullSVD, "save_mat": False, "rois": rois,
                       "sy": 0, "sx": 0, "savepath": tmp}
        sink = io.StringIO()
        ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
        with ctx:
            savename = process.run(filenames, sbin=sbin, motSVD=motSVD, movSVD=movSVD,
                                   proc=proc_in, savepath=tmp)
        out = np.load(savename, allow_pickle=True).item()
    finally:
        utils.get_frame_details, utils.svdecon = orig_details, orig_svdecon
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    return out
