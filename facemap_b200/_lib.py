"""ctypes binding of the C-ABI in include/facemap_b200.h (the only route from Python to the kernels).

There is no fallback: if the shared library is missing it is built in-tree with nvcc, and if that
fails or no sm_100 device is present the import / first call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_PKG, "lib", "libfacemap_b200.so")
FM_MAX_ROIS = 8


class FacemapB200Error(RuntimeError):
    pass


class Rois(C.Structure):
    """mirror of `fm_rois` (binned-coordinate rectangles of one view)."""
    _fields_ = [("n", C.c_int32),
                ("y0", C.c_int32 * FM_MAX_ROIS), ("y1", C.c_int32 * FM_MAX_ROIS),
                ("x0", C.c_int32 * FM_MAX_ROIS), ("x1", C.c_int32 * FM_MAX_ROIS)]

    @classmethod
    def from_rects(cls, rects):
        if len(rects) > FM_MAX_ROIS:
            raise FacemapB200Error(f"at most {FM_MAX_ROIS} motion ROIs per view are supported, got {len(rects)}")
        r = cls()
        r.n = len(rects)
        for i, (y0, y1, x0, x1) in enumerate(rects):
            r.y0[i], r.y1[i], r.x0[i], r.x1[i] = int(y0), int(y1), int(x0), int(x1)
        return r


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64

# name -> (restype, argtypes); must list every symbol declared in include/facemap_b200.h
SIGNATURES = {
    "fm_version": (_I, []),
    "fm_last_error": (C.c_char_p, []),
    "fm_check_device": (_I, [_I]),
    "fm_launch_count": (C.c_uint64, []),
    "fm_host_register": (_I, [_P, C.c_size_t]),
    "fm_host_unregister": (_I, [_P]),
    "fm_bin_motion": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _L, _P, C.POINTER(Rois), _P, _P, _P, _P, _P]),
    "fm_bin_motion_painted": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _L, _P, C.POINTER(Rois), _P, _P, _P, _P,
                                   _P, _P]),
    "fm_pairwise_plan": (_I, [_I, _I, _P, _P, _P]),
    "fm_motion_ref_scratch_bytes": (_I, [_I, _I, _I, _I, _I, C.POINTER(C.c_size_t)]),
    "fm_motion_ref": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P]),
    "fm_motion_ref_planes": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _I, _P, _P, _L, _L, _P, _P]),
    "fm_mean_ref": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "fm_mean_update": (_I, [_P, _P, _L, _I, _I, _I, _P]),
    "fm_gemm_tn": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _L, _P, _P, _I, _P, _L, _I, _I, _P]),
    "fm_gemm_i8_tn": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _L, _I, _P, _L, _I, _P]),
    "fm_project_i8": (_I, [_P, _P, _L, _P, _P, _P, _L, _P, _I, _I, _I, C.c_double, _P, _P, _L, _I, _P]),
    "fm_project_i8_blocked": (_I, [_P, _P, _L, _P, _P, _P, _L, _L, _P, _I, _I, _I, C.c_double, _P, _P, _L, _I, _P]),
    "fm_retile_u8": (_I, [_P, _L, _I, _L, _P, _L, _L, _P]),
    "fm_bin_motion_planes": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _L, _L, _P, C.POINTER(Rois), _P, _P, _P, _P]),
    "fm_slice_rows_i8": (_I, [_P, _L, _I, _L, _P, _P, _P, _L, _P, _P]),
    "fm_gram_assemble_i8": (_I, [_P, _I, _I, _L, _P, _P, _L, _P, _P]),
    "fm_gram_accumulate_i8": (_I, [_P, _I, _I, _L, _P, _P, _P]),
    "fm_gram_assemble_i8_rowside": (_I, [_P, _I, _I, _L, _P, _P, _P, _I, _P, _P]),
    "fm_split_lo": (_I, [_P, _L, _I, _I, _P, _L, _P]),
    "fm_slice_rows": (_I, [_P, _L, _I, _L, _P, _P, _P, _L, _P]),
    "fm_splitk_reduce": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _L, _P]),
    "fm_gemm_workspace_bytes": (_L, [_I, _L, _I]),
    "fm_row_means": (_I, [_P, _L, _I, _L, _P, _P]),
    "fm_gram_assemble": (_I, [_P, _I, _I, _L, _P, _L, _I, _P, _P]),
    "fm_row_dot": (_I, [_P, _L, _I, _L, _P, _P, _P]),
    "fm_gram_assemble_rowside": (_I, [_P, _I, _I, _L, _P, _P, _I, _I, _P, _P]),
    "fm_sign_from_right": (_I, [_P, _L, _I, _I, _P, _P, _L, _I, _P]),
    "fm_solver_create": (_I, [C.POINTER(_P)]),
    "fm_solver_destroy": (_I, [_P]),
    "fm_solver_check": (_I, [_P, _P]),
    "fm_eigh_topk_scratch_bytes": (C.c_size_t, [_I, _I, _I]),
    "fm_eigh_topk_batched": (_I, [_P, _I, _I, _I, _P, _P, _P, C.c_size_t, _P, _P]),
    "fm_syevd": (_I, [_P, _P, _I, _P, _I, _P]),
    "fm_syevd_batched": (_I, [_P, _P, _I, _I, _P, _P]),
    "fm_syevdx_top": (_I, [_P, _P, _I, _I, _P, _P]),
    "fm_topk_components": (_I, [_P, _P, _I, _I, _I, _P, _P, _L, _P, _P, _P, _P]),
    "fm_transpose": (_I, [_P, _L, _I, _I, _P, _L, _P]),
    "fm_gather_roi": (_I, [_P, _L, _I, _L, _I, _I, _I, _I, _I, _P, _L, _P]),
    "fm_gather_roi_u8": (_I, [_P, _L, _I, _L, _I, _I, _I, _I, _I, _P, _L, _P]),
    "fm_paint": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "fm_blink": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fm_running": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "fm_pupil_scratch_bytes": (_I, [_I, _I, _I, C.POINTER(C.c_size_t)]),
    "fm_pupil": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, C.c_float, C.c_double, _P, _I, _P, _P]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building in-tree if needed) libfacemap_b200.so and declare every signature."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIBPATH):
            if not build_if_missing:
                raise FacemapB200Error(f"{LIBPATH} is missing; run `python -m facemap_b200.build`")
            from . import build as _build
            _build.build()
        lib = C.CDLL(LIBPATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here = header/library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().fm_last_error().decode("utf-8", "replace")
        raise FacemapB200Error(f"{what or 'facemap_b200'} failed (code {rc}): {msg}")


_checked_devices = set()


def require_device(dev: int = 0) -> None:
    """Raise unless `dev` is a Blackwell part this library carries code for (no CPU fallback).
    The answer is cached per device: fm_check_device calls cudaGetDeviceProperties, which takes a driver-wide
    lock (tens of ms when something else, e.g. an NVML clock query, holds it) — not something to repeat per call."""
    dev = int(dev)
    if dev in _checked_devices:
        return
    check(load().fm_check_device(dev), "fm_check_device")
    _checked_devices.add(dev)


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
