"""Thin Python wrappers over the C-ABI (include/facemap_b200.h): torch tensors in, raw pointers and
the current CUDA stream out.  torch is used for device memory and streams only.

Reference seams (facemap = /root/reference/facemap):
  bin_motion      spatial_bin + |diff| + avg subtract + energy    process.py:34-43, 201-212, 448-463
  gemm_tn         `X @ U` and the products inside utils.svdecon     process.py:485-510, utils.py:751-776
  Solver.syevd    the LAPACK work inside svdecon                     utils.py:773-775
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import FacemapB200Error, Rois, check, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name, dims=None):
    if t is None:
        return
    if not t.is_cuda:
        raise FacemapB200Error(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise FacemapB200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise FacemapB200Error(f"{name}: expected {dims} dims, got {t.dim()}")


def round_up(x: int, m: int) -> int:
    return (int(x) + m - 1) // m * m


def alloc_matrix(rows: int, cols: int, device, zero: bool = False) -> torch.Tensor:
    """float32 [rows, cols] view whose row pitch is a multiple of 32 elements (128 B): a legal TMA
    operand for fm_gemm_tn whatever `cols` is."""
    ld = max(32, round_up(cols, 32))
    buf = (torch.zeros if zero else torch.empty)((max(rows, 1), ld), dtype=torch.float32, device=device)
    return buf[:rows, :cols]


def launch_count() -> int:
    return int(_lib.load().fm_launch_count())


def bin_motion(frames, sbin, *, prev_sum=None, avgmotion=None, avgframe=None, x_mot=None, x_mov=None,
               col0=0, energy=None, rois=None, roi_energy=None, last_sum=None, tsum_bin=None,
               tsum_adiff=None, ormask=None):
    """K1 on one view.  frames: uint8 [T,H,W] contiguous.  See fm_bin_motion in the header.
    ormask: optional uint8 [H, W] of 0x00 / 0xFF OR-ed into the frames on load (fm_bin_motion_painted)."""
    _need(frames, torch.uint8, "frames", 3)
    if not frames.is_contiguous():
        raise FacemapB200Error("frames must be contiguous [T, H, W]")
    T, H, W = frames.shape
    _need(prev_sum, torch.int32, "prev_sum")
    _need(avgmotion, torch.float32, "avgmotion")
    _need(avgframe, torch.float32, "avgframe")
    _need(x_mot, torch.float32, "x_mot", 2)
    _need(x_mov, torch.float32, "x_mov", 2)
    _need(energy, torch.int64, "energy")
    _need(roi_energy, torch.int64, "roi_energy")
    _need(last_sum, torch.int32, "last_sum")
    _need(tsum_bin, torch.int64, "tsum_bin")
    _need(tsum_adiff, torch.int64, "tsum_adiff")
    npix = (H // sbin) * (W // sbin)
    rows = T if prev_sum is not None else T - 1
    ldx = 0
    for x in (x_mot, x_mov):
        if x is not None:
            if x.stride(1) != 1 or x.shape[0] < rows or col0 + npix > x.shape[1]:
                raise FacemapB200Error(f"x operand too small or not row-major: {tuple(x.shape)}, rows {rows}")
            if ldx and x.stride(0) != ldx:
                raise FacemapB200Error("x_mot and x_mov must share a row pitch")
            ldx = x.stride(0)
    for v, n, name in ((avgmotion, npix, "avgmotion"), (avgframe, npix, "avgframe"), (prev_sum, npix, "prev_sum"),
                       (last_sum, npix, "last_sum"), (tsum_bin, npix, "tsum_bin"), (tsum_adiff, npix, "tsum_adiff"),
                       (energy, rows, "energy")):
        if v is not None and (v.numel() < n or not v.is_contiguous()):
            raise FacemapB200Error(f"{name}: needs {n} contiguous elements, has {v.numel()}")
    r = None
    if roi_energy is not None:
        r = rois if isinstance(rois, Rois) else Rois.from_rects(rois or [])
        if roi_energy.numel() < r.n * rows:
            raise FacemapB200Error("roi_energy too small")
    lib = _lib.load()
    if ormask is not None:
        _need(ormask, torch.uint8, "ormask", 2)
        if tuple(ormask.shape) != (H, W) or not ormask.is_contiguous():
            raise FacemapB200Error(f"ormask: expected a contiguous [{H}, {W}] uint8 tensor")
        check(lib.fm_bin_motion_painted(ptr(frames), T, H, W, int(sbin), ptr(prev_sum), ptr(avgmotion), ptr(avgframe),
                                        ptr(x_mot), ptr(x_mov), ldx, int(col0), ptr(energy),
                                        C.byref(r) if r is not None else None, ptr(roi_energy), ptr(last_sum),
                                        ptr(tsum_bin), ptr(tsum_adiff), ptr(ormask), _stream()), "fm_bin_motion_painted")
        return rows
    check(lib.fm_bin_motion(ptr(frames), T, H, W, int(sbin), ptr(prev_sum), ptr(avgmotion), ptr(avgframe),
                            ptr(x_mot), ptr(x_mov), ldx, int(col0), ptr(energy),
                            C.byref(r) if r is not None else None, ptr(roi_energy), ptr(last_sum),
                            ptr(tsum_bin), ptr(tsum_adiff), _stream()), "fm_bin_motion")
    return rows


def mean_update(acc, tsum, sbin, count, finalize_div=0):
    _need(acc, torch.float32, "acc")
    _need(tsum, torch.int64, "tsum")
    check(_lib.load().fm_mean_update(ptr(acc), ptr(tsum), acc.numel(), int(sbin), int(count), int(finalize_div),
                                     _stream()), "fm_mean_update")


def split_lo(X, out=None):
    """lo plane of the 3xTF32 split of X (same shape and pitch policy as alloc_matrix)."""
    _need(X, torch.float32, "X", 2)
    lo = alloc_matrix(X.shape[0], X.shape[1], X.device) if out is None else out
    if tuple(lo.shape) != tuple(X.shape) or lo.stride(1) != 1:
        raise FacemapB200Error("split_lo: out must have X's shape")
    check(_lib.load().fm_split_lo(ptr(X), X.stride(0), X.shape[0], X.shape[1], ptr(lo), lo.stride(0), _stream()),
          "fm_split_lo")
    return lo


def slice_rows(X, out=None):
    """Three 8-bit row-scaled slices (S0, S1, S2) of X, S0 + S1 + S2 == X up to 2^-24 of each row's largest
    magnitude (fm_slice_rows).  `out`: optional triple of [rows, cols] matrices sharing one row pitch."""
    _need(X, torch.float32, "X", 2)
    rows, cols = X.shape
    planes = tuple(alloc_matrix(rows, cols, X.device) for _ in range(3)) if out is None else tuple(out)
    lds = planes[0].stride(0)
    for pl in planes:
        _need(pl, torch.float32, "slice", 2)
        if tuple(pl.shape) != (rows, cols) or pl.stride(1) != 1 or pl.stride(0) != lds:
            raise FacemapB200Error("slice_rows: planes must have X's shape and one common row pitch")
    if X.stride(1) != 1:
        raise FacemapB200Error("slice_rows: X must be row-major")
    check(_lib.load().fm_slice_rows(ptr(X), X.stride(0), rows, cols, ptr(planes[0]), ptr(planes[1]), ptr(planes[2]),
                                    lds, _stream()), "fm_slice_rows")
    return planes


def gemm_tn(A, B, C_out=None, *, rscale=None, rbias=None, ksplit=1, workspace=None, upper_only=False, impl=0,
            B_lo=None, A_lo=None):
    """C[M,N] = rscale[m] * (A[M,K] @ B[N,K]^T) + rbias[m]; both operands K-contiguous float32.
    With ksplit > 1 the raw partial products go to `workspace` [ksplit, M, ldw] instead."""
    _need(A, torch.float32, "A", 2)
    _need(B, torch.float32, "B", 2)
    M, K = A.shape
    N, K2 = B.shape
    if K != K2 or A.stride(1) != 1 or B.stride(1) != 1:
        raise FacemapB200Error(f"gemm_tn: operands must be [M,K] and [N,K] row-major, got {tuple(A.shape)} {tuple(B.shape)}")
    ldc = ldw = 0
    if ksplit > 1:
        _need(workspace, torch.float32, "workspace", 3)
        if workspace.shape[0] < ksplit or workspace.shape[1] != M or workspace.shape[2] < N or workspace.stride(2) != 1 \
                or workspace.stride(0) != M * workspace.stride(1):
            raise FacemapB200Error(f"gemm_tn: workspace must be [ksplit, M, >=N] packed, got {tuple(workspace.shape)}")
        ldw = workspace.stride(1)
    else:
        _need(C_out, torch.float32, "C", 2)
        if C_out.shape[0] < M or C_out.shape[1] < N or C_out.stride(1) != 1:
            raise FacemapB200Error(f"gemm_tn: C must be row-major [>=M, >=N], got {tuple(C_out.shape)}")
        ldc = C_out.stride(0)
    _need(rscale, torch.float32, "rscale")
    _need(rbias, torch.float32, "rbias")
    _need(B_lo, torch.float32, "B_lo", 2)
    if B_lo is not None and (tuple(B_lo.shape) != tuple(B.shape) or B_lo.stride(1) != 1):
        raise FacemapB200Error("gemm_tn: B_lo must have B's shape")
    _need(A_lo, torch.float32, "A_lo", 2)
    if A_lo is not None and (tuple(A_lo.shape) != tuple(A.shape) or A_lo.stride(1) != 1 or B_lo is None):
        raise FacemapB200Error("gemm_tn: A_lo must have A's shape and needs B_lo")
    check(_lib.load().fm_gemm_tn(ptr(A), A.stride(0), ptr(A_lo), A_lo.stride(0) if A_lo is not None else 0,
                                 ptr(B), B.stride(0), ptr(B_lo),
                                 B_lo.stride(0) if B_lo is not None else 0, M, N, K, ptr(C_out), ldc, ptr(rscale),
                                 ptr(rbias), int(ksplit), ptr(workspace), ldw, int(bool(upper_only)), int(impl),
                                 _stream()), "fm_gemm_tn")


def gemm_i8_tn(A, B, C_out=None, *, ksplit=1, workspace=None, upper_only=False):
    """fm_gemm_i8_tn: int32 C[M,N] = A[M,K] @ B[N,K]^T for uint8 / int8 operands, exact.
    With ksplit > 1 the int32 partials go to `workspace` [ksplit, M, ldw]."""
    for t, name in ((A, "A"), (B, "B")):
        if t.dtype not in (torch.uint8, torch.int8) or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
            raise FacemapB200Error(f"gemm_i8_tn: {name} must be a row-major CUDA uint8 / int8 matrix")
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise FacemapB200Error("gemm_i8_tn: operands must be [M,K] and [N,K]")
    ldc = ldw = 0
    if ksplit > 1:
        _need(workspace, torch.int32, "workspace", 3)
        if workspace.shape[0] < ksplit or workspace.shape[1] != M or workspace.shape[2] < N or workspace.stride(2) != 1 \
                or workspace.stride(0) != M * workspace.stride(1):
            raise FacemapB200Error("gemm_i8_tn: workspace must be [ksplit, M, >=N] packed")
        ldw = workspace.stride(1)
    else:
        _need(C_out, torch.int32, "C", 2)
        if C_out.shape[0] < M or C_out.shape[1] < N or C_out.stride(1) != 1:
            raise FacemapB200Error("gemm_i8_tn: C must be row-major [>=M, >=N] int32")
        ldc = C_out.stride(0)
    check(_lib.load().fm_gemm_i8_tn(ptr(A), A.stride(0), int(A.dtype == torch.int8), ptr(B), B.stride(0),
                                    int(B.dtype == torch.int8), M, N, K, ptr(C_out), ldc, int(ksplit), ptr(workspace),
                                    ldw, int(bool(upper_only)), _stream()), "fm_gemm_i8_tn")


def slice_rows_i8(X, out=None):
    """fm_slice_rows_i8: (Q0 int8, Q1 uint8, Q2 uint8, exps int32) digit planes of the rows of X, row
    pitch a multiple of 16 bytes.  `out`: optional (Q0, Q1, Q2) [rows, cols] views sharing one pitch."""
    _need(X, torch.float32, "X", 2)
    rows, cols = X.shape
    if X.stride(1) != 1:
        raise FacemapB200Error("slice_rows_i8: X must be row-major")
    if out is None:
        ld = round_up(cols, 16)
        out = tuple(torch.empty((max(rows, 1), ld), dtype=dt, device=X.device)[:rows, :cols]
                    for dt in (torch.int8, torch.uint8, torch.uint8))
    Q0, Q1, Q2 = out
    ldq = Q0.stride(0)
    for q, dt in ((Q0, torch.int8), (Q1, torch.uint8), (Q2, torch.uint8)):
        if q.dtype != dt or tuple(q.shape) != (rows, cols) or q.stride(1) != 1 or q.stride(0) != ldq or ldq % 16:
            raise FacemapB200Error("slice_rows_i8: planes must be int8 / uint8 / uint8 [rows, cols] with one 16-byte pitch")
    exps = torch.empty(max(rows, 1), dtype=torch.int32, device=X.device)[:rows]
    check(_lib.load().fm_slice_rows_i8(ptr(X), X.stride(0), rows, cols, ptr(Q0), ptr(Q1), ptr(Q2), ldq, ptr(exps),
                                       _stream()), "fm_slice_rows_i8")
    return Q0, Q1, Q2, exps


def gram_accumulate_i8(P, ksplit, n, exps, G):
    """fm_gram_accumulate_i8: G += the weighted plane products of a further block of columns."""
    _need(P, torch.int32, "P", 4)
    _need(G, torch.float64, "G", 2)
    if P.shape[0] != 6 or P.shape[1] < ksplit or P.shape[2] != n or P.shape[3] < n or not P.is_contiguous():
        raise FacemapB200Error("gram_accumulate_i8: P must be a contiguous int32 [6, ksplit, n, >=n] stack")
    if tuple(G.shape) != (n, n) or not G.is_contiguous():
        raise FacemapB200Error("gram_accumulate_i8: G must be a contiguous [n, n] float64 matrix")
    check(_lib.load().fm_gram_accumulate_i8(ptr(P), int(ksplit), int(n), P.stride(2), ptr(exps), ptr(G), _stream()),
          "fm_gram_accumulate_i8")
    return G


def gram_assemble_i8(P, ksplit, n, exps, mean, ncols, out=None, rowside=None):
    """fm_gram_assemble_i8: P int32 [6, ksplit, n, ldp] plane products -> float64 centred Gram  X X^T - ncols m m^T.
    rowside = (q, mu): the row-side form of a wide matrix instead,  Y Y^T - q 1^T - 1 q^T + mu.mu
    (fm_gram_assemble_i8_rowside)."""
    _need(P, torch.int32, "P", 4)
    if P.shape[0] != 6 or P.shape[1] < ksplit or P.shape[2] != n or P.shape[3] < n or not P.is_contiguous():
        raise FacemapB200Error("gram_assemble_i8: P must be a contiguous int32 [6, ksplit, n, >=n] stack")
    _need(exps, torch.int32, "exps", 1)
    _need(mean, torch.float64, "mean")
    if out is None:
        out = torch.empty((n, n), dtype=torch.float64, device=P.device)
    if rowside is not None:
        q, mu = rowside
        _need(q, torch.float64, "q", 1)
        _need(mu, torch.float64, "mu", 1)
        if q.numel() < n:
            raise FacemapB200Error("gram_assemble_i8: q must hold one entry per row")
        check(_lib.load().fm_gram_assemble_i8_rowside(ptr(P), int(ksplit), int(n), P.stride(2), ptr(exps), ptr(q), ptr(mu),
                                                      mu.numel(), ptr(out), _stream()), "fm_gram_assemble_i8_rowside")
        return out
    check(_lib.load().fm_gram_assemble_i8(ptr(P), int(ksplit), int(n), P.stride(2), ptr(exps), ptr(mean), int(ncols),
                                          ptr(out), _stream()), "fm_gram_assemble_i8")
    return out


def bin_motion_planes(frames, sbin, *, prev_sum=None, q_mot=None, q_mov=None, col0=0, energy=None, rois=None,
                      roi_energy=None, last_sum=None, ormask=None):
    """fm_bin_motion_planes: K1 writing its integer results as byte planes.  q_mot / q_mov: (hi, lo)
    pairs of uint8 [rows, >= col0 + npix] matrices sharing one row pitch (hi may be None when sbin == 1); they receive
    |S_t - S_(t-1)| and S_t as 256 * hi + lo — the frame operands of `project_i8`."""
    _need(frames, torch.uint8, "frames", 3)
    if not frames.is_contiguous():
        raise FacemapB200Error("frames must be contiguous [T, H, W]")
    T, H, W = frames.shape
    npix = (H // sbin) * (W // sbin)
    rows = T if prev_sum is not None else T - 1
    _need(prev_sum, torch.int32, "prev_sum")
    _need(energy, torch.int64, "energy")
    _need(roi_energy, torch.int64, "roi_energy")
    _need(last_sum, torch.int32, "last_sum")
    if q_mot is None and q_mov is None:
        raise FacemapB200Error("bin_motion_planes: no planes to write")
    ldq, ptrs = 0, []
    for pair, name in ((q_mot, "q_mot"), (q_mov, "q_mov")):
        hi, lo = (None, None) if pair is None else pair
        if pair is not None and (lo is None or (hi is None and sbin > 1)):
            raise FacemapB200Error(f"bin_motion_planes: {name} needs (hi, lo) planes (hi optional only for sbin 1)")
        for q in (hi, lo):
            if q is not None:
                _need(q, torch.uint8, name, 2)
                if q.stride(1) != 1 or q.shape[0] < rows or col0 + npix > q.shape[1] or (ldq and q.stride(0) != ldq):
                    raise FacemapB200Error(f"bin_motion_planes: {name} planes must be row-major [>= {rows}, >= "
                                           f"{col0 + npix}] with one shared pitch")
                ldq = q.stride(0)
        ptrs += [ptr(hi), ptr(lo)]
    for v, n, name in ((prev_sum, npix, "prev_sum"), (last_sum, npix, "last_sum"), (energy, rows, "energy")):
        if v is not None and (v.numel() < n or not v.is_contiguous()):
            raise FacemapB200Error(f"{name}: needs {n} contiguous elements, has {v.numel()}")
    r = None
    if roi_energy is not None:
        r = rois if isinstance(rois, Rois) else Rois.from_rects(rois or [])
        if roi_energy.numel() < r.n * rows:
            raise FacemapB200Error("roi_energy too small")
    if ormask is not None:
        _need(ormask, torch.uint8, "ormask", 2)
        if tuple(ormask.shape) != (H, W) or not ormask.is_contiguous():
            raise FacemapB200Error(f"ormask: expected a contiguous [{H}, {W}] uint8 tensor")
    check(_lib.load().fm_bin_motion_planes(ptr(frames), T, H, W, int(sbin), ptr(prev_sum), *ptrs, ldq, int(col0),
                                           ptr(energy), C.byref(r) if r is not None else None, ptr(roi_energy),
                                           ptr(last_sum), ptr(ormask), _stream()), "fm_bin_motion_planes")
    return rows


def project_i8(D, Q, exps, alpha, bias, V, impl=0):
    """fm_project_i8: V[t, c] = alpha * sum_p D[t, p] U[c, p] - bias[c] on tcgen05 kind::i8, exactly.
    D: (hi, lo) uint8 [M, K] planes of one pitch (hi None for operands < 256); Q: the (Q0 int8, Q1, Q2 uint8) [N, K]
    digit planes and `exps` of `slice_rows_i8(U)`; bias: float64 [N] or None; V: float32 [M, >= N] row-major;
    impl: 0 = one CTA per tile, 1 = CTA pairs (cta_group::2)."""
    hi, lo = D
    Q0, Q1, Q2 = Q
    _need(lo, torch.uint8, "D_lo", 2)
    _need(hi, torch.uint8, "D_hi", 2)
    M, Kd = lo.shape
    N = Q0.shape[0]
    if lo.stride(1) != 1 or (hi is not None and (tuple(hi.shape) != (M, Kd) or hi.stride(1) != 1 or hi.stride(0) != lo.stride(0))):
        raise FacemapB200Error("project_i8: D planes must be row-major [M, K] with one shared pitch")
    for q, dt in ((Q0, torch.int8), (Q1, torch.uint8), (Q2, torch.uint8)):
        if q.dtype != dt or not q.is_cuda or tuple(q.shape) != (N, Kd) or q.stride(1) != 1 or q.stride(0) != Q0.stride(0):
            raise FacemapB200Error("project_i8: Q must be int8 / uint8 / uint8 [N, K] planes with one shared pitch")
    _need(exps, torch.int32, "exps", 1)
    _need(bias, torch.float64, "bias", 1)
    _need(V, torch.float32, "V", 2)
    if exps.numel() < N or (bias is not None and bias.numel() < N) or V.shape[0] < M or V.shape[1] < N or V.stride(1) != 1:
        raise FacemapB200Error("project_i8: exps / bias need N entries and V must be row-major [>= M, >= N]")
    check(_lib.load().fm_project_i8(ptr(hi), ptr(lo), lo.stride(0), ptr(Q0), ptr(Q1), ptr(Q2), Q0.stride(0), ptr(exps),
                                    M, N, Kd, float(alpha), ptr(bias), ptr(V), V.stride(0), int(impl), _stream()), "fm_project_i8")
    return V


def retile_u8(src, kb, out=None):
    """fm_retile_u8: a byte plane [rows, cols] (row-major, any pitch) -> its K-blocked copy, a contiguous
    [ceil(cols / kb), rows, kb] tensor (int8 in, int8 out), zero-filled behind `cols`."""
    if src.dtype not in (torch.uint8, torch.int8) or src.dim() != 2 or src.stride(1) != 1 or not src.is_cuda:
        raise FacemapB200Error("retile_u8: expected a row-major CUDA uint8 / int8 matrix")
    rows, cols = src.shape
    nblk = -(-cols // int(kb))
    if out is None:
        out = torch.empty((nblk, max(rows, 1), int(kb)), dtype=src.dtype, device=src.device)[:, :rows]
    if out.dtype != src.dtype or tuple(out.shape) != (nblk, rows, int(kb)) or out.stride(2) != 1 or out.stride(1) != int(kb):
        raise FacemapB200Error("retile_u8: out must be a [nblk, rows, kb] tensor with packed rows")
    for r0 in range(0, rows, 32768):
        r1 = min(rows, r0 + 32768)
        check(_lib.load().fm_retile_u8(ptr(src[r0:r1]), src.stride(0), r1 - r0, cols, ptr(out[:, r0:r1]), int(kb),
                                       out.stride(0), _stream()), "fm_retile_u8")
    return out


def project_i8_blocked(D, Q, exps, alpha, bias, V, K_total, impl=0):
    """fm_project_i8_blocked: project_i8 on K-blocked planes (retile_u8): D = (hi or None, lo) [nblk, M, kb],
    Q = (Q0 int8, Q1, Q2 uint8) [nblk, N, kb]; K_total = the number of valid columns."""
    hi, lo = D
    Q0, Q1, Q2 = Q
    nblk, M, kb = lo.shape
    N = Q0.shape[1]
    for t, dt, rows, name in ((lo, torch.uint8, M, "D_lo"), (hi, torch.uint8, M, "D_hi"), (Q0, torch.int8, N, "Q0"),
                              (Q1, torch.uint8, N, "Q1"), (Q2, torch.uint8, N, "Q2")):
        if t is None:
            continue
        if t.dtype != dt or not t.is_cuda or tuple(t.shape) != (nblk, rows, kb) or t.stride(2) != 1 or t.stride(1) != kb:
            raise FacemapB200Error(f"project_i8_blocked: {name} must be a {dt} [nblk, rows, kb] tensor with packed rows")
    if hi is not None and hi.stride(0) != lo.stride(0) or Q1.stride(0) != Q0.stride(0) or Q2.stride(0) != Q0.stride(0):
        raise FacemapB200Error("project_i8_blocked: the planes of an operand must share one block stride")
    if not (nblk - 1) * kb < int(K_total) <= nblk * kb:
        raise FacemapB200Error("project_i8_blocked: K_total does not match the number of blocks")
    _need(exps, torch.int32, "exps", 1)
    _need(bias, torch.float64, "bias", 1)
    _need(V, torch.float32, "V", 2)
    if exps.numel() < N or (bias is not None and bias.numel() < N) or V.shape[0] < M or V.shape[1] < N or V.stride(1) != 1:
        raise FacemapB200Error("project_i8_blocked: exps / bias need N entries and V must be row-major [>= M, >= N]")
    check(_lib.load().fm_project_i8_blocked(ptr(hi), ptr(lo), lo.stride(0), ptr(Q0), ptr(Q1), ptr(Q2), Q0.stride(0), int(kb),
                                            ptr(exps), M, N, int(K_total), float(alpha), ptr(bias), ptr(V), V.stride(0),
                                            int(impl), _stream()), "fm_project_i8_blocked")
    return V


def splitk_reduce(workspace, ksplit, C_out, *, rscale=None, rbias=None):
    """C[m,n] = rscale[m] * sum_s workspace[s,m,n] + rbias[m], summed in float64."""
    _need(workspace, torch.float32, "workspace", 3)
    _need(C_out, torch.float32, "C", 2)
    M = workspace.shape[1]
    N = min(workspace.shape[2], C_out.shape[1])
    if C_out.shape[0] < M or C_out.stride(1) != 1 or workspace.shape[0] < ksplit:
        raise FacemapB200Error("splitk_reduce: shape mismatch")
    check(_lib.load().fm_splitk_reduce(ptr(workspace), int(ksplit), M, N, workspace.stride(1), ptr(rscale), ptr(rbias),
                                       ptr(C_out), C_out.stride(0), _stream()), "fm_splitk_reduce")


def row_means(X, out=None):
    _need(X, torch.float32, "X", 2)
    rows, cols = X.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.float64, device=X.device)
    check(_lib.load().fm_row_means(ptr(X), X.stride(0), rows, cols, ptr(out), _stream()), "fm_row_means")
    return out


def gram_assemble(workspace, ksplit, n, mean, ncols, upper_only, out=None):
    if out is None:
        out = torch.empty((n, n), dtype=torch.float64, device=workspace.device)
    check(_lib.load().fm_gram_assemble(ptr(workspace), int(ksplit), int(n), workspace.stride(-2), ptr(mean), int(ncols),
                                       int(bool(upper_only)), ptr(out), _stream()), "fm_gram_assemble")
    return out


def row_dot(Y, w):
    """q[r] = sum_c Y[r,c] * w[c] in float64."""
    _need(Y, torch.float32, "Y", 2)
    _need(w, torch.float64, "w")
    q = torch.empty(Y.shape[0], dtype=torch.float64, device=Y.device)
    check(_lib.load().fm_row_dot(ptr(Y), Y.stride(0), Y.shape[0], Y.shape[1], ptr(w), ptr(q), _stream()), "fm_row_dot")
    return q


def gram_assemble_rowside(workspace, ksplit, n, q, mu, upper_only):
    G = torch.empty((n, n), dtype=torch.float64, device=workspace.device)
    check(_lib.load().fm_gram_assemble_rowside(ptr(workspace), int(ksplit), int(n), workspace.stride(-2), ptr(q), ptr(mu),
                                               mu.numel(), int(bool(upper_only)), ptr(G), _stream()),
          "fm_gram_assemble_rowside")
    return G


def sign_from_right(T, mu, Ut):
    """flip rows of Ut [k, p] by the sign rule on the right singular vectors; T = Ut @ Y [k, c]."""
    check(_lib.load().fm_sign_from_right(ptr(T), T.stride(0), Ut.shape[0], T.shape[1], ptr(mu), ptr(Ut), Ut.stride(0),
                                         Ut.shape[1], _stream()), "fm_sign_from_right")


EIGH_CLUSTER_NMAX = 1024       # largest matrix order fm_eigh_topk_batched takes (one reflector per shared-memory vector)


def eigh_topk_scratch_bytes(n, batch, k):
    return int(_lib.load().fm_eigh_topk_scratch_bytes(int(n), int(batch), int(k)))


class EigCheckFailed(FacemapB200Error):
    """the hand-written eigensolver's self-check flagged a batch (e.g. exactly repeated eigenvalues)"""


class Solver:
    """cuSOLVER handle + workspaces (fm_solver), and the hand-written batched solver for chunk-sized matrices."""

    def __init__(self):
        self._h = C.c_void_p()
        check(_lib.load().fm_solver_create(C.byref(self._h)), "fm_solver_create")

    def syevd(self, G, use_fp32=False):
        """G: float64 [n,n] symmetric, overwritten by eigenvectors (row j <-> j-th smallest eigenvalue).
        Returns lam (float64 [n], ascending).  Asynchronous; `check()` reads cuSOLVER's devInfo back."""
        _need(G, torch.float64, "G", 2)
        n = G.shape[0]
        if G.shape[1] != n or not G.is_contiguous():
            raise FacemapB200Error("syevd: G must be a contiguous square matrix")
        lam = torch.empty(n, dtype=torch.float64, device=G.device)
        check(_lib.load().fm_syevd(self._h, ptr(G), n, ptr(lam), int(bool(use_fp32)), _stream()), "fm_syevd")
        return lam

    def syevdx_top(self, G, k):
        """Only the k largest eigenpairs: rows 0..k-1 of G (ascending), returns lam [k]."""
        _need(G, torch.float64, "G", 2)
        n = G.shape[0]
        if G.shape[1] != n or not G.is_contiguous():
            raise FacemapB200Error("syevdx_top: G must be a contiguous square matrix")
        lam = torch.empty(n, dtype=torch.float64, device=G.device)
        check(_lib.load().fm_syevdx_top(self._h, ptr(G), n, int(k), ptr(lam), _stream()), "fm_syevdx_top")
        return lam[:k]

    def syevd_batched(self, G):
        """G: float64 [batch, n, n]; overwritten by eigenvectors (rows).  Returns lam [batch, n]."""
        _need(G, torch.float64, "G", 3)
        b, n, n2 = G.shape
        if n != n2 or not G.is_contiguous():
            raise FacemapB200Error("syevd_batched: G must be contiguous [batch, n, n]")
        lam = torch.empty((b, n), dtype=torch.float64, device=G.device)
        check(_lib.load().fm_syevd_batched(self._h, ptr(G), n, b, ptr(lam), _stream()), "fm_syevd_batched")
        return lam

    def check(self):
        """devInfo of every cuSOLVER solve since the last check (one D2H copy + stream synchronisation), and the
        self-check flags of the hand-written solver's batches (EigCheckFailed: the caller repeats the pass on
        cuSOLVER)."""
        check(_lib.load().fm_solver_check(self._h, _stream()), "fm_solver_check")
        pending, self._status = getattr(self, "_status", []), []
        for st in pending:
            flags = st.cpu()
            if bool(flags.any()):
                raise EigCheckFailed(f"hand-written eigensolver self-check flags {flags.tolist()}")

    def eigh_topk_batched(self, G, k, scratch=None):
        """fm_eigh_topk_batched: G float64 [batch, n, n] (destroyed), n <= 1024 -> (lam [batch, k], W [batch, k, n]),
        ascending order of eigenvalue.  The self-check flags are read by check()."""
        _need(G, torch.float64, "G", 3)
        b, n, n2 = G.shape
        if n != n2 or not G.is_contiguous():
            raise FacemapB200Error("eigh_topk_batched: G must be contiguous [batch, n, n]")
        lib = _lib.load()
        need = int(lib.fm_eigh_topk_scratch_bytes(n, b, int(k)))
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(need, dtype=torch.uint8, device=G.device)
        lam = torch.empty((b, int(k)), dtype=torch.float64, device=G.device)
        W = torch.empty((b, int(k), n), dtype=torch.float64, device=G.device)
        status = torch.empty(b, dtype=torch.int32, device=G.device)
        check(lib.fm_eigh_topk_batched(ptr(G), n, b, int(k), ptr(lam), ptr(W), ptr(scratch), need, ptr(status), _stream()),
              "fm_eigh_topk_batched")
        if not hasattr(self, "_status"):
            self._status = []
        self._status.append(status)
        return lam, W

    def close(self):
        if self._h:
            _lib.load().fm_solver_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topk_components(evec, lam, k, mean, Wt, want_scale=False, nev=None):
    """Top-k eigenpairs -> Wt[k, n] (sign-fixed, float32), sval[k], rbias[k], rscale[k] (if wanted).
    With want_scale the bias is pre-multiplied by rscale so that a GEMM epilogue
    `rscale*acc + rbias` yields (acc - w.mean)/sval."""
    n = evec.shape[1]
    nev = n if nev is None else int(nev)
    dev = evec.device
    sval = torch.empty(k, dtype=torch.float32, device=dev)
    rbias = torch.empty(k, dtype=torch.float32, device=dev)
    rscale = torch.empty(k, dtype=torch.float32, device=dev) if want_scale else None
    if Wt.shape[0] < k or Wt.shape[1] < n:
        raise FacemapB200Error("topk_components: Wt too small")
    check(_lib.load().fm_topk_components(ptr(evec), ptr(lam), n, nev, int(k), ptr(mean), ptr(Wt), Wt.stride(0), ptr(sval),
                                         ptr(rbias), ptr(rscale), _stream()), "fm_topk_components")
    return sval, rbias, rscale


def transpose(src, dst):
    """dst[c, r] = src[r, c]."""
    _need(src, torch.float32, "src", 2)
    _need(dst, torch.float32, "dst", 2)
    rows, cols = src.shape
    if dst.shape[0] < cols or dst.shape[1] < rows or src.stride(1) != 1 or dst.stride(1) != 1:
        raise FacemapB200Error("transpose: shape mismatch")
    check(_lib.load().fm_transpose(ptr(src), src.stride(0), rows, cols, ptr(dst), dst.stride(0), _stream()),
          "fm_transpose")


def gather_roi_u8(X, rows, col0, Lxb, rect, out):
    """fm_gather_roi_u8: ROI columns of one uint8 plane."""
    _need(X, torch.uint8, "X", 2)
    _need(out, torch.uint8, "out", 2)
    y0, y1, x0, x1 = rect
    check(_lib.load().fm_gather_roi_u8(ptr(X), X.stride(0), int(rows), int(col0), int(Lxb), int(y0), int(y1), int(x0),
                                       int(x1), ptr(out), out.stride(0), _stream()), "fm_gather_roi_u8")


def gather_roi(X, rows, col0, Lxb, rect, out):
    y0, y1, x0, x1 = rect
    check(_lib.load().fm_gather_roi(ptr(X), X.stride(0), int(rows), int(col0), int(Lxb), int(y0), int(y1), int(x0),
                                    int(x1), ptr(out), out.stride(0), _stream()), "fm_gather_roi")


# ---------------------------------------------------------------------------------------------
# full-resolution ROI processors (include/facemap_b200.h: fm_paint, fm_blink, fm_running, fm_pupil)
# ---------------------------------------------------------------------------------------------
def _frames3(frames):
    _need(frames, torch.uint8, "frames", 3)
    if not frames.is_contiguous():
        raise FacemapB200Error("frames must be contiguous [T, H, W]")
    return frames.shape


def paint(frames, ormask, out):
    """out[t] = frames[t] | ormask (uint8 [H, W] of 0x00 / 0xFF)."""
    T, H, W = _frames3(frames)
    _need(ormask, torch.uint8, "ormask", 2)
    _need(out, torch.uint8, "out", 3)
    if tuple(ormask.shape) != (H, W) or out.shape[0] < T or tuple(out.shape[1:]) != (H, W) \
            or not ormask.is_contiguous() or not out.is_contiguous():
        raise FacemapB200Error("paint: shape mismatch")
    check(_lib.load().fm_paint(ptr(frames), T, H, W, ptr(ormask), ptr(out), _stream()), "fm_paint")
    return out[:T]


def _crop_operand(t, rect, name, dtype=torch.uint8):
    y0, y1, x0, x1 = rect
    _need(t, dtype, name, 2)
    if tuple(t.shape) != (y1 - y0, x1 - x0) or not t.is_contiguous():
        raise FacemapB200Error(f"{name}: expected a contiguous [{y1 - y0}, {x1 - x0}] tensor, got {tuple(t.shape)}")


def blink(frames, rect, mask, term, out):
    """out[t] (float64) = sum over the painted ROI of term[pixel] (term: float64 [256])."""
    T, H, W = _frames3(frames)
    _crop_operand(mask, rect, "mask")
    _need(term, torch.float64, "term", 1)
    _need(out, torch.float64, "out", 1)
    if term.numel() != 256 or out.numel() < T or not out.is_contiguous():
        raise FacemapB200Error("blink: term must have 256 entries and out at least T")
    y0, y1, x0, x1 = (int(v) for v in rect)
    check(_lib.load().fm_blink(ptr(frames), T, H, W, y0, y1, x0, x1, ptr(mask), ptr(term), ptr(out), _stream()),
          "fm_blink")


def running(frames, rect, ormask, g2, means, out, prev_crop=None, prev_mean=None):
    """out[t] = (dy, dx) int32 between frame t and its predecessor (row 0 needs prev_crop / prev_mean)."""
    T, H, W = _frames3(frames)
    _crop_operand(ormask, rect, "ormask")
    _crop_operand(g2, rect, "g2", torch.float32)
    _need(means, torch.float32, "means", 1)
    _need(out, torch.int32, "out", 2)
    if prev_crop is not None:
        _crop_operand(prev_crop, rect, "prev_crop")
        _need(prev_mean, torch.float32, "prev_mean", 1)
    if means.numel() < T or out.shape[0] < T or out.shape[1] != 2 or not out.is_contiguous():
        raise FacemapB200Error("running: means / out too small")
    y0, y1, x0, x1 = (int(v) for v in rect)
    check(_lib.load().fm_running(ptr(frames), T, H, W, y0, y1, x0, x1, ptr(ormask), ptr(prev_crop), ptr(prev_mean),
                                 ptr(g2), ptr(means), ptr(out), _stream()), "fm_running")


def pupil_scratch_bytes(rect, blocks):
    y0, y1, x0, x1 = rect
    n = C.c_size_t(0)
    check(_lib.load().fm_pupil_scratch_bytes(int(y1 - y0), int(x1 - x0), int(blocks), C.byref(n)), "fm_pupil_scratch_bytes")
    return int(n.value)


def pupil_scratch(rect, blocks, device):
    return torch.empty(pupil_scratch_bytes(rect, blocks), dtype=torch.uint8, device=device)


def pupil(frames, rect, mask, dark_offset, sigma, scratch, blocks, out, reflector=None):
    """out[t] (float64 [T, 9]) = com(2), area, axdir(4), axlen(2) of the Gaussian pupil fit."""
    T, H, W = _frames3(frames)
    _crop_operand(mask, rect, "mask")
    if reflector is not None:
        _crop_operand(reflector, rect, "reflector")
    _need(scratch, torch.uint8, "scratch", 1)
    _need(out, torch.float64, "out", 2)
    if out.shape[0] < T or out.shape[1] != 9 or not out.is_contiguous():
        raise FacemapB200Error("pupil: out must be a contiguous [>=T, 9] float64 tensor")
    y0, y1, x0, x1 = (int(v) for v in rect)
    if scratch.numel() < pupil_scratch_bytes(rect, blocks):
        raise FacemapB200Error("pupil: scratch too small for this ROI / block count")
    check(_lib.load().fm_pupil(ptr(frames), T, H, W, y0, y1, x0, x1, ptr(mask), ptr(reflector),
                               C.c_float(float(dark_offset)), C.c_double(float(sigma)), ptr(scratch), int(blocks),
                               ptr(out), _stream()), "fm_pupil")


# ---------------------------------------------------------------------------------------------
# motion energy in the reference's floating-point arithmetic (fm_motion_ref)
# ---------------------------------------------------------------------------------------------
EXACT_BYTE_SUM_PIXELS = 65793      # 255 * 65 793 < 2^24: a float32 sum of that many byte values never rounds


def exact_groups(ln, comb, limit=EXACT_BYTE_SUM_PIXELS):
    """For sums of BYTE-valued elements (sbin 1): the maximal subtrees of NumPy's pairwise tree that hold at most
    `limit` elements.  Every partial sum inside them is an integer below 2^24, i.e. exact in float32 in any order, so
    a subtree's value is the integer sum of its leaves; only the merges above them round.  Returns (first leaf, number
    of leaves, comb) per group, left to right: the fold plan over the groups (comb = merges right after a group)."""
    import numpy as np

    ln, comb = np.asarray(ln, np.int64), np.asarray(comb, np.int64)
    stack, found, ends = [], [], {}
    for l in range(len(ln)):
        stack.append([l, 1, int(ln[l])])
        for _ in range(int(comb[l])):
            b = stack.pop()
            a = stack.pop()
            if a[2] + b[2] > limit:                       # a rounding merge: its exact children are maximal
                for node in (a, b):
                    if node[2] <= limit:
                        found.append((node[0], node[1]))
                e = b[0] + b[1] - 1
                ends[e] = ends.get(e, 0) + 1
                stack.append([a[0], a[1] + b[1], limit + 1])          # never exact again
            else:
                stack.append([a[0], a[1] + b[1], a[2] + b[2]])
    if not found:                                         # the whole tree is exact
        found = [(0, len(ln))]
    found.sort()
    first = np.array([f for f, _ in found], np.int32)
    count = np.array([c for _, c in found], np.int32)
    gcomb = np.array([ends.get(int(f + c - 1), 0) for f, c in found], np.uint8)
    assert int(count.sum()) == len(ln) and (first[1:] == first[:-1] + count[:-1]).all()
    return first, count, gcomb


def pairwise_plan(n, device, groups=False):
    """(lo int32, len int32, comb uint8) device tensors: the leaves of NumPy's pairwise sum over n elements.
    groups=True appends the exact-subtree fold plan of byte-valued sums (exact_groups): three more tensors."""
    import numpy as np

    cap = max(1, n // 64 + 2)
    lo, ln, comb = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.uint8)
    count = int(_lib.load().fm_pairwise_plan(int(n), cap, C.c_void_p(lo.ctypes.data), C.c_void_p(ln.ctypes.data),
                                             C.c_void_p(comb.ctypes.data)))
    if count < 1 or count > cap:
        raise FacemapB200Error(f"fm_pairwise_plan({n}) failed ({count})")
    arrays = [lo[:count].copy(), ln[:count].copy(), comb[:count].copy()]
    if groups:
        arrays += list(exact_groups(arrays[1], arrays[2]))
    return tuple(torch.from_numpy(a).to(device) for a in arrays)


def motion_ref_scratch(H, W, sbin, nleaves, blocks, device):
    n = C.c_size_t(0)
    check(_lib.load().fm_motion_ref_scratch_bytes(int(H), int(W), int(sbin), int(nleaves), int(blocks), C.byref(n)),
          "fm_motion_ref_scratch_bytes")
    return torch.empty(int(n.value), dtype=torch.uint8, device=device)


def motion_ref(frames, sbin, plan, scratch, blocks, out, prev_frame=None, ormask=None):
    """out[r] (float64) = the reference's `imbin_mot.sum(-1)` of one view, bit for bit (rows as bin_motion)."""
    T, H, W = _frames3(frames)
    lo, ln, comb = plan[:3]
    _need(out, torch.float64, "out", 1)
    rows = T if prev_frame is not None else T - 1
    if out.numel() < rows or not out.is_contiguous():
        raise FacemapB200Error("motion_ref: out too small")
    for t, name in ((prev_frame, "prev_frame"), (ormask, "ormask")):
        if t is not None:
            _need(t, torch.uint8, name, 2)
            if tuple(t.shape) != (H, W) or not t.is_contiguous():
                raise FacemapB200Error(f"{name}: expected a contiguous [{H}, {W}] uint8 tensor")
    check(_lib.load().fm_motion_ref(ptr(frames), T, H, W, int(sbin), ptr(prev_frame), ptr(ormask), ptr(lo), ptr(ln),
                                    ptr(comb), int(lo.numel()), ptr(scratch), int(blocks), ptr(out), _stream()),
          "fm_motion_ref")
    return rows


def motion_ref_planes(frames, plan, scratch, blocks, out, plane, col0=0, energy=None, prev_frame=None, ormask=None):
    """fm_motion_ref_planes (sbin 1): one pass over the frames of a view for the reference-arithmetic motion sums
    (`out`, float64 [rows]), the |difference| bytes (`plane`: uint8 [rows, >= col0 + H * W] row-major) and the integer
    energy (`energy`: int64 [rows], +=)."""
    T, H, W = _frames3(frames)
    lo, ln, comb = plan[:3]
    gf, gn, gc = plan[3:6] if len(plan) >= 6 else (None, None, None)
    _need(out, torch.float64, "out", 1)
    _need(plane, torch.uint8, "plane", 2)
    _need(energy, torch.int64, "energy", 1)
    rows = T if prev_frame is not None else T - 1
    if out.numel() < rows or not out.is_contiguous() or (energy is not None and (energy.numel() < rows or not energy.is_contiguous())):
        raise FacemapB200Error("motion_ref_planes: out / energy too small")
    if plane.shape[0] < rows or plane.stride(1) != 1 or plane.shape[1] < col0 + H * W:
        raise FacemapB200Error("motion_ref_planes: plane must be row-major [>= rows, >= col0 + H * W]")
    for t, name in ((prev_frame, "prev_frame"), (ormask, "ormask")):
        if t is not None:
            _need(t, torch.uint8, name, 2)
            if tuple(t.shape) != (H, W) or not t.is_contiguous():
                raise FacemapB200Error(f"{name}: expected a contiguous [{H}, {W}] uint8 tensor")
    check(_lib.load().fm_motion_ref_planes(ptr(frames), T, H, W, ptr(prev_frame), ptr(ormask), ptr(lo), ptr(ln), ptr(comb),
                                           int(lo.numel()), ptr(gf), ptr(gn), ptr(gc), 0 if gf is None else int(gf.numel()),
                                           ptr(scratch), int(blocks), ptr(out), ptr(plane), plane.stride(0),
                                           int(col0), ptr(energy), _stream()), "fm_motion_ref_planes")
    return rows


def mean_ref(frames, sbin, mean_bin, mean_adiff):
    """Segment means of one view in the reference's float64 arithmetic (fm_mean_ref; non-power-of-two sbin)."""
    T, H, W = _frames3(frames)
    npix = (H // sbin) * (W // sbin)
    for t, name in ((mean_bin, "mean_bin"), (mean_adiff, "mean_adiff")):
        _need(t, torch.float64, name, 1)
        if t.numel() < npix or not t.is_contiguous():
            raise FacemapB200Error(f"{name}: needs {npix} contiguous float64 elements")
    check(_lib.load().fm_mean_ref(ptr(frames), T, H, W, int(sbin), ptr(mean_bin), ptr(mean_adiff), _stream()),
          "fm_mean_ref")


_STREAMS = {}


def shared_stream(device, name):
    """One stream per (device, purpose) for the life of the process.  Creating a stream — like cudaMemGetInfo — goes
    through the kernel-mode driver, and such calls wait whenever something else holds it (an NVML poller on the box:
    single queries were seen to take 50 - 100 ms, profiles/r02_logs/bench_r2z2.json); kernel launches do not.  The
    per-pass side streams (stage 3's early bin pass, the host sink, the frame upload) are therefore made once."""
    key = (torch.device(device).index or 0, name)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(device)
    return _STREAMS[key]


def free_memory(device, need):
    """Free device memory for a decision about `need` bytes — from the allocator's own books (no driver call) when
    the request is under a quarter of what they show, from cudaMemGetInfo otherwise."""
    total = torch.cuda.get_device_properties(device).total_memory
    est = total - torch.cuda.memory_reserved(device)
    if need <= est // 4:
        return est
    return torch.cuda.mem_get_info(device)[0]
