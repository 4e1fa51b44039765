"""Device pipeline of the motion/movie-SVD path: the three stages of facemap.process.run rebuilt on
the C-ABI kernels.  Python here only sequences kernel launches and owns buffers (torch = memory,
streams, collectives); every array operation on frames or matrices is a kernel of this library
(or the one cuSOLVER call).

Reference map (facemap = /root/reference/facemap):
  stage1_means      subsampled_mean           process.py:53-102
  stage2_masks      compute_SVD               process.py:105-299  (svdecon: utils.py:751-776)
  stage3_project    process_ROIs (SVD parts)  process.py:302-530

svdecon is evaluated in its Gram form (matlab/utils/svdecon.m:16-43; the dead block at
utils.py:753-772) with the PCA centring of the live code (sklearn _pca.py:733-736) folded in as a
rank-1 correction:  G = X X^T - p m m^T,  G = W L W^T,  U S = X^T W - 1 (m^T W).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import kernels as K
from ._lib import FacemapB200Error, Rois
from .utils import binned_inds

NCOMPS = 500        # process.py:764
CHUNK_SVD = 1000    # process.py:134
BUDGET_SVD = 15000  # process.py:135
NC_CHUNK = 250      # process.py:136
NUM_SMS = 148
BATCHED_EIG_MIN = 5


class StageTimer:
    """CUDA-event stopwatch: `with timer("name"):` accumulates device time per label; nothing is
    synchronised until report().  Events come from a process-wide pool (creating and destroying ~600
    events per pass showed up as host-side stalls)."""

    _pool = []

    def __init__(self, enabled=True):
        self.enabled = enabled
        self.spans = []
        self.origin = None
        if enabled:
            self.origin = self._event()
            self.origin.record()

    @classmethod
    def _event(cls):
        return cls._pool.pop() if cls._pool else torch.cuda.Event(enable_timing=True)

    class _Span:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            if self.owner.enabled:
                self.a = self.owner._event()
                self.b = self.owner._event()
                self.a.record()
            return self

        def __exit__(self, *exc):
            if self.owner.enabled:
                self.b.record()
                self.owner.spans.append((self.name, self.a, self.b))
            return False

    def __call__(self, name):
        return StageTimer._Span(self, name)

    def report(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.spans:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def timeline(self):
        """device-time offset (ms since construction) at which the LAST span of every label ended."""
        torch.cuda.synchronize()
        out = {}
        if self.origin is not None:
            for name, a, b in self.spans:
                out[name] = max(out.get(name, 0.0), self.origin.elapsed_time(b))
        return out

    def release(self):
        """hand the events back to the pool (call after report / timeline)."""
        for _, a, b in self.spans:
            StageTimer._pool.append(a)
            StageTimer._pool.append(b)
        if self.origin is not None:
            StageTimer._pool.append(self.origin)
        self.spans, self.origin = [], None


# tcgen05 accumulates in float32 with truncation: every MMA into a running sum loses up to one ulp
# of it, always toward zero (measured on B200: -5.4e-8 relative per accumulation, scripts/probe_trunc.py),
# so the bias grows with the length of a K chain.  The product GEMM (impl 0, outputs wider than 128) bounds
# its TMEM chains to 128 K inside the kernel and sums them in float32 registers with round-to-nearest;
# split-K over CTAs is then only needed to fill the machine.  Narrow outputs (<= 128 columns) run the
# single-accumulator kernel and are cut into K chains here, reduced in float64.
# Gram matrices feed an eigensolve whose SMALL eigenvalues are sensitive to unstructured error: their K range is
# still cut into many short partial sums reduced in float64 (fm_gram_assemble), so that at most 2-3 chains are
# added in float32 before the exact reduction.
KCHAIN_NARROW = 256
KCHAIN_GRAM = 320     # <= 3 in-kernel chains per float64-reduced partial
# The left singular vectors (U S = X^T W, U = Y W / S) feed the next Gram / are the masks themselves: their small
# components come out of large cancelling sums, so they too keep float64-reduced partials (2 in-kernel chains each).
# The projection of every frame (trace tolerance 1e-3) runs as one launch: 150 chains summed in registers.
KCHAIN_LEFT = 256
LEFT_BLOCK_PIXELS = 131072   # pixels per left-vector GEMM (a multiple of 32: block boundaries stay 128-byte aligned)
KCHAIN_PROJ = 1200    # only when the projection runs the single-accumulator kernel (impl 9)


def _fill_ksplit(tiles, kdim):
    """split-K factor that only serves parallelism: ~2 waves of CTAs, at least 1024 K elements each."""
    want = math.ceil(2 * NUM_SMS / max(tiles, 1)) if tiles < NUM_SMS else 1
    return int(max(1, min(want, kdim // 1024 if kdim >= 2048 else 1, 64)))


def _gram_ksplit(n, kdim):
    """split-K factor for an n x n upper-triangular Gram over `kdim`."""
    if n <= 128:
        return int(max(1, min(math.ceil(kdim / KCHAIN_NARROW), 64)))
    mt, nt = math.ceil(n / 128), math.ceil(n / 256)
    tiles = sum(1 for i in range(mt) for j in range(nt) if (j + 1) * 256 - 1 >= i * 128)
    fill = math.ceil(2 * NUM_SMS / max(tiles, 1)) if tiles < NUM_SMS else 1
    return int(max(1, min(max(fill, math.ceil(kdim / KCHAIN_GRAM)), math.ceil(kdim / 32), 192)))


def _gemm_ksplit(M, N, kdim, chain=None):
    if chain is not None or N <= 128:
        return int(max(1, min(math.ceil(kdim / (chain or KCHAIN_NARROW)), 64)))
    return _fill_ksplit(math.ceil(M / 128) * math.ceil(N / 256), kdim)


_SOLVERS = {}


def get_solver(device):
    """One cuSOLVER handle (+ its device workspace) per device for the life of the process: creating
    and destroying it per call costs cudaMalloc/cudaFree synchronisations."""
    key = torch.device(device).index or 0
    if key not in _SOLVERS:
        _SOLVERS[key] = K.Solver()
    return _SOLVERS[key]


class SvdWorkspace:
    """Reusable device buffers for the Gram -> eigensolve -> left-vector sequence."""

    def __init__(self, device, eig_fp32=False, gemm_impl=0, timer=None, gram_mode=None, proj_mode=None, eig_mode=None):
        self.device = device
        self.timer = timer if timer is not None else StageTimer(enabled=False)
        self.solver = get_solver(device)
        # per-role kernel choice (tuning / tests): 0 = chained accumulation, 9 = one TMEM chain per partial
        import os as _os
        self.gram_impl = int(_os.environ.get("FM_GRAM_IMPL", gemm_impl))
        self.left_impl = int(_os.environ.get("FM_LEFT_IMPL", gemm_impl))
        self.proj_impl = int(_os.environ.get("FM_PROJ_IMPL", gemm_impl))
        self.eig_fp32 = eig_fp32
        self.gemm_impl = gemm_impl
        # how the Gram matrices that feed the eigensolves are formed:
        #   "i8"            (default) s8/u8/u8 digit planes and exact integer products on tcgen05 kind::i8 (i8_gram) at
        #                   both SVD levels: all 500 singular values within 3e-6 of the exact float64 SVD
        #   "i8_merge"      only the second-level (merge) Gram that way; chunk Grams as "tf32x3"
        #   "tf32x3"        one 3xTF32 GEMM with float64-reduced K partials (centred_gram): tail singular values 5e-4 (motion)
        #                   to 3e-2 (movie) off at component 500 (profiles/r01_accuracy.md) — kept for comparison
        #   "sliced_merge"  the merge Gram from 8-bit float row slices with an exact leading product (sliced_gram)
        #   "sliced"        chunk Grams as well (~4x the Gram time of "tf32x3")
        self.gram_mode = _os.environ.get("FM_GRAM_MODE", "i8") if gram_mode is None else gram_mode
        if self.gram_mode not in ("tf32x3", "sliced_merge", "sliced", "i8_merge", "i8"):
            raise FacemapB200Error(f"unknown gram_mode {self.gram_mode!r}")
        # pass A of sliced_gram on the one-product kernel (fm_gemm_tn impl 10) instead of the 3xTF32 kernel with a
        # zero lo plane: a third of the MMAs, no zero plane to stage
        self.sliced_single = _os.environ.get("FM_SLICED_SINGLE", "0") == "1"
        # how stage 3 projects the frames onto the masks:
        #   "i8"      (default) K1 writes the integer energies / block sums as byte planes, the masks become s8/u8/u8
        #             digits and fm_project_i8 forms the products exactly on tcgen05 kind::i8 (Stage3): 2x the 3xTF32
        #             rate at sbin 4, 9x at sbin 1 on large views.  Two byte planes hold block sums up to sbin 16;
        #             beyond that the default is "tf32x3" (an explicit "i8" is refused there, not silently replaced)
        #   "tf32x3"  K1 writes float32 operands, one 3xTF32 GEMM per batch
        self.proj_mode = _os.environ.get("FM_PROJ_MODE") if proj_mode is None else proj_mode
        self.proj_mode_explicit = self.proj_mode is not None
        if self.proj_mode is None:
            self.proj_mode = "i8"
        if self.proj_mode not in ("tf32x3", "i8"):
            raise FacemapB200Error(f"unknown proj_mode {self.proj_mode!r}")
        # who solves the per-chunk eigenproblems (n = 999, a batch of them):
        #   "cluster"   (default) the hand-written solver, one thread-block cluster per matrix (csrc/eigh_cluster.cu)
        #   "cusolver"  cusolverDnXsyevBatched / syevd (host-driven: ~4.7 ms per matrix batched, 12 ms serial)
        self.eig_mode = _os.environ.get("FM_EIG", "cluster") if eig_mode is None else eig_mode
        if self.eig_mode not in ("cluster", "cusolver"):
            raise FacemapB200Error(f"unknown eig_mode {self.eig_mode!r}")
        # merge-level eigensolve: "subspace" (block subspace iteration + Rayleigh-Ritz for large problems,
        # subspace_topk) or "cusolver" (syevdx always); a pass repeated on cuSOLVER (eig_mode "cusolver") uses syevdx
        self.merge_eig = _os.environ.get("FM_MERGE_EIG", MERGE_EIG_DEFAULT) if self.eig_mode == "cluster" else "cusolver"
        if self.merge_eig not in ("subspace", "cusolver"):
            raise FacemapB200Error(f"unknown FM_MERGE_EIG {self.merge_eig!r}")
        self.proj_i8_impl = int(_os.environ.get("FM_PROJ_I8_IMPL", "0"))      # 1 = CTA pairs (fm_project_i8 impl)
        self._ws = None
        self._bufs = {}

    def gemm(self, A, B, C_out, rscale=None, rbias=None, label="gemm", B_lo=None, chain=None, impl=None):
        """C = rscale * (A @ B^T) + rbias.  `chain` cuts K into partial sums of that length reduced in float64;
        without it K is split only where the tile count cannot fill the GPU."""
        ksplit = _gemm_ksplit(A.shape[0], B.shape[0], A.shape[1], chain)
        impl = self.gemm_impl if impl is None else impl
        if impl not in (0, 9):
            B_lo = None
        if ksplit == 1:
            with self.timer(label + ".mma"):
                K.gemm_tn(A, B, C_out, rscale=rscale, rbias=rbias, impl=impl, B_lo=B_lo)
            return
        M, N = A.shape[0], B.shape[0]
        ld = max(32, K.round_up(N, 32))
        need = ksplit * M * ld
        buf = self._bufs.get("splitk")
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.float32, device=self.device)
            self._bufs["splitk"] = buf
        part = buf[:need].view(ksplit, M, ld)
        with self.timer(label + ".mma"):
            K.gemm_tn(A, B, ksplit=ksplit, workspace=part, impl=impl, B_lo=B_lo)
        with self.timer(label + ".reduce"):
            K.splitk_reduce(part, ksplit, C_out, rscale=rscale, rbias=rbias)

    def workspace(self, ksplit, n):
        ld = max(32, K.round_up(n, 32))
        need = ksplit * n * ld
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=self.device)
        return self._ws[:need].view(ksplit, n, ld)

    def raw(self, tag, numel, dtype):
        """reusable flat buffer of `numel` elements of `dtype` (byte planes, int32 partials)"""
        buf = self._bufs.get(tag)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
            self._bufs[tag] = buf
        return buf[:numel]

    def release(self, keep=("X3",)):
        """drop the reusable buffers of a finished phase (all but the tags starting with one of `keep`: stage 3's
        operands may already be filling on the side stream) so that the next phase's do not add to them"""
        for tag in [t for t in self._bufs if not t.startswith(tuple(keep))]:
            del self._bufs[tag]
        self._ws = None

    def matrix(self, tag, rows, cols):
        ld = max(32, K.round_up(cols, 32))
        need = max(rows, 1) * ld
        buf = self._bufs.get(tag)
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.float32, device=self.device)
            self._bufs[tag] = buf
        return buf[:need].view(max(rows, 1), ld)[:rows, :cols]


def gram_is_i8(ws: SvdWorkspace, label):
    return ws.gram_mode == "i8" or (ws.gram_mode == "i8_merge" and label.endswith("merge"))


def gram_is_sliced(ws: SvdWorkspace, label, n):
    """sliced_gram serves the levels `ws.gram_mode` names; it needs the 128 x 256 chain kernel (K dealt in 16-wide
    k-blocks), i.e. impl 0 and an output wider than 128."""
    if ws.gram_mode not in ("sliced", "sliced_merge") or ws.gram_impl != 0 or n <= 128:
        return False
    return ws.gram_mode == "sliced" or label.endswith("merge")


def centred_gram(ws: SvdWorkspace, X, timer, label, out=None):
    """mean over columns of every row of X [n, p] and the centred Gram  X X^T - p m m^T  (float64)."""
    n, p = X.shape
    if gram_is_i8(ws, label):
        return i8_gram(ws, X, timer, label, out=out)
    if gram_is_sliced(ws, label, n):
        return sliced_gram(ws, X, timer, label, out=out)
    with timer(label + ".colmean"):
        mean = K.row_means(X)
    ksplit = _gram_ksplit(n, p)
    part = ws.workspace(ksplit, n)
    # A = B = X: one lo plane serves both operands, TMA stages all four tiles (no in-kernel split)
    X_lo = None
    if ws.gram_impl in (0, 9):
        with timer(label + ".gram.split"):
            X_lo = K.split_lo(X, out=ws.matrix("gram.lo", n, p))
    with timer(label + ".gram.mma"):
        if ksplit > 1:
            K.gemm_tn(X, X, ksplit=ksplit, workspace=part, upper_only=True, impl=ws.gram_impl, B_lo=X_lo, A_lo=X_lo)
        else:
            K.gemm_tn(X, X, part[0], upper_only=True, impl=ws.gram_impl, B_lo=X_lo, A_lo=X_lo)
    with timer(label + ".gram.assemble"):
        G = K.gram_assemble(part, ksplit, n, mean, p, True, out=out)
    return mean, G


SLICE_KBLOCKS = 15     # k-blocks of 16 per exact partial of the top-slice product: 2^16 * 240 < 2^24


def sliced_gram(ws: SvdWorkspace, X, timer, label, out=None):
    """centred_gram with the rows of X cut into three 8-bit slices X = S0 + S1 + S2 (fm_slice_rows):

        pass A   S0 S0^T                       K partials of <= 240: every product and every partial sum is an
                                               integer (times a power of two) below 2^24, so the tensor cores'
                                               truncating float32 accumulation never rounds — exact
        pass B   S0 S1^T + S1 S0^T + S1 S1^T   2^-8 of A:  the 3xTF32 kernel with (lo, hi) = (S0, S1)
        pass C   ... with (S0, S2)             2^-16 of A
        pass D   ... with (S1, S2)             2^-24 of A  (S2 S2^T enters twice: 2^-32)

    all into one stack of partials that fm_gram_assemble sums in float64.  What is left is the float32 rounding of
    passes B-D (5e-8 of 2^-8 of an entry) and the 2^-25 representation error of the slices, instead of 5e-8 of the
    entry itself.  scripts/model_accuracy.py (a NumPy model of this arithmetic, `passes` / `passesm`): with both
    levels sliced every one of the 500 singular values lands within 1e-4 of the exact oracle's, motion and movie,
    where the one-GEMM Gram leaves the tail at 1e-4 .. 4e-3.  Pass A runs on the same kernel with an all-zero lo
    plane, or (ws.sliced_single) on its one-product instantiation, fm_gemm_tn impl 10."""
    n, p = X.shape
    with timer(label + ".colmean"):
        mean = K.row_means(X)
    with timer(label + ".gram.split"):
        S0, S1, S2 = K.slice_rows(X, out=tuple(ws.matrix(f"gram.slice{d}", n, p) for d in range(3)))
        Z = None
        if not ws.sliced_single:
            Z = ws.matrix("gram.zero", n, p)
            Z.zero_()
    ks_a = max(1, math.ceil(math.ceil(p / 16) / SLICE_KBLOCKS))
    mt, nt = math.ceil(n / 128), math.ceil(n / 256)
    tiles = sum(1 for i in range(mt) for j in range(nt) if (j + 1) * 256 - 1 >= i * 128)
    ks_f = _fill_ksplit(tiles, p)
    part = ws.workspace(ks_a + 3 * ks_f, n)
    with timer(label + ".gram.mma"):
        at = 0
        for hi, lo, ks in ((S0, Z, ks_a), (S1, S0, ks_f), (S2, S0, ks_f), (S2, S1, ks_f)):
            impl = 0 if lo is not None else 10
            if ks > 1:
                K.gemm_tn(hi, hi, ksplit=ks, workspace=part[at:at + ks], upper_only=True, impl=impl, B_lo=lo, A_lo=lo)
            else:
                K.gemm_tn(hi, hi, part[at], upper_only=True, impl=impl, B_lo=lo, A_lo=lo)
            at += ks
    with timer(label + ".gram.assemble"):
        G = K.gram_assemble(part, at, n, mean, p, True, out=out)
    return mean, G


I8_KMAX = 32768        # K per int32 partial of fm_gemm_i8_tn (255^2 * 32768 < 2^31)
I8_BLOCK_COLS = 262144  # columns of a Gram operand sliced into digit planes at once (8 partials of I8_KMAX)


def i8_gram(ws: SvdWorkspace, X, timer, label, out=None, rowside_mu=None):
    """centred_gram on the integer tensor cores: the rows of X become s8 / u8 / u8 digit planes
    (fm_slice_rows_i8: 23 bits below each row's bound, unbiased), the six plane products are formed by fm_gemm_i8_tn
    with exact int32 accumulation, fm_gram_assemble_i8 weighs and sums them in float64.  No rounding anywhere but the
    2^-24 digit representation: all 500 singular values land within 3e-6 of the exact oracle's (measured on B200,
    profiles/r02_accuracy.md).  3 bytes per element are staged per product instead of the 8 (hi + lo planes) of the
    3xTF32 Gram.
    rowside_mu: X is a wide matrix Y [p, c] centred over its COLUMNS (means `rowside_mu` [c]); returns (q, G) with
    q = Y mu and G = Yc Yc^T (merge_components_rowside)."""
    n, p = X.shape
    with timer(label + ".colmean"):
        mean = K.row_means(X) if rowside_mu is None else K.row_dot(X, rowside_mu)
    # operands too wide to slice at once (sbin 1 on a large view: p = 1.3 M columns) go in column blocks, each with
    # its own digit planes, row exponents and int32 partials; the float64 assembly accumulates them
    blk = p if (p <= I8_BLOCK_COLS or rowside_mu is not None) else I8_BLOCK_COLS
    ldp = K.round_up(n, 4)
    G = out
    for c0 in range(0, p, blk):
        c1 = min(p, c0 + blk)
        w = c1 - c0
        with timer(label + ".gram.split"):
            ldq = K.round_up(w, 16)
            raw = ws.raw("gram.i8", 3 * n * ldq, torch.uint8).view(3, n, ldq)
            Q0, Q1, Q2, exps = K.slice_rows_i8(X[:, c0:c1], out=(raw[0].view(torch.int8)[:, :w], raw[1][:, :w], raw[2][:, :w]))
        # split K where the int32 accumulator demands it, and where the tile count alone cannot fill the machine
        ks = max(1, math.ceil(math.ceil(w / 64) / (I8_KMAX // 64)), _fill_ksplit(math.ceil(n / 128) * math.ceil(n / 256), w))
        P = ws.raw("gram.i8.P", 6 * ks * n * ldp, torch.int32).view(6, ks, n, ldp)
        with timer(label + ".gram.mma"):
            for d, (A, B, upper) in enumerate(((Q0, Q0, True), (Q1, Q1, True), (Q2, Q2, True),
                                               (Q0, Q1, False), (Q0, Q2, False), (Q1, Q2, False))):
                if ks > 1:
                    K.gemm_i8_tn(A, B, ksplit=ks, workspace=P[d], upper_only=upper)
                else:
                    K.gemm_i8_tn(A, B, P[d, 0], upper_only=upper)
        with timer(label + ".gram.assemble"):
            if c0 == 0:
                G = K.gram_assemble_i8(P, ks, n, exps, mean, p, out=out,
                                       rowside=None if rowside_mu is None else (mean, rowside_mu))
            else:
                K.gram_accumulate_i8(P, ks, n, exps, G)
    return mean, G


def gram_eig(ws: SvdWorkspace, X, k, timer, want_scale, label):
    """Centred Gram of the rows of X [n, p], eigensolve, top-k.  Returns (Wt [k, n], sval, rbias,
    rscale): the right singular vectors of X^T (sign-fixed), singular values, and the epilogue
    terms for the left-vector GEMM."""
    n, p = X.shape
    mean, G = centred_gram(ws, X, timer, label)
    evec = G
    with timer(label + ".syevd"):
        top = None
        if getattr(ws, "merge_eig", "cusolver") == "subspace" and not ws.eig_fp32 and subspace_eligible(n, k):
            top = subspace_topk(G[:n, :n], k, rr=lambda Hm, kk: _rr_cluster(ws, Hm, kk))
        if top is not None:
            lam, evec, nev = top[0], top[1], k
        elif ws.eig_fp32 or k == n:
            lam, nev = ws.solver.syevd(G, use_fp32=ws.eig_fp32), n
        else:
            lam, nev = ws.solver.syevdx_top(G, k), k      # only the k largest eigenpairs (rows 0..k-1 of G)
    with timer(label + ".topk"):
        Wt = ws.matrix(label + ".Wt", k, n)
        sval, rbias, rscale = K.topk_components(evec, lam, k, mean, Wt, want_scale=want_scale, nev=nev)
    return Wt, sval, rbias, rscale


MERGE_EIG_DEFAULT = "subspace"
SUBSPACE_MIN_N = 3000          # the merge of >= 12 chunks (3 000 ... 3 750 stacked components); smaller ones stay direct
SUBSPACE_EXTRA = 250           # guard vectors next to the k wanted ones
SUBSPACE_MULTS = 8             # multiplications by G before the first Rayleigh-Ritz step
SUBSPACE_MORE = 4              # ... and per further round when the residuals are not there yet
SUBSPACE_ROUNDS = 3
_SUBSPACE_START = {}


def subspace_eligible(n, k):
    return n >= SUBSPACE_MIN_N and 64 <= k and 4 * (k + SUBSPACE_EXTRA) <= n


def _rr_cluster(ws, Hm, k):
    """top-k eigenpairs of the projected matrix on the hand-written solver: (lam [k], Z [k, b]) ascending; its
    self-check flags are read by ws.solver.check() like those of the chunk problems"""
    lam, Wt = ws.solver.eigh_topk_batched(Hm[None].contiguous(), k)
    return lam[0], Wt[0]


def _rr_lapack(Hm, k):
    th, Z = torch.linalg.eigh(Hm)
    return th[-k:], Z[:, -k:].T.contiguous()


def subspace_topk(G, k, rr=_rr_lapack, mults=None, seed=0):
    """The k largest eigenpairs of the symmetric positive semi-definite G [n, n] (float64) by block subspace
    iteration with Rayleigh-Ritz: Q <- orth(G Q) for a block of k + 250 vectors (float64 GEMMs, Cholesky-QR after
    every multiplication: the spectrum spans 1e6 and more), then the eigenpairs of Q^T G Q.  Measured on the merge
    problem of the benchmark (n = 3 750, k = 500, B200): 8 multiplications bring every singular value within 1e-8
    and every gap-resolved vector within 1e-7 of LAPACK's, in 25 ms against 69 - 74 ms for cusolverDnXsyevdx
    (profiles/r02_eig.md, scripts/probe_merge_subspace.py).

    Accepted only if the a-posteriori estimates are well inside the path's gates, with up to two more rounds of
    multiplications; otherwise, or if a Cholesky factor breaks down, None: the caller solves directly.  With the
    residuals r_j = G w_j - theta_j w_j: (a) eigenvalues — the residuals point out of the block, where G's spectrum
    lies below `edge` (the mean of the block's unwanted Ritz values, an upper bound of it), so theta_j is off by about
    |r_j|^2 / (theta_j - edge); required: 1e-5 theta_j (singular value to 5e-6; the gate is 1e-4).  (b) vectors — a
    component whose singular value stands 4e-4 S_0 clear of both neighbours (what the tests call resolved) turns by
    at most |r_j| / gap_j; required: 2.5e-4 (the gate is 1e-3).
    Returns (lam [k] ascending, W [k, n] rows = eigenvectors, ascending) — what fm_topk_components reads with nev = k.
    Deterministic: the start block is a fixed pseudo-random matrix."""
    n = G.shape[0]
    b = k + SUBSPACE_EXTRA
    key = (str(G.device), n, b, seed)
    if key not in _SUBSPACE_START:
        g = torch.Generator(device="cpu")
        g.manual_seed(1234 + seed)
        _SUBSPACE_START[key] = torch.randn((n, b), dtype=torch.float64, generator=g).to(G.device)
    infos = []

    def orth(Z):
        L, info = torch.linalg.cholesky_ex(Z.T @ Z, check_errors=False)
        infos.append(info)
        return torch.linalg.solve_triangular(L, Z.T, upper=False).T.contiguous()

    Q = orth(_SUBSPACE_START[key])
    todo = SUBSPACE_MULTS if mults is None else int(mults)
    for _ in range(SUBSPACE_ROUNDS):
        for _ in range(todo):
            Q = orth(G @ Q)
        GQ = G @ Q
        Hm = Q.T @ GQ
        Hm = ((Hm + Hm.T) / 2).contiguous()
        trace = torch.trace(Hm)                                     # before rr: a solver may destroy its input
        th, Zt = rr(Hm, k)                                          # ascending; rows of Zt are the vectors
        W = Zt @ Q.T                                                # [k, n]
        R = Zt @ GQ.T - th[:, None] * W
        rn = torch.linalg.vector_norm(R, dim=1)
        tiny = 1e-300
        thc = th.clamp_min(0)
        edge = ((trace - th.sum()) / (b - k)).clamp_min(0)
        est = rn * rn / (thc - edge).clamp_min(tiny)
        bad_a = (est > 1e-5 * thc).any()
        S = torch.sqrt(thc)
        lo_nb = torch.cat([torch.sqrt(edge).reshape(1), S[:-1]])               # the neighbour below (ascending order)
        gap_dn, gap_up = S - lo_nb, torch.cat([S[1:] - S[:-1], S[-1:]])
        resolved = torch.minimum(gap_dn, gap_up) >= 4e-4 * S[-1]
        gap_lam = torch.minimum(thc - torch.cat([edge.reshape(1), thc[:-1]]),
                                torch.cat([thc[1:] - thc[:-1], thc[-1:]])).clamp_min(tiny)
        bad_b = (resolved & (rn > 2.5e-4 * gap_lam)).any()
        bad = bad_a | bad_b | torch.stack([i.reshape(()) != 0 for i in infos]).any() | ~torch.isfinite(rn).all()
        infos.clear()
        if not bool(bad):                                           # the one host synchronisation of the solve
            return th, W.contiguous()
        todo = SUBSPACE_MORE
    return None


def left_components(ws, X, evec, lam, mean, k, out_Yt, timer, label, nev=None):
    """(U S)^T = W^T X - (W^T m) 1^T for the top-k eigenvectors (rows of `evec`, `nev` of them in ascending order:
    all n after syevd, the top ones after the hand-written solver) -> out_Yt [k, p]."""
    n, p = X.shape
    with timer(label + ".topk"):
        Wt = ws.matrix(label + ".Wt", k, n)
        _, rbias, _ = K.topk_components(evec, lam, k, mean, Wt, want_scale=False, nev=nev)
    # pixels in blocks: the transposed operand, its lo plane and the split-K partials stay bounded for any view size
    for p0 in range(0, p, LEFT_BLOCK_PIXELS):
        p1 = min(p, p0 + LEFT_BLOCK_PIXELS)
        with timer(label + ".transpose"):
            Xt = ws.matrix(label + ".Xt", p1 - p0, n)
            K.transpose(X[:, p0:p1], Xt)
            Xt_lo = K.split_lo(Xt, out=ws.matrix(label + ".Xt_lo", p1 - p0, n)) if ws.left_impl in (0, 9) else None
        ws.gemm(Wt, Xt, out_Yt[:, p0:p1], rbias=rbias, label=label + ".left", B_lo=Xt_lo, chain=KCHAIN_LEFT,
                impl=ws.left_impl)


def chunk_components(ws, X, k, out_Yt, timer, label="chunk"):
    """svdecon(X^T, k) of one chunk: writes (U S)^T = W^T X - (W^T m) 1^T into out_Yt [k, p]
    (facemap/process.py:246-251; ROI variant :213-242)."""
    mean, G = centred_gram(ws, X, timer, label)
    with timer(label + ".syevd"):
        lam = ws.solver.syevd(G, use_fp32=ws.eig_fp32)
    left_components(ws, X, G, lam, mean, k, out_Yt, timer, label)


def merge_components_rowside(ws, Yt, k, timer, label="merge"):
    """svdecon(Y, k) when Y [p, c] is wide (p < c: the merge of a small ROI): eigensolve of the p x p
    matrix Yc Yc^T, whose eigenvectors are the masks themselves; the sign comes from the right vectors."""
    c, p = Yt.shape
    with timer(label + ".colmean"):
        mu = K.row_means(Yt)                                   # column means of Y = row means of Yt
    with timer(label + ".transpose"):
        Y = ws.matrix(label + ".Y", p, c)
        K.transpose(Yt, Y)
    if gram_is_i8(ws, label):
        _, G = i8_gram(ws, Y, timer, label, rowside_mu=mu)
    else:
        with timer(label + ".transpose"):
            Y_lo = K.split_lo(Y, out=ws.matrix(label + ".Y_lo", p, c)) if ws.gram_impl in (0, 9) else None
        ksplit = _gram_ksplit(p, c)
        part = ws.workspace(ksplit, p)
        with timer(label + ".gram.mma"):
            if ksplit > 1:
                K.gemm_tn(Y, Y, ksplit=ksplit, workspace=part, upper_only=True, impl=ws.gram_impl, B_lo=Y_lo, A_lo=Y_lo)
            else:
                K.gemm_tn(Y, Y, part[0], upper_only=True, impl=ws.gram_impl, B_lo=Y_lo, A_lo=Y_lo)
        with timer(label + ".gram.assemble"):
            q = K.row_dot(Y, mu)
            G = K.gram_assemble_rowside(part, ksplit, p, q, mu, True)
    with timer(label + ".syevd"):
        if ws.eig_fp32 or k == p:
            lam, nev = ws.solver.syevd(G, use_fp32=ws.eig_fp32), p
        else:
            lam, nev = ws.solver.syevdx_top(G, k), k
    with timer(label + ".topk"):
        Ut = K.alloc_matrix(k, p, Yt.device)
        sval, _, _ = K.topk_components(G, lam, k, None, Ut, want_scale=False, nev=nev)
    # sign rule on the right singular vectors: T = Ut Y  (V S = T^T - mu (1^T U))
    T = ws.matrix(label + ".T", k, c)
    ws.gemm(Ut, Yt, T, label=label + ".left", chain=KCHAIN_LEFT, impl=ws.left_impl)
    with timer(label + ".topk"):
        K.sign_from_right(T, mu, Ut)
    with timer(label + ".transpose"):
        U = torch.empty((p, k), dtype=torch.float32, device=Yt.device)
        K.transpose(Ut, U)
    return Ut, U, sval


def merge_components(ws, Yt, k, timer, label="merge"):
    """svdecon(Y, k) for Y = [U_1 S_1 ... ] given as Yt [c, p] (facemap/process.py:264-295).
    Returns (Ut [k, p], U [p, k], sval [k])."""
    c, p = Yt.shape
    if p < c:
        return merge_components_rowside(ws, Yt, k, timer, label)
    Wt, sval, rbias, rscale = gram_eig(ws, Yt, k, timer, True, label)
    Ut = K.alloc_matrix(k, p, Yt.device)
    for p0 in range(0, p, LEFT_BLOCK_PIXELS):
        p1 = min(p, p0 + LEFT_BLOCK_PIXELS)
        with timer(label + ".transpose"):
            Y = ws.matrix(label + ".Y", p1 - p0, c)
            K.transpose(Yt[:, p0:p1], Y)
            Y_lo = K.split_lo(Y, out=ws.matrix(label + ".Y_lo", p1 - p0, c)) if ws.left_impl in (0, 9) else None
        ws.gemm(Wt, Y, Ut[:, p0:p1], rscale=rscale, rbias=rbias, label=label + ".left", B_lo=Y_lo, chain=KCHAIN_LEFT,
                impl=ws.left_impl)
    with timer(label + ".transpose"):
        U = torch.empty((p, k), dtype=torch.float32, device=Yt.device)
        K.transpose(Ut, U)
    return Ut, U, sval


class Geometry:
    def __init__(self, Ly, Lx, sbin, rois):
        self.Ly, self.Lx, self.sbin = list(Ly), list(Lx), int(sbin)
        self.Lyb, self.Lxb, self.ir = binned_inds(Ly, Lx, sbin)
        self.npix_v = [int(a) * int(b) for a, b in zip(self.Lyb, self.Lxb)]
        self.col0 = [0] + list(np.cumsum(self.npix_v)[:-1].astype(int))
        self.p = int(sum(self.npix_v))
        # motion ROIs (rind == 1), in list order: (view, y0, y1, x0, x1) in binned coordinates
        self.mot_rois = []
        for r in (rois or []):
            if r["rind"] == 1:
                yb, xb = r["yrange_bin"], r["xrange_bin"]
                if yb.size == 0 or xb.size == 0:
                    raise FacemapB200Error("motion ROI is empty after binning")
                v, y0, y1, x0, x1 = int(r["ivid"]), int(yb[0]), int(yb[-1]) + 1, int(xb[0]), int(xb[-1]) + 1
                # the reference would fail (or silently wrap) on a rectangle outside its view; here it would read
                # device memory out of bounds — refuse it by name.  (The reproduced x-stop rule, floor(x) / sbin, can
                # reach Lxb + 1 for a ROI touching the right border of a view whose width is no multiple of sbin.)
                if not 0 <= v < len(self.Ly):
                    raise FacemapB200Error(f"motion ROI {len(self.mot_rois)}: ivid {v} outside the {len(self.Ly)} views")
                if not (0 <= y0 < y1 <= int(self.Lyb[v]) and 0 <= x0 < x1 <= int(self.Lxb[v])):
                    raise FacemapB200Error(
                        f"motion ROI {len(self.mot_rois)} on view {v}: binned rows [{y0}, {y1}) x columns [{x0}, {x1}) "
                        f"outside the {int(self.Lyb[v])} x {int(self.Lxb[v])} binned view")
                self.mot_rois.append((v, y0, y1, x0, x1))
        self.roi_npix = [(y1 - y0) * (x1 - x0) for (_, y0, y1, x0, x1) in self.mot_rois]
        self.rois_on_view = [[j for j, r in enumerate(self.mot_rois) if r[0] == v] for v in range(len(Ly))]


def access_plan(source, dist=None):
    """(priority ranges, home range) of one pass for this rank: the stage-1 segments and stage-2
    chunks it will read, in order, and the frame range it sweeps in stage 3 (with its halo)."""
    from .parallel import assign_contiguous, assign_round_robin, frame_ranges

    N = source.nframes
    rank, world = (0, 1) if dist is None else (dist.rank, dist.world)
    nt1, tf1 = stage1_plan(N, source.block_frames)
    nt2, nsegs, _, tf2 = stage2_plan(N)
    pri = [(int(tf1[i]), int(tf1[i]) + nt1) for i in assign_round_robin(len(tf1), world)[rank]]
    pri += [(int(tf2[i]), int(tf2[i]) + nt2) for i in assign_contiguous(nsegs, world)[rank]]
    f0, f1 = frame_ranges(N, world)[rank]
    return pri, (max(0, f0 - 1), f1)


# ---------------------------------------------------------------------------------------------
# stage 1
# ---------------------------------------------------------------------------------------------
def stage1_plan(N, block_frames):
    """facemap/process.py:61-67."""
    nf = min(1000, N)
    nt0 = int(min(100, min(block_frames)))
    nsegs = int(np.floor(nf / nt0))
    tf = np.floor(np.linspace(0, N - nt0, nsegs)).astype(int)
    return nt0, tf


def stage1_means(source, geo: Geometry, device, timer, segments=None):
    """avgframe, avgmotion (float32 [p] device tensors).  `segments` restricts the work to a
    subset of segment indices and returns the raw per-segment integer sums instead (multi-GPU)."""
    N = source.nframes
    nt0, tf = stage1_plan(N, source.block_frames)
    p = geo.p
    avgframe = torch.zeros(p, dtype=torch.float32, device=device)
    avgmotion = torch.zeros(p, dtype=torch.float32, device=device)
    nseg = len(tf)
    todo = range(nseg) if segments is None else segments
    sums = {}
    todo = list(todo)
    for pos, i in enumerate(todo):
        t = int(tf[i])
        with timer("stage1.fetch"):
            frames = source.fetch(t, t + nt0, device)
            if pos + 1 < len(todo):
                source.prefetch(int(tf[todo[pos + 1]]), int(tf[todo[pos + 1]]) + nt0, device)
        # power-of-two sbin: exact integer time sums; otherwise the reference's float64 means (fm_mean_ref)
        dt = torch.float64 if stage1_uses_reference(geo.sbin) else torch.int64
        tsb = torch.zeros(p, dtype=dt, device=device)
        tsa = torch.zeros(p, dtype=dt, device=device)
        with timer("stage1.bin"):
            for v, fr in enumerate(frames):
                c0, n = geo.col0[v], geo.npix_v[v]
                if dt == torch.float64:
                    K.mean_ref(fr, geo.sbin, tsb[c0:c0 + n], tsa[c0:c0 + n])
                else:
                    K.bin_motion(fr, geo.sbin, tsum_bin=tsb[c0:c0 + n], tsum_adiff=tsa[c0:c0 + n])
        sums[i] = (tsb, tsa, int(frames[0].shape[0]))
    if segments is not None:
        return sums, nseg
    with timer("stage1.mean"):
        return stage1_finish(sums, nseg, geo, device)


def stage1_uses_reference(sbin):
    """Non-power-of-two bins: the reference's float64 means of means are not exact, so stage 1 reproduces
    its arithmetic (fm_mean_ref) instead of dividing exact integer sums."""
    return bool(sbin & (sbin - 1))


def stage1_finish(all_sums, nseg, geo, device):
    """float32 accumulation of the per-segment sums / means (any order of arrival) in the reference's segment
    order (process.py:84-86, 95-96)."""
    avgframe = torch.zeros(geo.p, dtype=torch.float32, device=device)
    avgmotion = torch.zeros(geo.p, dtype=torch.float32, device=device)
    for i in range(nseg):
        tsb, tsa, T = all_sums[i]
        if tsb.dtype == torch.float64:                      # segment means: avg32 = f32(f64(avg32) + mean)
            avgframe = (avgframe.double() + tsb).float()
            avgmotion = (avgmotion.double() + tsa).float()
        else:
            K.mean_update(avgframe, tsb, geo.sbin, T)
            K.mean_update(avgmotion, tsa, geo.sbin, max(T - 1, 1))
    K.mean_update(avgframe, None, geo.sbin, 1, finalize_div=nseg)
    K.mean_update(avgmotion, None, geo.sbin, 1, finalize_div=nseg)
    return avgframe, avgmotion


# ---------------------------------------------------------------------------------------------
# stage 2
# ---------------------------------------------------------------------------------------------
def stage2_plan(N, ncomps=NCOMPS):
    """facemap/process.py:134-141."""
    nt0 = min(CHUNK_SVD, N)
    nsegs = int(min(np.floor(BUDGET_SVD / nt0), np.floor(N / nt0)))
    nc = min(NC_CHUNK, nt0 - 1)
    if nsegs == 1:
        nc = min(ncomps, nt0 - 1)
    tf = np.floor(np.linspace(0, N - nt0 - 1, nsegs)).astype(int)
    return nt0, nsegs, nc, tf


def _bin_chunk(frames, geo, avgframe, avgmotion, mods, X, prev=None, last=None, energy=None, roi_energy=None,
               ormask=None, planes=False, skip=()):
    """K1 over every view of one fetched chunk into the modality matrices X[m] ([rows, p]).
    ormask: per-view paint masks (fullres.FullRes.ormask) OR-ed into the frames on load, or None.
    planes: X[m] are (hi, lo) uint8 plane pairs that receive the integer results (fm_bin_motion_planes)."""
    rows = None
    for v, fr in enumerate(frames):
        if v in skip:               # served by another kernel (Stage3: fm_motion_ref_planes)
            continue
        c0, n = geo.col0[v], geo.npix_v[v]
        rects = [geo.mot_rois[j][1:] for j in geo.rois_on_view[v]]
        if planes and mods:
            rows = K.bin_motion_planes(
                fr, geo.sbin, prev_sum=None if prev is None else prev[v], q_mot=X.get("mot"), q_mov=X.get("mov"),
                col0=c0, energy=None if energy is None else energy[v],
                rois=Rois.from_rects(rects) if (roi_energy is not None and rects) else None,
                roi_energy=None if (roi_energy is None or not rects) else roi_energy[v],
                last_sum=None if last is None else last[v], ormask=None if ormask is None else ormask[v])
            continue
        rows = K.bin_motion(
            fr, geo.sbin,
            prev_sum=None if prev is None else prev[v],
            avgmotion=avgmotion[c0:c0 + n] if "mot" in mods else None,
            avgframe=avgframe[c0:c0 + n] if "mov" in mods else None,
            x_mot=X.get("mot"), x_mov=X.get("mov"), col0=c0,
            energy=None if energy is None else energy[v],
            rois=Rois.from_rects(rects) if (roi_energy is not None and rects) else None,
            roi_energy=None if (roi_energy is None or not rects) else roi_energy[v],
            last_sum=None if last is None else last[v],
            ormask=None if ormask is None else ormask[v])
    return rows


def stage2_chunks(source, geo, avgframe, avgmotion, mods, fullSVD, ws, device, timer, chunks=None,
                  x_budget_bytes=24 << 30, Yt_out=None):
    """Per-chunk components.  Returns Yt[m][idx] = [n_chunks * k, p_idx] stacked blocks (idx 0 = full
    frame, 1+j = motion ROI j) for the chunk indices in `chunks` (default all).

    The chunks are processed in groups: K1 + centred Gram for every (chunk, modality, target) of the
    group, ONE batched eigensolve for all of them (they share n = nt0 - 1), then the left-vector
    GEMMs.  A group holds as many chunks as fit `x_budget_bytes` of chunk matrices."""
    N = source.nframes
    nt0, nsegs, nc, tf = stage2_plan(N)
    nroi = len(geo.mot_rois)
    todo = list(range(nsegs)) if chunks is None else list(chunks)
    sizes = [geo.p] + geo.roi_npix
    ks = [min(nc, s) for s in sizes]
    # Yt_out: the caller's buffers for the blocks (multi-GPU: windows on the gathered buffer, parallel.Shard)
    Yt = Yt_out if Yt_out is not None else {
        m: [K.alloc_matrix(len(todo) * ks[i], sizes[i], device) if (i > 0 or fullSVD) else None
            for i in range(1 + nroi)] for m in mods}
    n = nt0 - 1
    if n < 1 or not todo:
        return Yt, ks, nsegs
    targets = ([0] if fullSVD else []) + list(range(1, 1 + nroi))
    per_chunk = len(mods) * 4 * n * (K.round_up(geo.p, 32) + sum(K.round_up(s, 32) for s in geo.roi_npix))
    group = int(max(1, min(len(todo), x_budget_bytes // max(per_chunk, 1))))
    for g0 in range(0, len(todo), group):
        slots = list(range(g0, min(len(todo), g0 + group)))
        jobs = []            # (slot, modality, target, X view, mean, label)
        G = torch.empty((len(slots) * len(mods) * len(targets), n, n), dtype=torch.float64, device=device)
        for gi, slot in enumerate(slots):
            t = int(tf[todo[slot]])
            with timer("stage2.fetch"):
                frames = source.fetch(t, t + nt0, device)
                if slot + 1 < len(todo):
                    source.prefetch(int(tf[todo[slot + 1]]), int(tf[todo[slot + 1]]) + nt0, device)
            X = {m: ws.matrix(f"X.{m}.{gi}", n, geo.p) for m in mods}
            with timer("stage2.bin"):
                rows = _bin_chunk(frames, geo, avgframe, avgmotion, mods, X)
            assert rows == n
            for m in mods:
                for idx in targets:
                    if idx == 0:
                        Xt, label = X[m], "stage2.chunk"                       # process.py:246-259
                    else:                                                      # process.py:213-242
                        v, y0, y1, x0, x1 = geo.mot_rois[idx - 1]
                        Xt = ws.matrix(f"Xroi.{m}.{idx}.{gi}", n, sizes[idx])
                        with timer("stage2.roi_gather"):
                            K.gather_roi(X[m], n, geo.col0[v], int(geo.Lxb[v]), (y0, y1, x0, x1), Xt)
                        label = "stage2.roi"
                    mean, _ = centred_gram(ws, Xt, timer, label, out=G[len(jobs)])
                    jobs.append((slot, m, idx, Xt, mean, label))
        evec, nev = G, None
        if getattr(ws, "eig_mode", "cusolver") == "cluster" and not ws.eig_fp32 and 2 <= n <= K.EIGH_CLUSTER_NMAX:
            # hand-written solver: only the eigenpairs the targets keep, no host round trips (csrc/eigh_cluster.cu)
            nev = max(ks[idx] for (_, _, idx, _, _, _) in jobs)
            with timer("stage2.chunk.eigh"):
                lam, evec = ws.solver.eigh_topk_batched(G, nev, scratch=ws.raw("eigh.scratch", K.eigh_topk_scratch_bytes(
                    n, G.shape[0], nev), torch.uint8))
        else:
            with timer("stage2.chunk.syevd"):
                # the batched solver carries ~50 ms of fixed cost on B200 (15 x 999^2: 68 ms, 2 x 999^2: 53 ms)
                # against 12 ms per serial syevd: it pays off from 5 matrices on
                if ws.eig_fp32 or len(jobs) < BATCHED_EIG_MIN:
                    lam = torch.stack([ws.solver.syevd(G[i], use_fp32=ws.eig_fp32) for i in range(len(jobs))])
                else:
                    lam = ws.solver.syevd_batched(G)
        for i, (slot, m, idx, Xt, mean, label) in enumerate(jobs):
            kk = ks[idx]
            left_components(ws, Xt, evec[i], lam[i], mean, kk, Yt[m][idx][slot * kk:(slot + 1) * kk], timer, label, nev=nev)
    return Yt, ks, nsegs


def merge_cost(shape, ncomps, merge_eig):
    """Cost of one merge in units of cuSOLVER columns (~19 us each): the order of its eigenproblem — the smaller side
    of Yt [c, p] — unless the problem goes to the subspace solver (gram_eig: p >= c and subspace_eligible), which takes
    what cuSOLVER takes for ~1 500 columns (29 ms)."""
    c, p = shape
    n = min(c, p)
    k = min(ncomps, p - 1)
    if merge_eig == "subspace" and p >= c and subspace_eligible(c, k):
        return 1500
    return n


def deal_merges(sizes, world):
    """Which rank solves which merge: sizes = {target: cost (merge_cost)} -> {target: rank}, heaviest first
    onto the least loaded rank.  The cost of a direct solve is PROPORTIONAL TO n: cuSOLVER's syevd(x) spends about
    19 us per matrix column whatever the size (999 columns 12 ms, 3 750 columns 71 ms: profiles/r02_eig.md), not n^3.
    One target or one rank: nothing to deal ({})."""
    if world <= 1 or len(sizes) <= 1:
        return {}
    owner, load = {}, [0.0] * world
    for t in sorted(sizes, key=lambda t: -float(sizes[t])):          # sorted() is stable: ties keep the target order
        r = min(range(world), key=lambda q: load[q])
        owner[t] = r
        load[r] += float(sizes[t])
    return owner


def stage2_merge(Yt, ks, nsegs, geo, mods, fullSVD, ws, timer, ncomps=NCOMPS, nc=None, mot_on=True, shard=None):
    """Second-level SVD (process.py:264-295).  Returns per modality lists Ut, U, and Sv (the
    singular values of the LAST target processed, quirk Q3).
    shard (parallel.Shard, several ranks): every rank holds all stacked chunk components after the all-gather, and
    the merges of different targets (modalities x full frame / ROIs) are independent, so they are dealt across the
    ranks instead of replicated — heaviest first — and every result is broadcast by its owner (deterministic: the
    same bits as the replicated solve).  With ONE target (motion SVD of the full frame only) there is nothing to deal
    and every rank solves it."""
    # which targets get merged, and who solves them
    todo = []
    for m in mods:
        for idx, Y in enumerate(Yt[m]):
            if Y is not None and nsegs > 1 and idx < (1 if mot_on else 0) + (len(Yt[m]) - 1):
                todo.append((m, idx))
    world = 1 if shard is None else shard.world
    owner = deal_merges({t: merge_cost(tuple(Yt[t[0]][t[1]].shape), ncomps, getattr(ws, "merge_eig", "cusolver"))
                         for t in todo}, world)
    # the solves this rank owns come first (cuSOLVER blocks the host, and a broadcast waiting in the stream would block
    # the solves queued behind it), the broadcasts of all dealt targets follow in target order
    solved = {}
    for (m, idx) in todo:
        src = owner.get((m, idx))
        if src is None or src == shard.rank:
            Y = Yt[m][idx]
            solved[(m, idx)] = merge_components(ws, Y, min(ncomps, Y.shape[1] - 1), timer,
                                                "stage2.merge" if idx == 0 else "stage2.roi_merge")
    for (m, idx) in todo:
        src = owner.get((m, idx))
        if src is None:
            continue
        Y = Yt[m][idx]
        k, p_idx = min(ncomps, Y.shape[1] - 1), Y.shape[1]
        if src == shard.rank:
            Ut, U, sv_t = solved[(m, idx)]
        else:
            Ut, U, sv_t = K.alloc_matrix(k, p_idx, ws.device), None, torch.empty(k, dtype=torch.float32, device=ws.device)
        with timer("stage2.merge.bcast"):
            Ut, sv_t = shard.broadcast_merge(Ut, sv_t, src)
        if U is None:
            U = torch.empty((p_idx, k), dtype=torch.float32, device=ws.device)
            K.transpose(Ut, U)
        solved[(m, idx)] = (Ut, U, sv_t)
    out = {}
    for m in mods:
        Uts, Us = [], []
        sv = torch.zeros(500, dtype=torch.float32, device=ws.device)
        for idx, Y in enumerate(Yt[m]):
            if Y is None:                       # fullSVD off: placeholder like process.py:153-160
                Uts.append(None)
                Us.append(torch.zeros((0, 1), dtype=torch.float32, device=ws.device))
                continue
            p_idx = Y.shape[1]
            # quirk Q8: the reference's merge loop runs over range(len(U_mot)) and indexes both lists with it
            # (process.py:264-295); with the motion SVD off, U_mot holds only the ROI placeholders, so the movie targets
            # from that position on (all of them without ROIs, the last ROI otherwise) keep their stacked chunk blocks
            if (m, idx) in solved:
                Ut, U, sv = solved[(m, idx)]
            else:                               # quirk Q6 / Q8: not merged -> the chunks' U S as they are, Sv stays zero
                Ut = Y
                # the reference's full-frame buffer is nsegs * nc wide and only cut to its filled columns inside
                # the merge (process.py:146-151, 267): fewer binned pixels than nc leave zero components behind
                if idx == 0 and nc is not None and Y.shape[0] < nsegs * nc:
                    Ut = K.alloc_matrix(nsegs * nc, p_idx, ws.device, zero=True)
                    Ut[:Y.shape[0]].copy_(Y)
                U = torch.empty((p_idx, Ut.shape[0]), dtype=torch.float32, device=ws.device)
                K.transpose(Ut, U)
            Uts.append(Ut)
            Us.append(U)
        out[m] = (Uts, Us, sv)
    return out


# ---------------------------------------------------------------------------------------------
# stage 3
# ---------------------------------------------------------------------------------------------
def projection_rows_per_batch(p, k, budget_bytes=6 << 30, elem_bytes=4):
    """Rows per projection GEMM: whole waves of 128-row tiles over the 148 SMs, bounded by the
    X-buffer budget (elem_bytes per binned pixel: 4 for the float32 operand, 1 or 2 for byte planes)."""
    ntn = max(1, math.ceil(k / 256))
    rows_wave = max(1, NUM_SMS // ntn) * 128
    rows = 2 * rows_wave
    cap = max(128, (budget_bytes // (elem_bytes * max(p, 1))) // 128 * 128)
    return int(min(rows, cap))


PROJ_KB = 4096                       # K-block width of the blocked projection operands (fm_retile_u8)
PROJ_BLOCKED_MIN_PITCH = 262144      # plane row pitch (bytes) from which the projection reads K-blocked copies
MOTION_REF_BLOCKS = 148 * 4
MOTION_REF_SCRATCH_MAX = 1 << 30        # bytes of per-block scratch of fm_motion_ref per view


def motion_ref_blocks(npix, nleaves, sbin):
    """Persistent blocks of fm_motion_ref for one view: as many as fill the machine, fewer when the per-block
    scratch (two rows of binned values + the leaf sums, fm_motion_ref_scratch_bytes) would pass the budget —
    a 1280x1024 view at sbin 1 needs 10.5 MB per block."""
    per_block = int(nleaves) * 4 if sbin == 1 else (2 * int(npix) + int(nleaves)) * 8
    return int(max(8, min(MOTION_REF_BLOCKS, MOTION_REF_SCRATCH_MAX // per_block)))


def motion_needs_reference(sbin, npix):
    """True when `M += imbin_mot.sum(-1)` (process.py:460) is not decided by the exact integer energy: float64
    means of means for a non-power-of-two sbin (float32 rounding ties across views), float32 sums beyond 2^24
    for sbin == 1."""
    if sbin & (sbin - 1):
        return True
    return sbin == 1 and npix * 255 >= (1 << 24)


class Stage3:
    """Projection of every frame in [f0, f1) onto the masks (process.py:394-515), in two phases:

      bin      K1 over the frames -> X rows (+ integer energies); needs only the stage-1 averages
      project  X @ masks (K4) in whole-wave batches; needs the masks

    When the frames are resident in HBM and all of X fits a budget, `start_early()` enqueues the whole
    bin phase on a side stream right after stage 1, so it runs underneath the stage-2 eigensolves
    (which leave the SMs mostly idle) and only the GEMMs remain on the critical path.  Otherwise the
    two phases alternate batch by batch through one X buffer.

    Results: V[m] = list of [f1-f0, k] tensors (0 = full frame, 1+j = ROI j); integer energies
    E [nviews, f1-f0], E_roi [nroi, f1-f0], row i <-> frame f0+i holding sum |S_t - S_{t-1}| (row of
    frame 0 stays zero).  The reference's row conventions (Q1/Q2) are applied by the caller."""

    def __init__(self, source, geo, avgframe, avgmotion, mods, fullSVD, ws, device, timer, frame_range=None,
                 fetch_frames=2368, kmax=NCOMPS, fullres=None):
        self.source, self.geo, self.avgframe, self.avgmotion = source, geo, avgframe, avgmotion
        self.fullres = fullres          # facemap_b200.fullres.FullRes: pupil / blink / running on the same sweep
        self.mods, self.fullSVD, self.ws, self.device, self.timer = mods, fullSVD, ws, device, timer
        N = source.nframes
        self.f0, self.f1 = (0, N) if frame_range is None else frame_range
        self.n_out = self.f1 - self.f0
        self.nroi, self.nv = len(geo.mot_rois), len(geo.Ly)
        self.E = torch.zeros((self.nv, self.n_out), dtype=torch.int64, device=device)
        self.E_roi = torch.zeros((max(self.nroi, 1), self.n_out), dtype=torch.int64, device=device)
        self.carry = [[torch.zeros(n, dtype=torch.int32, device=device) for n in geo.npix_v] for _ in range(2)]
        self.flip = 0
        # integer-plane operands for fm_project_i8 (ws.proj_mode == "i8"); asked for but not servable is an error, not
        # a silent change of path
        self.i8 = ws.proj_mode == "i8" and bool(mods)
        if self.i8 and geo.sbin > 16:
            if getattr(ws, "proj_mode_explicit", True):
                raise FacemapB200Error("proj_mode 'i8': two byte planes hold block sums up to sbin 16")
            self.i8 = False                   # default mode: bins beyond 16 take the float32 operands
        self.ldq = K.round_up(geo.p, 16)
        # row pitches beyond this put every row of an operand tile on its own 2 MB page: project from K-blocked copies
        self.blocked = self.i8 and self.ldq > PROJ_BLOCKED_MIN_PITCH
        self.rows_batch = projection_rows_per_batch(geo.p, kmax, elem_bytes=(1 if geo.sbin == 1 else 2) if self.i8 else 4)
        # fetch plan: GEMM batches of <= rows_batch rows, each filled by several fetches
        self.have0 = self.f0 > 0                         # a shard starting mid-video carries a 1-frame halo
        self.batches, t, have = [], self.f0, self.have0
        while t < self.f1:
            fetches, filled = [], 0
            while t < self.f1 and filled < self.rows_batch:
                n = min(fetch_frames, self.f1 - t, self.rows_batch - filled + (0 if have else 1))
                rows = n if have else n - 1
                fetches.append((t, n, rows, have))
                filled += rows
                t += n
                have = True
            self.batches.append((fetches, filled))
        self.flat = [f for b, _ in self.batches for f in b]
        self.pos = 0
        self.first_row_frame = self.f0 if self.have0 else self.f0 + 1
        self.rows_total = max(0, self.f1 - self.first_row_frame)
        self.X_full = None
        self.early_done = None
        self._halo_done = False
        # views whose float32 motion accumulator the exact integer energy cannot decide (fm_motion_ref)
        self.Eref, self.ref_views = None, []
        if fullSVD and "mot" in mods:
            self.ref_views = [v for v in range(self.nv) if motion_needs_reference(geo.sbin, geo.npix_v[v])]
        if self.ref_views:
            self.Eref = torch.zeros((self.nv, self.n_out), dtype=torch.float64, device=device)
            self.ref_plan = {v: K.pairwise_plan(geo.npix_v[v], device, groups=geo.sbin == 1) for v in self.ref_views}
            self.ref_blocks = {v: motion_ref_blocks(geo.npix_v[v], self.ref_plan[v][0].numel(), geo.sbin)
                               for v in self.ref_views}
            self.ref_scratch = {v: K.motion_ref_scratch(geo.Ly[v], geo.Lx[v], geo.sbin, self.ref_plan[v][0].numel(),
                                                        self.ref_blocks[v], device) for v in self.ref_views}
            self.ref_last = {v: None for v in self.ref_views}
        # full-resolution views (sbin 1) that need the reference-arithmetic sums anyway: ONE pass writes those, the
        # operand plane and the integer energy (fm_motion_ref_planes) instead of K1 plus fm_motion_ref
        self.fused_views = [v for v in self.ref_views
                            if self.i8 and geo.sbin == 1 and list(mods) == ["mot"] and not geo.rois_on_view[v]]

    def _operands(self, tag, rows):
        """per-modality stage-3 operands for `rows` difference rows: float32 [rows, p], or (hi, lo) byte planes"""
        if not self.i8:
            return {m: self.ws.matrix(tag + m, rows, self.geo.p) for m in self.mods}
        out = {}
        for m in self.mods:
            buf = self.ws.raw(tag + "q." + m, 2 * max(rows, 1) * self.ldq, torch.uint8).view(2, max(rows, 1), self.ldq)
            out[m] = (buf[0, :rows, :self.geo.p] if self.geo.sbin > 1 else None, buf[1, :rows, :self.geo.p])
        return out

    def _rows_of(self, X, r0, r1):
        if not self.i8:
            return {m: X[m][r0:r1] for m in self.mods}
        return {m: (None if X[m][0] is None else X[m][0][r0:r1], X[m][1][r0:r1]) for m in self.mods}

    # ---- bin phase -------------------------------------------------------------------------------
    def _seed_halo(self):
        if self.have0 and not self._halo_done:
            with self.timer("stage3.fetch"):
                halo = self.source.fetch(self.f0 - 1, self.f0, self.device)
            if self.fullres is not None:
                with self.timer("stage3.fullres"):
                    self.fullres.seed(halo)
            for v, fr in enumerate(halo):
                K.bin_motion(fr, self.geo.sbin, last_sum=self.carry[self.flip][v],
                             ormask=None if self.fullres is None else self.fullres.ormask[v])
            for v in self.ref_views:
                self.ref_last[v] = halo[v][-1].clone()
        self._halo_done = True

    def _bin_fetches(self, fetches, X, row0):
        """K1 for a list of fetches; rows land in X[m][row0 + ...]."""
        geo, filled = self.geo, 0
        for (t, n, rows, have) in fetches:
            with self.timer("stage3.fetch"):
                frames = self.source.fetch(t, t + n, self.device)
                if self.pos + 1 < len(self.flat):
                    nxt = self.flat[self.pos + 1]
                    self.source.prefetch(nxt[0], nxt[0] + nxt[1], self.device)
            self.pos += 1
            if self.fullres is not None:
                # full-resolution ROIs read the raw frames; K1 ORs their paint in as it loads (process.py:543)
                with self.timer("stage3.fullres"):
                    self.fullres.chunk(frames, t)
            first_frame = t if have else t + 1
            with self.timer("stage3.bin"):
                r = row0 + filled
                Xv = self._rows_of(X, r, r + rows) if rows > 0 else {m: None for m in self.mods}
                o = first_frame - self.f0
                energy = [self.E[v, o:o + rows] for v in range(self.nv)]
                roi_e = []
                for v in range(self.nv):
                    js = geo.rois_on_view[v]
                    roi_e.append(None)
                    if js:      # ROI energies of one view must be contiguous [n_on_view, rows]
                        roi_e[v] = torch.zeros((len(js), max(rows, 1)), dtype=torch.int64, device=self.device)
                _bin_chunk(frames, geo, self.avgframe, self.avgmotion, self.mods if rows > 0 else [], Xv,
                           prev=self.carry[self.flip] if have else None, last=self.carry[1 - self.flip],
                           energy=energy if rows > 0 else None,
                           roi_energy=roi_e if (self.nroi and rows > 0) else None,
                           ormask=None if self.fullres is None else self.fullres.ormask, planes=self.i8,
                           skip=self.fused_views)
                if self.nroi and rows > 0:
                    for v in range(self.nv):
                        for q, j in enumerate(geo.rois_on_view[v]):
                            self.E_roi[j, o:o + rows] = roi_e[v][q, :rows]
                for v in self.ref_views:
                    if rows > 0 and v in self.fused_views:
                        K.motion_ref_planes(frames[v], self.ref_plan[v], self.ref_scratch[v], self.ref_blocks[v],
                                            self.Eref[v, o:o + rows], Xv["mot"][1], col0=geo.col0[v], energy=energy[v],
                                            prev_frame=self.ref_last[v] if have else None,
                                            ormask=None if self.fullres is None else self.fullres.ormask[v])
                    elif rows > 0:
                        K.motion_ref(frames[v], geo.sbin, self.ref_plan[v], self.ref_scratch[v], self.ref_blocks[v],
                                     self.Eref[v, o:o + rows], prev_frame=self.ref_last[v] if have else None,
                                     ormask=None if self.fullres is None else self.fullres.ormask[v])
                    self.ref_last[v] = frames[v][-1].clone()
            self.flip = 1 - self.flip
            filled += rows
        return filled

    def start_early(self, budget_bytes=24 << 30):
        """Enqueue the whole bin phase on a side stream (see class docstring).  Returns True if taken."""
        if self.rows_total == 0 or not self.mods or not self.source.resident(self.f0, self.f1):
            return False
        need = len(self.mods) * self.rows_total * (2 * self.ldq if self.i8 else max(32, K.round_up(self.geo.p, 32)) * 4)
        if need > budget_bytes or need > 0.5 * K.free_memory(self.device, need):
            return False
        self.X_full = self._operands("X3full.", self.rows_total)
        main = torch.cuda.current_stream()
        side = K.shared_stream(self.device, "stage3.early")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._seed_halo()
            row0 = 0
            # one event per batch: project() starts on a batch as soon as it is binned, so that with frames still
            # arriving from the host the projections, and their copies back to the host, spread over the upload
            self.early_events = []
            for fetches, filled in self.batches:
                row0 += self._bin_fetches(fetches, self.X_full, row0)
                ev = torch.cuda.Event()
                ev.record(side)
                self.early_events.append(ev)
            self.early_done = torch.cuda.Event()
            self.early_done.record(side)
        self._side = side
        return True

    # ---- project phase ---------------------------------------------------------------------------
    def project(self, masks, sink=None, V_host=None):
        """sink: optional host sink (process._HostSink); the projections are then copied to pinned host
        arrays batch by batch on its stream and returned as CPU tensors."""
        geo, mods, ws, timer, device = self.geo, self.mods, self.ws, self.timer, self.device
        V, V_host = {}, dict(V_host or {})
        for m in mods:
            V[m] = [None if Ut is None else torch.zeros((self.n_out, Ut.shape[0]), dtype=torch.float32, device=device)
                    for Ut in masks[m][0]]
            if sink is not None:
                if m not in V_host:
                    V_host[m] = [None if v is None else sink.alloc(v.shape) for v in V[m]]
                for h in V_host[m]:
                    if h is not None and self.first_row_frame > self.f0:
                        h[:self.first_row_frame - self.f0].zero_()      # frame 0 has no difference row
        # the masks are the B operand of every projection GEMM: split them once (lo plane), TMA stages both
        with timer("stage3.split_masks"):
            if self.i8:
                # digits of the masks and the rank-one term of the subtracted average, once per run
                # (targets: 0 = full frame, 1 + j = motion ROI j with the ROI's columns of the average)
                avg = {"mot": self.avgmotion.double(), "mov": self.avgframe.double()}
                roi_cols = []
                for (v, y0, y1, x0, x1) in geo.mot_rois:
                    yy = torch.arange(y0, y1, device=device)[:, None] * int(geo.Lxb[v])
                    roi_cols.append((int(geo.col0[v]) + yy + torch.arange(x0, x1, device=device)[None, :]).reshape(-1))
                Ut_q = {m: [None if Ut is None else K.slice_rows_i8(Ut) for Ut in masks[m][0]] for m in mods}
                Ut_bias = {m: [None if Ut is None else K.row_dot(Ut, avg[m] if i == 0 else avg[m][roi_cols[i - 1]].contiguous())
                               for i, Ut in enumerate(masks[m][0])] for m in mods}
                Ut_lo = None
            else:
                Ut_lo = {m: [None if Ut is None else K.split_lo(Ut) for Ut in masks[m][0]] for m in mods}
        if self.X_full is not None:
            X = self.X_full
        else:
            X = self._operands("X3.", self.rows_batch)
            self.source.begin_sweep()
            self._seed_halo()
        row0 = 0
        Ut_qb = {}                      # K-blocked copies of the mask digits (wide views)
        for ib, (fetches, filled) in enumerate(self.batches):
            if self.X_full is None:
                self._bin_fetches(fetches, X, 0)
                base = 0
            else:
                torch.cuda.current_stream().wait_event(self.early_events[ib])
                base = row0
            if filled == 0:
                continue
            o = self.first_row_frame + row0 - self.f0
            for m in mods:
                Uts = masks[m][0]
                if self.i8:
                    alpha = 1.0 / (geo.sbin * geo.sbin)
                    hi, lo = self._rows_of(X, base, base + filled)[m]
                    if self.fullSVD and Uts[0] is not None and self.blocked:
                        # wide view: both operands in the K-blocked layout (a 128-row tile = 128 * PROJ_KB contiguous
                        # bytes instead of one 2 MB page per row; fm_project_i8_blocked)
                        with timer("stage3.project.retile"):
                            nblk = -(-geo.p // PROJ_KB)
                            if m not in Ut_qb:
                                Ut_qb[m] = tuple(K.retile_u8(q, PROJ_KB) for q in Ut_q[m][0][:3])
                            Db = []
                            for tag, pl in (("hi", hi), ("lo", lo)):
                                Db.append(None if pl is None else K.retile_u8(pl, PROJ_KB, out=ws.raw(
                                    "X3blk." + tag, nblk * filled * PROJ_KB, torch.uint8).view(nblk, filled, PROJ_KB)))
                        with timer("stage3.project.mma"):
                            K.project_i8_blocked(tuple(Db), Ut_qb[m], Ut_q[m][0][3], alpha, Ut_bias[m][0],
                                                 V[m][0][o:o + filled], geo.p, impl=ws.proj_i8_impl)
                        if sink is not None:
                            sink.copy(V[m][0][o:o + filled], V_host[m][0][o:o + filled])
                    elif self.fullSVD and Uts[0] is not None:
                        with timer("stage3.project.mma"):
                            K.project_i8((hi, lo), Ut_q[m][0][:3], Ut_q[m][0][3], alpha, Ut_bias[m][0], V[m][0][o:o + filled],
                                         impl=ws.proj_i8_impl)
                        if sink is not None:
                            sink.copy(V[m][0][o:o + filled], V_host[m][0][o:o + filled])
                    for j, (v, y0, y1, x0, x1) in enumerate(geo.mot_rois):
                        ldr = K.round_up(geo.roi_npix[j], 16)
                        for r0 in range(0, filled, 32768):
                            r1 = min(filled, r0 + 32768)
                            buf = ws.raw("X3roi.q", 2 * (r1 - r0) * ldr, torch.uint8).view(2, r1 - r0, ldr)
                            with timer("stage3.roi_gather"):
                                for src, dst in ((hi, buf[0]), (lo, buf[1])):
                                    if src is not None:
                                        K.gather_roi_u8(src[r0:r1], r1 - r0, geo.col0[v], int(geo.Lxb[v]), (y0, y1, x0, x1), dst)
                            with timer("stage3.roi_project.mma"):
                                K.project_i8((None if hi is None else buf[0][:, :geo.roi_npix[j]], buf[1][:, :geo.roi_npix[j]]),
                                             Ut_q[m][1 + j][:3], Ut_q[m][1 + j][3], alpha, Ut_bias[m][1 + j],
                                             V[m][1 + j][o + r0:o + r1], impl=ws.proj_i8_impl)
                            if sink is not None:
                                sink.copy(V[m][1 + j][o + r0:o + r1], V_host[m][1 + j][o + r0:o + r1])
                    continue
                Xm = X[m][base:base + filled]
                if self.fullSVD and Uts[0] is not None:
                    ws.gemm(Xm, Uts[0], V[m][0][o:o + filled], label="stage3.project", B_lo=Ut_lo[m][0], impl=ws.proj_impl,
                            chain=KCHAIN_PROJ if ws.proj_impl == 9 else None)
                    if sink is not None:
                        sink.copy(V[m][0][o:o + filled], V_host[m][0][o:o + filled])
                for j, (v, y0, y1, x0, x1) in enumerate(geo.mot_rois):
                    for r0 in range(0, filled, 32768):
                        r1 = min(filled, r0 + 32768)
                        Xr = ws.matrix("X3roi", r1 - r0, geo.roi_npix[j])
                        with timer("stage3.roi_gather"):
                            K.gather_roi(Xm[r0:r1], r1 - r0, geo.col0[v], int(geo.Lxb[v]), (y0, y1, x0, x1), Xr)
                        ws.gemm(Xr, Uts[1 + j], V[m][1 + j][o + r0:o + r1], label="stage3.roi_project",
                                B_lo=Ut_lo[m][1 + j], impl=ws.proj_impl,
                                chain=KCHAIN_PROJ if ws.proj_impl == 9 else None)
                        if sink is not None:
                            sink.copy(V[m][1 + j][o + r0:o + r1], V_host[m][1 + j][o + r0:o + r1])
            row0 += filled
        if self.X_full is not None:
            torch.cuda.current_stream().wait_event(self.early_done)     # the energies of the whole pass
        if sink is not None:
            return V_host, self.E, self.E_roi
        return V, self.E, self.E_roi


def stage3_project(source, geo, avgframe, avgmotion, masks, mods, fullSVD, ws, device, timer, frame_range=None,
                   fetch_frames=2368):
    """One-shot form of Stage3 (bin and project alternating batch by batch)."""
    kmax = max([Ut.shape[0] for m in mods for Ut in masks[m][0] if Ut is not None] + [1])
    return Stage3(source, geo, avgframe, avgmotion, mods, fullSVD, ws, device, timer, frame_range, fetch_frames,
                  kmax).project(masks)
