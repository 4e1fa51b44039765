"""The integer projection path on the CPU: (1) the arithmetic fm_project_i8 implements (tests/i8_model.py) against the
exact float64 projection, with the float32 GEMM the reference uses (facemap/process.py:485) beside it; (2) the int32
bound of a 16 384-K chain at the extreme digits; (3) Stage3's orchestration in FM_PROJ_MODE=i8 with NumPy stand-ins
for the kernels that restate their contracts (the kernels themselves have not run on a GPU yet:
tests/test_project_i8_gpu.py is opt-in)."""
import numpy as np
import pytest
import torch

from facemap_b200 import svd

from . import i8_model as M8


def _masks(rs, k, p):
    """orthonormal-ish rows with a graded dynamic range, like the masks of a merge SVD"""
    U = rs.standard_normal((k, p)).astype(np.float32) * np.exp(rs.uniform(-3, 0, size=(1, p))).astype(np.float32)
    return (U / np.linalg.norm(U, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("sbin,p", [(4, 3000), (1, 2500), (16, 700)])
def test_model_matches_exact_projection(sbin, p):
    rs = np.random.RandomState(sbin)
    T, k = 64, 40
    D = rs.randint(0, 255 * sbin * sbin + 1, size=(T, p))
    D[3] = 255 * sbin * sbin                                   # a saturated frame difference
    D[4] = 0
    U = _masks(rs, k, p)
    avg = (D.mean(axis=0) / sbin ** 2).astype(np.float32)
    alpha = 1.0 / sbin ** 2
    got = M8.project_i8_model(D, U, alpha, avg).astype(np.float64)
    X64 = D.astype(np.float64) * alpha - avg.astype(np.float64)
    want = X64 @ U.astype(np.float64).T
    # the sum of magnitudes behind every output: what a float32 GEMM's rounding scales with
    scale = (np.abs(D) * alpha + np.abs(avg).astype(np.float64)) @ np.abs(U).astype(np.float64).T
    # digits: 2^-24 of a mask row's bound per entry, unbiased; float32 store: 2^-24 of the result
    assert (np.abs(got - want) / scale).max() < 1e-7
    # the reference's own arithmetic: float32 operand, float32 GEMM
    ref = ((D * alpha - avg.astype(np.float64)).astype(np.float32) @ U.T).astype(np.float64)
    assert np.abs(got - want).max() <= 4 * np.abs(ref - want).max() + 1e-12


def test_chain_bound_at_extreme_digits():
    """two u8 x u8 products share an accumulator: 2 * 255^2 * 16 384 < 2^31; a chain twice as long would not fit"""
    K = M8.KCHAIN
    hi = np.full((1, K), 255, np.uint8)
    lo = np.full((1, K), 255, np.uint8)
    q0 = np.full((1, K), -128, np.int8)
    q1 = np.full((1, K), 255, np.uint8)
    q2 = np.full((1, K), 255, np.uint8)
    cls = M8.class_sums([hi, lo], (q0, q1, q2))
    assert cls.shape == (1, 4, 1, 1)
    assert cls[0, 2, 0, 0] == 2 * 255 * 255 * K < 2 ** 31
    assert 2 * 255 * 255 * (2 * K) >= 2 ** 31
    with pytest.raises(AssertionError):
        M8.class_sums([np.tile(hi, 2), np.tile(lo, 2)], (np.tile(q0, 2), np.tile(q1, 2), np.tile(q2, 2)), chain=2 * K)
    # int64 across chains: 2^22 K at the extreme leading product stays below 2^63
    assert 255 * 128 * (1 << 22) * (1 << 24) < 2 ** 63


class FakeKernels:
    """contracts of the kernels Stage3 calls in i8 mode, on CPU tensors"""

    def __init__(self):
        self.log = []

    @staticmethod
    def round_up(x, m):
        return (int(x) + m - 1) // m * m

    def bin_motion_planes(self, frames, sbin, *, prev_sum=None, q_mot=None, q_mov=None, col0=0, energy=None, rois=None,
                          roi_energy=None, last_sum=None, ormask=None):
        fr = frames.numpy() if ormask is None else frames.numpy() | ormask.numpy()[None]
        S = M8.block_sums(fr, sbin)
        npix = S.shape[1]
        prev = S[:1] if prev_sum is None else prev_sum.numpy().astype(np.int64)[None, :]
        cur = S[1:] if prev_sum is None else S
        d = np.abs(np.diff(np.concatenate([prev, cur]), axis=0))
        rows = d.shape[0]
        for pair, val in ((q_mot, d), (q_mov, cur)):
            if pair is not None:
                hi, lo = pair
                assert lo.dtype == torch.uint8 and lo.stride(0) % 16 == 0 and (hi is None or hi.stride(0) == lo.stride(0))
                h, l = M8.planes_of(val)
                lo[:rows, col0:col0 + npix] = torch.from_numpy(l)
                if hi is not None:
                    hi[:rows, col0:col0 + npix] = torch.from_numpy(h)
                else:
                    assert not h.any()
        if energy is not None:
            energy[:rows] = torch.from_numpy(d.sum(axis=1))
        if roi_energy is not None and rois is not None:
            img = d.reshape(rows, -1, frames.shape[2] // sbin)
            for q in range(rois.n):
                roi_energy.view(rois.n, -1)[q, :rows] = torch.from_numpy(
                    img[:, rois.y0[q]:rois.y1[q], rois.x0[q]:rois.x1[q]].sum(axis=(1, 2)))
        if last_sum is not None:
            last_sum.copy_(torch.from_numpy(S[-1].astype(np.int32)))
        self.log.append(("bin", rows))
        return rows

    @staticmethod
    def gather_roi_u8(X, rows, col0, Lxb, rect, out):
        assert X.dtype == torch.uint8 and out.dtype == torch.uint8
        y0, y1, x0, x1 = rect
        cols = [col0 + y * Lxb + x for y in range(y0, y1) for x in range(x0, x1)]
        out[:rows, :len(cols)] = X[:rows, cols]

    @staticmethod
    def slice_rows_i8(X, out=None):
        q0, q1, q2, exps = M8.mask_digits(X.numpy())
        return torch.from_numpy(q0), torch.from_numpy(q1), torch.from_numpy(q2), torch.from_numpy(exps)

    @staticmethod
    def row_dot(Y, w):
        assert w.dtype == torch.float64
        return torch.from_numpy(Y.numpy().astype(np.float64) @ w.numpy())

    def project_i8(self, D, Q, exps, alpha, bias, V, impl=0):
        hi, lo = D
        Dv = lo.numpy().astype(np.int64) + (0 if hi is None else 256 * hi.numpy().astype(np.int64))
        planes = [lo.numpy()] if hi is None else [hi.numpy(), lo.numpy()]
        cls = M8.class_sums(planes, tuple(q.numpy() for q in Q))
        nc = cls.shape[1]
        acc = sum(cls[c, j] * (1 << (8 * (nc - 1 - j))) for c in range(cls.shape[0]) for j in range(nc))
        v = acc.astype(np.float64) * (alpha * np.exp2(exps.numpy().astype(np.float64) - 23))[None, :] - bias.numpy()[None, :]
        V[:Dv.shape[0], :v.shape[1]] = torch.from_numpy(v.astype(np.float32))
        self.log.append(("project", Dv.shape[0]))
        return V


class ArraySource:
    def __init__(self, views):
        self.views = views
        self.nframes = views[0].shape[0]

    def fetch(self, t0, t1, device):
        return [torch.from_numpy(v[t0:t1]) for v in self.views]

    def prefetch(self, t0, t1, device):
        pass

    def resident(self, t0, t1):
        return False

    def begin_sweep(self):
        self.swept = True


@pytest.mark.parametrize("sbin,mods", [(4, ["mot"]), (2, ["mot", "mov"]), (1, ["mot"])])
def test_stage3_i8_orchestration(monkeypatch, sbin, mods):
    rs = np.random.RandomState(10 + sbin)
    N, k = 61, 12
    views = [rs.randint(0, 256, size=(N, 24, 36)).astype(np.uint8), rs.randint(0, 256, size=(N, 16, 20)).astype(np.uint8)]
    ry, rx = (np.arange(1, 3), np.arange(2, 5)) if sbin == 4 else (np.arange(2, 6), np.arange(3, 9))
    roi = {"rind": 1, "ivid": 1, "yrange_bin": ry, "xrange_bin": rx}
    geo = svd.Geometry([24, 16], [36, 20], sbin, [roi])
    S = np.concatenate([M8.block_sums(v, sbin) for v in views], axis=1)
    d = np.abs(np.diff(S, axis=0))
    avgmotion = (d.mean(axis=0) / sbin ** 2).astype(np.float32)
    avgframe = (S.mean(axis=0) / sbin ** 2).astype(np.float32)
    U = {m: _masks(rs, k, geo.p) for m in mods}
    Ur = {m: _masks(rs, 5, geo.roi_npix[0]) for m in mods}
    cols = [geo.col0[1] + y * int(geo.Lxb[1]) + x for y in ry for x in rx]
    fake = FakeKernels()
    monkeypatch.setattr(svd, "K", fake)
    ws = object.__new__(svd.SvdWorkspace)
    ws.device, ws._bufs, ws._ws = torch.device("cpu"), {}, None
    ws.proj_mode, ws.proj_impl, ws.proj_i8_impl = "i8", 0, 0
    timer = svd.StageTimer(enabled=False)
    st = svd.Stage3(ArraySource(views), geo, torch.from_numpy(avgframe), torch.from_numpy(avgmotion), mods, True, ws,
                    torch.device("cpu"), timer, fetch_frames=17)
    assert st.i8
    masks = {m: ([torch.from_numpy(U[m]), torch.from_numpy(Ur[m])], None) for m in mods}
    V, E, E_roi = st.project(masks)
    assert np.array_equal(E.numpy()[:, 1:].sum(axis=0), d.sum(axis=1)) and not E.numpy()[:, 0].any()
    assert np.array_equal(E_roi.numpy()[0, 1:], d[:, cols].sum(axis=1))
    X = {"mot": d / sbin ** 2 - avgmotion.astype(np.float64), "mov": S[1:] / sbin ** 2 - avgframe.astype(np.float64)}
    for m in mods:
        got = V[m][0].numpy().astype(np.float64)
        want = X[m] @ U[m].astype(np.float64).T
        assert not got[0].any()                                 # frame 0 has no difference row
        assert np.abs(got[1:] - want).max() <= 2e-6 * np.abs(want).max()
        want_r = X[m][:, cols] @ Ur[m].astype(np.float64).T     # the motion ROI: gathered planes, its own masks
        assert np.abs(V[m][1].numpy()[1:] - want_r).max() <= 2e-6 * np.abs(want_r).max()
    assert [r for op, r in fake.log if op == "bin"][:2] == [16, 16]       # per view: 17 frames, no predecessor -> 16 rows


def test_stage3_i8_refuses_what_it_cannot_serve():
    geo = svd.Geometry([32], [32], 32, None)
    ws = object.__new__(svd.SvdWorkspace)
    ws.device, ws._bufs, ws._ws, ws.proj_mode, ws.proj_impl, ws.proj_i8_impl = torch.device("cpu"), {}, None, "i8", 0, 0
    src = ArraySource([np.zeros((5, 32, 32), np.uint8)])
    with pytest.raises(svd.FacemapB200Error):
        svd.Stage3(src, geo, None, None, ["mot"], True, ws, torch.device("cpu"), svd.StageTimer(enabled=False))
    # the DEFAULT projection mode is "i8" too, but not asked for by name: bins beyond 16 take the float32 operands
    ws.proj_mode_explicit = False
    s3 = svd.Stage3(src, geo, None, None, ["mot"], True, ws, torch.device("cpu"), svd.StageTimer(enabled=False))
    assert s3.i8 is False


class FloatFakeKernels:
    """contracts of the kernels Stage3 calls on its DEFAULT path (float32 operands, fm_gemm_tn), on CPU tensors"""

    round_up = staticmethod(FakeKernels.round_up)

    def __init__(self):
        self.gemms = []

    @staticmethod
    def bin_motion(frames, sbin, *, prev_sum=None, avgmotion=None, avgframe=None, x_mot=None, x_mov=None, col0=0,
                   energy=None, rois=None, roi_energy=None, last_sum=None, tsum_bin=None, tsum_adiff=None, ormask=None):
        S = M8.block_sums(frames.numpy(), sbin)
        npix = S.shape[1]
        prev = S[:1] if prev_sum is None else prev_sum.numpy().astype(np.int64)[None, :]
        cur = S[1:] if prev_sum is None else S
        d = np.abs(np.diff(np.concatenate([prev, cur]), axis=0))
        rows = d.shape[0]
        if x_mot is not None:
            x_mot[:rows, col0:col0 + npix] = torch.from_numpy((d / sbin ** 2 - avgmotion.numpy().astype(np.float64)).astype(np.float32))
        if x_mov is not None:
            x_mov[:rows, col0:col0 + npix] = torch.from_numpy((cur / sbin ** 2 - avgframe.numpy().astype(np.float64)).astype(np.float32))
        if energy is not None:
            energy[:rows] = torch.from_numpy(d.sum(axis=1))
        if roi_energy is not None and rois is not None:
            Lxb = frames.shape[2] // sbin
            img = d.reshape(rows, -1, Lxb)
            for q in range(rois.n):
                roi_energy.view(rois.n, -1)[q, :rows] = torch.from_numpy(
                    img[:, rois.y0[q]:rois.y1[q], rois.x0[q]:rois.x1[q]].sum(axis=(1, 2)))
        if last_sum is not None:
            last_sum.copy_(torch.from_numpy(S[-1].astype(np.int32)))
        return rows

    @staticmethod
    def split_lo(X, out=None):
        return torch.zeros_like(X)

    @staticmethod
    def gather_roi(X, rows, col0, Lxb, rect, out):
        y0, y1, x0, x1 = rect
        cols = [col0 + y * Lxb + x for y in range(y0, y1) for x in range(x0, x1)]
        out[:rows] = X[:rows, cols]

    def gemm_tn(self, A, B, C_out=None, *, rscale=None, rbias=None, ksplit=1, workspace=None, upper_only=False, impl=0,
                B_lo=None, A_lo=None):
        assert rscale is None and rbias is None
        P = torch.from_numpy((A.numpy().astype(np.float64) @ B.numpy().astype(np.float64).T).astype(np.float32))
        if ksplit > 1:                                      # narrow outputs: K partials, summed by splitk_reduce
            workspace[:ksplit].zero_()
            workspace[0, :A.shape[0], :B.shape[0]] = P
        else:
            C_out[:A.shape[0], :B.shape[0]] = P
        self.gemms.append((A.shape[0], B.shape[0], A.shape[1]))

    @staticmethod
    def splitk_reduce(workspace, ksplit, C_out, *, rscale=None, rbias=None):
        C_out.copy_(workspace[:ksplit].sum(dim=0)[:C_out.shape[0], :C_out.shape[1]])


@pytest.mark.parametrize("fetch_frames", [17, 1000])
def test_stage3_default_float_orchestration(monkeypatch, fetch_frames):
    """the default path through Stage3 (the one measured on the B200) with stand-in kernels: two views, motion and
    movie operands, one motion ROI — guards the operand bookkeeping shared with the plane mode"""
    rs = np.random.RandomState(77)
    N, k, sbin = 61, 12, 2
    views = [rs.randint(0, 256, size=(N, 24, 36)).astype(np.uint8), rs.randint(0, 256, size=(N, 16, 20)).astype(np.uint8)]
    roi = {"rind": 1, "ivid": 1, "yrange_bin": np.arange(2, 6), "xrange_bin": np.arange(3, 9)}
    geo = svd.Geometry([24, 16], [36, 20], sbin, [roi])
    S = np.concatenate([M8.block_sums(v, sbin) for v in views], axis=1)
    d = np.abs(np.diff(S, axis=0))
    avgmotion = (d.mean(axis=0) / sbin ** 2).astype(np.float32)
    avgframe = (S.mean(axis=0) / sbin ** 2).astype(np.float32)
    mods = ["mot", "mov"]
    U = {m: _masks(rs, k, geo.p) for m in mods}
    Ur = {m: _masks(rs, 5, geo.roi_npix[0]) for m in mods}
    fake = FloatFakeKernels()
    monkeypatch.setattr(svd, "K", fake)
    ws = object.__new__(svd.SvdWorkspace)
    ws.device, ws._bufs, ws._ws = torch.device("cpu"), {}, None
    ws.proj_mode, ws.proj_impl, ws.gemm_impl = "tf32x3", 0, 0
    ws.timer = svd.StageTimer(enabled=False)
    st = svd.Stage3(ArraySource(views), geo, torch.from_numpy(avgframe), torch.from_numpy(avgmotion), mods, True, ws,
                    torch.device("cpu"), ws.timer, fetch_frames=fetch_frames)
    assert not st.i8
    masks = {m: ([torch.from_numpy(U[m]), torch.from_numpy(Ur[m])], None) for m in mods}
    V, E, E_roi = st.project(masks)
    assert np.array_equal(E.numpy()[:, 1:].sum(axis=0), d.sum(axis=1))
    c0, Lxb = geo.col0[1], int(geo.Lxb[1])
    cols = [c0 + y * Lxb + x for y in range(2, 6) for x in range(3, 9)]
    assert np.array_equal(E_roi.numpy()[0, 1:], d[:, cols].sum(axis=1))
    X = {"mot": d / sbin ** 2 - avgmotion.astype(np.float64), "mov": S[1:] / sbin ** 2 - avgframe.astype(np.float64)}
    for m in mods:
        want = X[m] @ U[m].astype(np.float64).T
        assert np.abs(V[m][0].numpy()[1:] - want).max() <= 2e-6 * np.abs(want).max() and not V[m][0].numpy()[0].any()
        want_r = X[m][:, cols] @ Ur[m].astype(np.float64).T
        assert np.abs(V[m][1].numpy()[1:] - want_r).max() <= 2e-6 * np.abs(want_r).max()


def test_dry_run_of_the_opt_in_gpu_tests(monkeypatch):
    """the bodies of tests/test_project_i8_gpu.py on the CPU with stand-in kernels: their shapes, dtypes, pitches and
    expected values are consistent with the model, so a GPU minute is never spent on a bug of the test itself"""
    import facemap_b200.kernels as Kreal

    from . import test_project_i8_gpu as T

    fake = FakeKernels()

    def alloc_matrix(rows, cols, device, zero=False):
        return torch.zeros((rows, -(-cols // 32) * 32), dtype=torch.float32)[:, :cols]

    def project_i8(D, Q, exps, alpha, bias, V, impl=0):
        hi, lo = D
        planes = [lo.numpy()] if hi is None else [hi.numpy(), lo.numpy()]
        cls = M8.class_sums(planes, tuple(q.numpy() for q in Q))
        nc = cls.shape[1]
        acc = sum(cls[c, j] * (1 << (8 * (nc - 1 - j))) for c in range(cls.shape[0]) for j in range(nc))
        v = acc.astype(np.float64) * (np.float64(alpha) * np.exp2(exps.numpy().astype(np.float64) - 23))[None, :]
        if bias is not None:
            v = v - bias.numpy()[None, :]
        V[:lo.shape[0], :v.shape[1]] = torch.from_numpy(v.astype(np.float32))

    monkeypatch.setattr(Kreal, "alloc_matrix", alloc_matrix)
    monkeypatch.setattr(Kreal, "slice_rows_i8", FakeKernels.slice_rows_i8)
    monkeypatch.setattr(Kreal, "row_dot", FakeKernels.row_dot)
    monkeypatch.setattr(Kreal, "project_i8", project_i8)
    monkeypatch.setattr(Kreal, "bin_motion_planes", fake.bin_motion_planes)
    cpu = torch.device("cpu")
    for (M, N, K, sbin) in [(128, 128, 64, 4), (77, 130, 1000, 1), (5, 7, 33, 16), (260, 250, 16384 + 64, 2)]:
        T.test_project_i8_bit_exact_against_model(cpu, M, N, K, sbin, 1)
    T.test_project_i8_extreme_digits_fill_the_int32_chain(cpu, 0)
    for (sbin, H, W) in [(1, 40, 64), (4, 64, 96), (3, 45, 60), (4, 37, 53)]:
        T.test_bin_motion_planes_exact(cpu, sbin, H, W)
