"""Host-side logic of the product (geometry, plans, frame sources, output conventions) against
the oracle restatement and hand-computed cases.  CPU only."""
import os

import numpy as np
import pytest

from facemap_b200 import utils
from oracle import facemap_oracle as orc
from tests.helpers import motion_roi


def test_binned_inds_and_placement_match_oracle():
    rng = np.random.RandomState(0)
    for trial in range(200):
        nv = rng.randint(1, 7)
        Ly = list(rng.randint(8, 200, nv))
        Lx = list(rng.randint(8, 200, nv))
        sbin = int(rng.randint(1, 8))
        a = utils.binned_inds(Ly, Lx, sbin)
        b = orc.binned_geometry(Ly, Lx, sbin)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert all(np.array_equal(x, y) for x, y in zip(a[2], b[2]))
        LY, LX, sy, sx = utils.video_placement(a[0], a[1])
        LY2, LX2, sy2, sx2 = orc.canvas_layout(a[0], a[1])
        assert (LY, LX) == (LY2, LX2) and np.array_equal(sy, sy2) and np.array_equal(sx, sx2)
        X = rng.rand(int((a[0] * a[1]).sum()), 3).astype(np.float32)
        assert np.array_equal(utils.multivideo_reshape(X, LY, LX, sy, sx, a[0], a[1], a[2]),
                              orc.scatter_canvas(X, LY, LX, sy, sx, a[0], a[1], a[2]))


def test_roi_bin_ranges_keep_reference_quirk_q5():
    r = motion_roi(0, 10, 70, 17, 90)
    yb, xb = utils.roi_bin_ranges(r, 4)
    yo, xo = orc.roi_bin_ranges(r, 4)
    assert np.array_equal(yb, yo) and np.array_equal(xb, xo)
    assert xb.size == 19 and yb.size == 15        # x stop 89/4 = 22.25 rounds up (Q5)


@pytest.mark.parametrize("N", [3, 99, 100, 700, 1000, 1001, 1999, 2000, 2300, 14999, 15000, 16001, 100000])
def test_stage_plans_match_oracle(N):
    from facemap_b200 import svd

    if N >= 100:
        nt0, tf = svd.stage1_plan(N, [N])
        nt0o, tfo = orc.stage1_sample_times(N, [N])
        assert nt0 == nt0o and np.array_equal(tf, tfo)
    a, b = svd.stage2_plan(N), orc.stage2_plan(N)
    assert a[:3] == b[:3] and np.array_equal(a[3], b[3])


def test_motion_layout_quirk_q1():
    from facemap_b200.process import _motion_from_energy

    for N in (1300, 300):
        e = np.arange(N, dtype=np.int64) * 16          # energy of frame t = 16 t  (row 0 unused)
        m = _motion_from_energy([e], 4, N)
        nt1 = min(500, N)
        assert np.array_equal(m[:nt1 - 1], np.arange(1, nt1, dtype=np.float32))
        assert m[nt1 - 1] == 0
        assert np.array_equal(m[nt1:], np.arange(nt1, N, dtype=np.float32))
    # two views accumulate in float32 like `M[0][t:] += imbin_mot.sum(-1)` (process.py:460)
    e0 = np.full(600, 2 ** 25 + 3, np.int64)
    e1 = np.full(600, 5, np.int64)
    m = _motion_from_energy([e0, e1], 1, 600)
    assert m[10] == np.float32(np.float32(2 ** 25 + 3) + np.float32(5))


def test_frame_sources_clamp_like_get_frames():
    import torch

    from facemap_b200.frames import HostArraySource, as_source

    v = np.arange(10 * 2 * 3, dtype=np.uint8).reshape(10, 2, 3)
    src = as_source([v])
    assert isinstance(src, HostArraySource)
    assert src.clamp(8, 508) == (8, 10) and src.clamp(-5, 3) == (0, 3) and src.clamp(0, 10) == (0, 10)
    assert src.nframes == 10 and src.Ly == [2] and src.Lx == [3] and src.block_frames == [10]
    assert src.h2d_bytes(4) == 24
    with pytest.raises(ValueError):
        HostArraySource([v, v[:5]])
    with pytest.raises(TypeError):
        as_source(42)
    with pytest.raises(ValueError):
        HostArraySource([torch.zeros(3, 4, 4)])


def test_gram_ksplit_and_batch_sizes():
    from facemap_b200 import svd

    assert svd._gram_ksplit(999, 19200) >= 8
    assert svd._gram_ksplit(3750, 19200) >= 1
    assert 1 <= svd._gram_ksplit(999, 300) <= 10
    r = svd.projection_rows_per_batch(19200, 500)
    assert r % 128 == 0 and r * 19200 * 4 <= 6 << 30
    assert svd.projection_rows_per_batch(1310720, 500) % 128 == 0


# ---------------------------------------------------------------------------------------------
# host side of the full-resolution ROI processors (facemap_b200/fullres.py) against oracle/roi_oracle.py
# ---------------------------------------------------------------------------------------------
def test_fullres_paint_plan_reproduces_in_place_painting():
    """R1: every ROI's crop mask, the running masks and the per-view mask that stage 3 bins equal what the
    reference's in-place writes (process.py:543, 567) leave behind, overlapping rectangles included"""
    from facemap_b200 import fullres
    from oracle import roi_oracle
    from tests.helpers import fullres_case

    for name in ("eye_all", "eye_reflector"):
        views, sbin, mot, mov, rois, full = fullres_case(name)
        views = [v[:3] for v in views]
        Ly, Lx = [v.shape[1] for v in views], [v.shape[2] for v in views]
        painters, runners, ormask = fullres.paint_plan(rois, Ly, Lx)
        ims = [v.copy() for v in views]
        pup, bl, run = roi_oracle.processing_order(rois)
        assert [p["index"] for p in painters] == pup + bl and [r["index"] for r in runners] == run
        for pl, i in zip(painters, pup + bl):
            r = rois[i]
            y0, y1, x0, x1 = roi_oracle.roi_rect(r)
            assert pl["rect"] == (y0, y1, x0, x1)
            crop = ims[r["ivid"]][:, y0:y1, x0:x1]
            crop[:, ~r["ellipse"]] = 255                                     # the reference's write through a view
            assert np.array_equal(np.where(pl["mask"] > 0, 255, views[r["ivid"]][:, y0:y1, x0:x1]), crop), (name, i)
        for rn in runners:
            y0, y1, x0, x1 = rn["rect"]
            assert np.array_equal(views[rn["view"]][:, y0:y1, x0:x1] | rn["ormask"], ims[rn["view"]][:, y0:y1, x0:x1])
        for v in range(len(views)):
            assert np.array_equal(views[v] | (ormask[v] if ormask[v] is not None else 0), ims[v])


def test_fullres_tables_match_oracle():
    from facemap_b200 import fullres
    from oracle import roi_oracle

    for ny, nx in ((21, 30), (24, 92), (85, 97)):                                   # running: g^2 in float32
        g = roi_oracle.running_filter(ny, nx).astype(np.complex64)
        assert np.array_equal(np.real(g * g), fullres.gaussian_fourier_sq(ny, nx))
    crop = np.arange(256).astype(np.uint8).reshape(1, 16, 16)                        # blink: R3 dtype rules
    for sat in (100.5, 100, 0, 255, 254.9):
        t = fullres.blink_terms(sat)
        assert t.dtype == np.float64 and abs(roi_oracle.blink_chunk(crop, sat)[0] - t.sum()) <= 1e-9 * max(1.0, t.sum())
    assert fullres.blink_terms(100)[200] == float((100 - 200) % 256)                 # uint8 wrap-around
    assert fullres.blink_terms(100.0)[200] == 0.0                                    # float clamp


def test_fullres_smooth_trace_matches_oracle_bitwise():
    import warnings

    from facemap_b200 import fullres
    from oracle import roi_oracle

    rs = np.random.RandomState(0)
    for trial in range(30):
        n = int(rs.choice([5, 29, 30, 31, 100, 700]))
        x = rs.randn(n).cumsum() + 10
        x[rs.rand(n) < 0.1] = np.nan
        if trial % 5 == 0:
            x[:] = np.nan
        if trial % 7 == 0:
            x[rs.randint(n)] = 1e3
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(fullres.smooth_trace(x), roi_oracle.smooth_trace(x), equal_nan=True), (trial, n)


def test_motion_needs_reference_rule():
    from facemap_b200 import svd

    assert not svd.motion_needs_reference(4, 19200) and not svd.motion_needs_reference(1, 24 * 32)
    assert svd.motion_needs_reference(3, 100) and svd.motion_needs_reference(12, 332) and svd.motion_needs_reference(7, 560)
    assert not svd.motion_needs_reference(1, 65793) and svd.motion_needs_reference(1, 65794)      # 255 * p >= 2^24
    assert svd.stage1_uses_reference(7) and not svd.stage1_uses_reference(8) and not svd.stage1_uses_reference(1)


def test_motion_ref_block_count_respects_scratch_budget():
    from facemap_b200 import svd

    assert svd.motion_ref_blocks(34080, 400, 3) == svd.MOTION_REF_BLOCKS                  # 640x480 at sbin 3: 548 KB per block
    assert svd.motion_ref_blocks(1280 * 1024, 13000, 1) == svd.MOTION_REF_BLOCKS          # sbin 1 keeps leaf sums only: 52 KB
    big = svd.motion_ref_blocks(1280 * 1024 // 9, 1500, 3)                                # a 1280x1024 view at sbin 3: 2.3 MB per block
    assert 8 <= big < svd.MOTION_REF_BLOCKS and big * (2 * (1280 * 1024 // 9) + 1500) * 8 <= svd.MOTION_REF_SCRATCH_MAX
    assert svd.motion_ref_blocks(10 ** 9, 10 ** 7, 7) == 8                                # never below 8 blocks


def test_single_chunk_masks_keep_the_references_zero_columns(monkeypatch):
    """svd.stage2_merge, nsegs == 1: the chunk's U S as is (quirk Q6), zero-padded to nsegs * nc components when the
    view has fewer binned pixels than that (facemap/process.py:146-151, 264-267); ROI targets are not padded"""
    import torch

    from facemap_b200 import kernels, svd

    class FakeK:
        alloc_matrix = staticmethod(kernels.alloc_matrix)

        @staticmethod
        def transpose(src, dst):
            dst.copy_(src.t())

    monkeypatch.setattr(svd, "K", FakeK)
    ws = object.__new__(svd.SvdWorkspace)
    ws.device = torch.device("cpu")
    g = torch.Generator().manual_seed(0)
    Y0, Y1 = torch.randn(384, 384, generator=g), torch.randn(60, 60, generator=g)
    out = svd.stage2_merge({"mot": [Y0, Y1]}, None, 1, None, ["mot"], True, ws, svd.StageTimer(enabled=False), nc=500)
    Uts, Us, sv = out["mot"]
    assert tuple(Uts[0].shape) == (500, 384) and tuple(Us[0].shape) == (384, 500)
    assert torch.equal(Uts[0][:384], Y0) and not Uts[0][384:].any()
    assert torch.equal(Us[0][:, :384], Y0.t()) and not Us[0][:, 384:].any()
    assert Uts[0].stride(1) == 1 and Uts[0].stride(0) % 4 == 0          # still a legal GEMM operand
    assert tuple(Uts[1].shape) == (60, 60) and tuple(Us[1].shape) == (60, 60) and torch.equal(Us[1], Y1.t())
    assert tuple(sv.shape) == (500,) and not sv.any()
    # enough pixels: nothing changes
    Y2 = torch.randn(500, 560, generator=g)
    Uts, Us, _ = svd.stage2_merge({"mot": [Y2]}, None, 1, None, ["mot"], True, ws, svd.StageTimer(enabled=False), nc=500)["mot"]
    assert Uts[0] is Y2 and tuple(Us[0].shape) == (560, 500)


def test_merge_loop_follows_the_references_len_u_mot_quirk(monkeypatch):
    """svd.stage2_merge with the motion SVD off: targets at positions >= number of ROIs stay unmerged (quirk Q8)"""
    import torch

    from facemap_b200 import kernels, svd

    merged_calls = []

    class FakeK:
        alloc_matrix = staticmethod(kernels.alloc_matrix)

        @staticmethod
        def transpose(src, dst):
            dst.copy_(src.t())

    def fake_merge(ws, Y, k, timer, label):
        merged_calls.append((label, tuple(Y.shape), k))
        return Y[:k], Y[:k].t().contiguous(), torch.ones(k)

    monkeypatch.setattr(svd, "K", FakeK)
    monkeypatch.setattr(svd, "merge_components", fake_merge)
    ws = object.__new__(svd.SvdWorkspace)
    ws.device = torch.device("cpu")
    t = svd.StageTimer(enabled=False)
    full, r0, r1 = torch.randn(500, 396), torch.randn(80, 40), torch.randn(180, 90)
    # movie only, two ROIs: full frame and first ROI merged, last ROI kept as stacked chunk blocks
    Uts, Us, sv = svd.stage2_merge({"mov": [full, r0, r1]}, None, 2, None, ["mov"], True, ws, t, nc=250, mot_on=False)["mov"]
    assert [c[0] for c in merged_calls] == ["stage2.merge", "stage2.roi_merge"]
    assert tuple(Us[0].shape) == (396, 395) and tuple(Us[1].shape) == (40, 39) and tuple(Us[2].shape) == (90, 180)
    assert torch.equal(Us[2], r1.t()) and sv.numel() == 39
    # movie only, no ROI: nothing merged, singular values stay zero
    merged_calls.clear()
    Uts, Us, sv = svd.stage2_merge({"mov": [full]}, None, 2, None, ["mov"], True, ws, t, nc=250, mot_on=False)["mov"]
    assert not merged_calls and tuple(Us[0].shape) == (396, 500) and not sv.any()
    # motion SVD on: everything merged, for both modalities
    merged_calls.clear()
    svd.stage2_merge({"mot": [full, r0], "mov": [full, r0]}, None, 2, None, ["mot", "mov"], True, ws, t, nc=250, mot_on=True)
    assert len(merged_calls) == 4


def test_geometry_refuses_rois_outside_their_view():
    """a rectangle outside its binned view would make fm_gather_roi read out of bounds: refused by name"""
    import pytest

    from facemap_b200 import svd

    def roi(ivid, yb, xb):
        return {"rind": 1, "ivid": ivid, "yrange_bin": np.asarray(yb), "xrange_bin": np.asarray(xb)}

    svd.Geometry([48], [64], 4, [roi(0, range(2, 12), range(3, 16))])                # touches the border: fine
    for bad in (roi(1, range(2, 5), range(3, 6)), roi(0, range(2, 13), range(3, 6)), roi(0, range(2, 5), range(3, 17)),
                roi(0, range(-1, 5), range(3, 6))):
        with pytest.raises(svd.FacemapB200Error, match="motion ROI 0"):
            svd.Geometry([48], [64], 4, [bad])


def test_save_npy_writes_what_np_save_writes(tmp_path):
    """process.save_npy (protocol-5 buffers, parallel positioned writes) against np.save on a proc-like dict: np.load
    returns the same object either way (facemap/process.py:612 writes the file with np.save)"""
    from facemap_b200.process import save_npy

    rs = np.random.RandomState(0)
    big = rs.rand(3000, 500).astype(np.float32)                       # 6 MB: goes through the threaded path
    proc = {"filenames": [["a.avi", "b.avi"]], "Ly": [480, 96], "sbin": 4, "rois": None, "save_mat": False,
            "motSVD": [big, np.zeros((0, 1), np.float32)], "motMask_reshape": [big.reshape(60, 50, 500)],
            "motSv": rs.rand(500).astype(np.float32), "avgframe": [rs.rand(7).astype(np.float32)],
            "pupil": [{"area": rs.rand(10), "com": rs.rand(10, 2)}], "sybin": np.array([0, 3]), "empty": []}
    a, b = str(tmp_path / "a.npy"), str(tmp_path / "b.npy")
    np.save(a, proc)
    save_npy(b, proc, threads=3, piece=1 << 20)
    x, y = np.load(a, allow_pickle=True).item(), np.load(b, allow_pickle=True).item()
    assert list(x) == list(y)

    def same(u, v):
        if isinstance(u, dict):
            return list(u) == list(v) and all(same(u[k], v[k]) for k in u)
        if isinstance(u, list):
            return len(u) == len(v) and all(same(p, q) for p, q in zip(u, v))
        if isinstance(u, np.ndarray):
            return u.dtype == v.dtype and u.shape == v.shape and np.array_equal(u, v)
        return u == v

    assert same(x, y)
    assert abs(os.path.getsize(a) - os.path.getsize(b)) < 4096


def test_exact_groups_fold_equals_numpy_float32_sum():
    """kernels.exact_groups: for byte-valued data the maximal subtrees of NumPy's pairwise tree with <= 65 793 elements
    never round in float32, so summing them as integers and folding only the merges above them reproduces
    `a.astype(float32).sum()` bit for bit (what fm_motion_ref_planes does on the device)"""
    import ctypes as C

    from facemap_b200 import _lib
    from facemap_b200 import kernels as K

    def plan(n):
        cap = max(1, n // 64 + 2)
        lo, ln, comb = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.uint8)
        c = int(_lib.load().fm_pairwise_plan(int(n), cap, C.c_void_p(lo.ctypes.data), C.c_void_p(ln.ctypes.data),
                                             C.c_void_p(comb.ctypes.data)))
        return lo[:c], ln[:c], comb[:c]

    rs = np.random.RandomState(0)
    for n in (35, 4096, 65793, 65794, 66013, 93000, 198039, 1024 * 1280, 1000003):
        lo, ln, comb = plan(n)
        gf, gn, gc = K.exact_groups(ln, comb)
        assert max(int(ln[f:f + c].sum()) for f, c in zip(gf, gn)) <= K.EXACT_BYTE_SUM_PIXELS
        assert (len(gf) == 1) == (n <= K.EXACT_BYTE_SUM_PIXELS)
        for lo_v, hi_v in ((200, 256), (0, 256), (255, 256)):
            a = rs.randint(lo_v, hi_v, size=n).astype(np.uint8)
            want = a.astype(np.float32).sum(dtype=np.float32)
            csum = np.concatenate(([0], np.cumsum(a.astype(np.int64))))
            leaf = csum[lo + ln] - csum[lo]
            lsum = np.concatenate(([0], np.cumsum(leaf)))
            stack = []
            for f, c, cmb in zip(gf, gn, gc):
                v = np.float32(lsum[f + c] - lsum[f])
                for _ in range(int(cmb)):
                    v = np.float32(stack.pop() + v)
                stack.append(v)
            assert len(stack) == 1 and stack[0] == want, (n, lo_v)


def test_pinned_pool_reuses_a_buffer_only_after_its_last_user_is_gone():
    """process._PinnedPool with a plain allocator: whatever still refers to a handed-out buffer — the tensor, a slice
    of it, a NumPy array made from it, a view of that array — keeps it out of the pool."""
    import gc

    import torch

    from facemap_b200.process import _PinnedPool

    sizes = []
    pool = _PinnedPool(allocate=lambda n: (sizes.append(n), torch.empty(n, dtype=torch.uint8))[1])
    a = pool.take((1000, 500))
    arr, ptr_a = a.numpy(), a.data_ptr()
    a[3:5] = 7.0
    assert arr[3, 0] == 7.0 and arr.flags.writeable
    del a
    gc.collect()
    b = pool.take((1000, 500))
    assert b.data_ptr() != ptr_a                          # the array is still held
    view = arr[10:20]
    del arr
    gc.collect()
    c = pool.take((1000, 500))
    assert c.data_ptr() != ptr_a                          # ... and so is a view of it
    piece, ptr_c = c[5:9], c.data_ptr()
    del c
    gc.collect()
    assert pool.take((1000, 500)).data_ptr() != ptr_c     # a tensor slice holds its buffer too
    del view
    gc.collect()
    d = pool.take((999, 500))                             # a slightly smaller request shares the 2 MB size class
    assert d.data_ptr() == ptr_a
    assert sizes == [2 << 20] * 4 and pool.stats["reused"] == 1
    assert tuple(pool.take((0, 1)).shape) == (0, 1) and len(sizes) == 4          # empty results take no buffer
    small = pool.take((3,), torch.int32)
    assert small.dtype == torch.int32 and sizes[-1] == 4096
    del piece
    gc.collect()
    assert pool.take((1000, 500)).data_ptr() == ptr_c
    # free buffers beyond keep_bytes are dropped instead of kept
    tiny = _PinnedPool(allocate=lambda n: torch.empty(n, dtype=torch.uint8), keep_bytes=4096)
    x, y = tiny.take((1024,)), tiny.take((1024,))
    del x, y
    gc.collect()
    assert tiny.free_bytes == 4096


def test_merges_are_dealt_by_a_cost_linear_in_the_problem_order():
    """svd.deal_merges: configs[2]'s eight merge problems (2 modalities x (full frame 3750 + ROIs 2400, 1888, 2500)).
    cuSOLVER's eigensolve costs ~19 us per column, so the full-frame problems must not each fill a rank on their own."""
    from facemap_b200.svd import deal_merges

    sizes = {(m, i): n for m in ("mot", "mov") for i, n in enumerate((3750, 2400, 1888, 2500))}
    assert deal_merges(sizes, 1) == {} and deal_merges({("mot", 0): 3750}, 8) == {}
    for world, worst in ((2, 10538), (4, 5638), (8, 3750)):
        own = deal_merges(sizes, world)
        assert set(own) == set(sizes) and set(own.values()) <= set(range(world))
        load = [sum(n for t, n in sizes.items() if own[t] == r) for r in range(world)]
        assert max(load) == worst, (world, load)
    # deterministic: every rank computes the same table
    assert deal_merges(sizes, 4) == deal_merges(dict(sizes), 4)


class _FiredEvent:
    def synchronize(self):
        pass


def test_early_file_is_the_final_pickle_or_nothing(tmp_path):
    """process._EarlyFile on CPU tensors: rows written while 'the pass runs' land where the final pickle puts them
    (mask and its canvas view: twice), rows changed on the host afterwards are written again, and a final dict whose
    layout is not the announced one leaves no file behind."""
    import torch

    from facemap_b200.process import _EarlyFile, save_npy

    rng = np.random.default_rng(5)
    N, p, k = 6000, 1200, 500

    def build_full(out):
        return {"filenames": [["a.avi"]], "sbin": 2, "motion": out["motion"], "motSv": out["motSv"],
                "motMask": out["motMask"], "motMask_reshape": out["motMask_reshape"], "motSVD": out["motSVD"], "rois": None}

    def scenario(path, final_motion_len=N, second_begin=False, copy_v=False):
        V = torch.zeros((N, k), dtype=torch.float32)
        U = torch.zeros((p, k), dtype=torch.float32)
        S = torch.zeros(k, dtype=torch.float32)
        early = _EarlyFile(path, build_full)
        U.copy_(torch.from_numpy(rng.standard_normal((p, k)).astype(np.float32)))
        early.landed(U, _FiredEvent())                       # the masks land before the layout is announced
        like = {"motion": [np.empty(N, np.float32)], "motSv": S.numpy(), "motMask": [U.numpy()],
                "motMask_reshape": [U.numpy().reshape(30, 40, k)], "motSVD": [V.numpy()]}
        early.begin(like)
        if second_begin:
            early.begin(like)
        for a in range(0, N, 1000):
            V[a:a + 1000] = torch.from_numpy(rng.standard_normal((1000, k)).astype(np.float32))
            if a != 3000:                                    # one batch is never announced: written at the end
                early.landed(V[a:a + 1000], _FiredEvent())
        S.copy_(torch.arange(k, dtype=torch.float32))
        v = V.numpy()
        v[0] = v[1]
        early.dirty(v[0:1])
        out = {"motion": [rng.standard_normal(final_motion_len).astype(np.float32)], "motSv": S.numpy(),
               "motMask": [U.numpy()], "motMask_reshape": [U.numpy().reshape(30, 40, k)],
               "motSVD": [v.copy() if copy_v else v]}
        full = build_full(out)
        return early, full

    path = str(tmp_path / "a_proc.npy")
    early, full = scenario(path)
    assert early.finish(full) is True
    assert sum(b - a for a, b in early.done) == 4 * (5 * 1000 * k + 2 * p * k)       # five batches, the mask twice
    plain = str(tmp_path / "plain.npy")
    save_npy(plain, full)
    assert open(path, "rb").read() == open(plain, "rb").read()
    assert sorted(os.listdir(tmp_path)) == ["a_proc.npy", "plain.npy"]
    back = np.load(path, allow_pickle=True).item()
    assert np.array_equal(back["motSVD"][0], full["motSVD"][0]) and np.array_equal(back["motSVD"][0][0], back["motSVD"][0][1])
    assert np.array_equal(back["motMask_reshape"][0], full["motMask"][0].reshape(30, 40, k))
    # not the announced layout, or not the announced buffers -> nothing is left, the caller writes the file the plain way
    for kw in ({"final_motion_len": N + 1}, {"second_begin": True}, {"copy_v": True}):
        path2 = str(tmp_path / "b_proc.npy")
        early, full = scenario(path2, **kw)
        assert early.finish(full) is False
        assert sorted(os.listdir(tmp_path)) == ["a_proc.npy", "plain.npy"]
    # a pass that raises
    early, _ = scenario(str(tmp_path / "c_proc.npy"))
    early.abandon()
    assert sorted(os.listdir(tmp_path)) == ["a_proc.npy", "plain.npy"]


def test_subspace_topk_matches_lapack_or_declines():
    """svd.subspace_topk (torch, runs on the CPU here): on a decaying spectrum the top-k eigenpairs agree with LAPACK
    far inside the path's gates and come back in the layout fm_topk_components reads (ascending rows); on a flat
    spectrum, where the block cannot converge in the allowed rounds, it declines (None: the caller solves directly)."""
    import torch

    from facemap_b200 import svd

    rng = np.random.default_rng(11)
    n, k = 1200, 200
    Qr, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = (np.arange(1, n + 1, dtype=np.float64) ** -1.6) * (1 + 0.2 * rng.random(n))        # S ~ k^-0.8: a video's tail
    lam = np.sort(lam)[::-1]
    G = torch.from_numpy((Qr * lam) @ Qr.T)
    G = (G + G.T) / 2
    calls = []

    def rr(Hm, kk):
        calls.append(Hm.shape[0])
        return svd._rr_lapack(Hm, kk)

    th, W = svd.subspace_topk(G, k, rr=rr)
    assert calls == [k + svd.SUBSPACE_EXTRA]                         # accepted at the first Rayleigh-Ritz step
    assert th.shape == (k,) and W.shape == (k, n) and W.is_contiguous() and bool((th[1:] >= th[:-1]).all())
    assert np.max(np.abs(th.numpy()[::-1] - lam[:k]) / lam[:k]) < 1e-6
    lead = np.abs(np.sum(W.numpy()[::-1][:20].T * Qr[:, :20], axis=0))       # well-separated leading vectors
    assert np.sqrt(np.max(1 - np.minimum(lead, 1.0) ** 2)) < 1e-5
    assert float((W @ W.T - torch.eye(k, dtype=torch.float64)).abs().max()) < 1e-10
    th2, W2 = svd.subspace_topk(G, k, rr=rr)                         # deterministic
    assert torch.equal(th, th2) and torch.equal(W, W2)
    # flat tail: lambda_{b+1} / lambda_k = 0.98 -> no convergence in 8 + 4 + 4 multiplications
    lam_flat = np.r_[np.linspace(1.0, 0.5, 50), np.linspace(0.2, 0.19, n - 50)]
    Gf = torch.from_numpy((Qr * lam_flat) @ Qr.T)
    assert svd.subspace_topk((Gf + Gf.T) / 2, k) is None
    assert svd.subspace_eligible(3750, 500) and not svd.subspace_eligible(2500, 500) and not svd.subspace_eligible(3750, 999)


def test_merge_cost_knows_which_solver_a_merge_gets():
    from facemap_b200.svd import merge_cost

    assert merge_cost((3750, 38400), 500, "cusolver") == 3750          # direct solve: the order of the problem
    assert merge_cost((3750, 38400), 500, "subspace") == 1500          # subspace iteration: 29 ms ~ 1 500 cuSOLVER columns
    assert merge_cost((3750, 2400), 500, "subspace") == 2400           # wide ROI merge: row-side problem, always direct
    assert merge_cost((1000, 38400), 500, "subspace") == 1000          # too small for the subspace solver
