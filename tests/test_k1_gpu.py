"""K1 (fused bin + diff + |.| + avg subtract + energy + running sums) against the oracle.
Integer outputs must be bit-exact; float outputs bit-exact for power-of-two sbin and within
1 ulp otherwise (north_star tolerance)."""
import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc
from tests.helpers import ulp_distance

pytestmark = pytest.mark.gpu


def _video(T, H, W, seed):
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, size=(1, H, W))
    drift = rng.randint(-40, 41, size=(T, H, W))
    return np.clip(base + np.cumsum(drift, axis=0) // 4 + rng.randint(0, 30, size=(T, H, W)), 0, 255).astype(np.uint8)


def _oracle(frames, sbin, avgmot, avgfrm):
    T, H, W = frames.shape
    Lyb, Lxb = H // sbin, W // sbin
    S = orc.bin_sums_int(frames, sbin, Lyb, Lxb)                       # int64 [T, p]
    b = orc.bin_frames(frames, sbin, Lyb, Lxb)                          # reference float binning
    d = np.abs(np.diff(b, axis=0))
    xm = (d - avgmot).astype(np.float32)
    xf = (b[1:] - avgfrm).astype(np.float32)
    dS = np.abs(np.diff(S, axis=0))
    return S, xm, xf, dS


CASES = [
    # T, H, W, sbin   (vector path needs W % min(16, 4*sbin) == 0)
    (40, 64, 96, 4), (37, 60, 100, 4), (33, 48, 64, 2), (20, 24, 32, 1), (25, 64, 128, 8),
    (19, 64, 64, 16), (21, 50, 70, 3), (18, 61, 67, 4), (16, 30, 50, 5), (140, 32, 48, 4),
    # non-power-of-two bins on 4-byte aligned rows: vector kernel instantiations for sbin 3, 5, 6, 7, 10, 12
    (20, 60, 96, 3), (14, 70, 140, 7), (10, 96, 144, 12), (15, 60, 100, 5), (9, 66, 132, 6),
    # ... with a cropped remainder / a partial last pixel group (the row-end guard of the vector kernel)
    (6, 42, 640, 3), (6, 49, 648, 7), (5, 60, 104, 10), (7, 36, 76, 6),
]


@pytest.mark.parametrize("T,H,W,sbin", CASES)
def test_k1_matches_oracle(cuda_device, T, H, W, sbin):
    from facemap_b200 import kernels as K

    frames = _video(T, H, W, seed=T + H + sbin)
    Lyb, Lxb = H // sbin, W // sbin
    p = Lyb * Lxb
    rng = np.random.RandomState(1)
    avgmot = (rng.rand(p) * 3).astype(np.float32)
    avgfrm = (rng.rand(p) * 200).astype(np.float32)
    S, xm, xf, dS = _oracle(frames, sbin, avgmot, avgfrm)

    dev = cuda_device
    fr = torch.from_numpy(frames).to(dev)
    pad = 5
    Xm = K.alloc_matrix(T - 1, p + pad, dev, zero=True)
    Xf = K.alloc_matrix(T - 1, p + pad, dev, zero=True)
    energy = torch.zeros(T - 1, dtype=torch.int64, device=dev)
    last = torch.zeros(p, dtype=torch.int32, device=dev)
    tsb = torch.zeros(p, dtype=torch.int64, device=dev)
    tsa = torch.zeros(p, dtype=torch.int64, device=dev)
    rows = K.bin_motion(fr, sbin, avgmotion=torch.from_numpy(avgmot).to(dev), avgframe=torch.from_numpy(avgfrm).to(dev),
                        x_mot=Xm, x_mov=Xf, col0=pad, energy=energy, last_sum=last, tsum_bin=tsb, tsum_adiff=tsa)
    torch.cuda.synchronize()
    assert rows == T - 1
    assert np.array_equal(energy.cpu().numpy(), dS.sum(axis=1))
    assert np.array_equal(last.cpu().numpy(), S[-1])
    assert np.array_equal(tsb.cpu().numpy(), S.sum(axis=0))
    assert np.array_equal(tsa.cpu().numpy(), dS.sum(axis=0))
    gm, gf = Xm.cpu().numpy()[:, pad:], Xf.cpu().numpy()[:, pad:]
    assert np.all(Xm.cpu().numpy()[:, :pad] == 0), "columns before col0 must be untouched"
    pow2 = sbin & (sbin - 1) == 0
    if pow2:
        assert np.array_equal(gm, xm) and np.array_equal(gf, xf)
    else:
        assert ulp_distance(gm, xm) <= 1 and ulp_distance(gf, xf) <= 1


@pytest.mark.parametrize("T,H,W,sbin", [(50, 64, 96, 4), (30, 50, 70, 3), (45, 32, 64, 1)])
def test_k1_carry_is_seamless(cuda_device, T, H, W, sbin):
    """Processing [0,a) then [a,T) with the carried block sums equals one pass (process.py:449-453)."""
    from facemap_b200 import kernels as K

    frames = _video(T, H, W, seed=7)
    Lyb, Lxb = H // sbin, W // sbin
    p = Lyb * Lxb
    dev = cuda_device
    fr = torch.from_numpy(frames).to(dev)
    whole = K.alloc_matrix(T - 1, p, dev)
    e_whole = torch.zeros(T - 1, dtype=torch.int64, device=dev)
    K.bin_motion(fr, sbin, x_mot=whole, energy=e_whole)
    a = 17
    part = K.alloc_matrix(T - 1, p, dev)
    e_part = torch.zeros(T - 1, dtype=torch.int64, device=dev)
    carry = torch.zeros(p, dtype=torch.int32, device=dev)
    K.bin_motion(fr[:a].contiguous(), sbin, x_mot=part[:a - 1], energy=e_part[:a - 1], last_sum=carry)
    rows = K.bin_motion(fr[a:].contiguous(), sbin, prev_sum=carry, x_mot=part[a - 1:], energy=e_part[a - 1:])
    torch.cuda.synchronize()
    assert rows == T - a
    assert torch.equal(whole, part) and torch.equal(e_whole, e_part)


def test_k1_roi_energy(cuda_device):
    from facemap_b200 import kernels as K

    T, H, W, sbin = 30, 64, 96, 4
    frames = _video(T, H, W, seed=3)
    Lyb, Lxb = H // sbin, W // sbin
    S = orc.bin_sums_int(frames, sbin, Lyb, Lxb).reshape(T, Lyb, Lxb)
    dS = np.abs(np.diff(S, axis=0))
    rects = [(2, 9, 3, 20), (0, 16, 0, 24), (5, 6, 7, 8)]
    dev = cuda_device
    fr = torch.from_numpy(frames).to(dev)
    energy = torch.zeros(T - 1, dtype=torch.int64, device=dev)
    roi_e = torch.zeros((len(rects), T - 1), dtype=torch.int64, device=dev)
    K.bin_motion(fr, sbin, energy=energy, rois=rects, roi_energy=roi_e)
    torch.cuda.synchronize()
    got = roi_e.cpu().numpy()
    for q, (y0, y1, x0, x1) in enumerate(rects):
        assert np.array_equal(got[q], dS[:, y0:y1, x0:x1].sum(axis=(1, 2)))
    assert np.array_equal(energy.cpu().numpy(), dS.sum(axis=(1, 2)))


def test_stage1_mean_update_bit_exact(cuda_device):
    """fm_mean_update reproduces subsampled_mean's float32 += float64 accumulation (process.py:82-96)."""
    from facemap_b200 import kernels as K

    for sbin in (1, 2, 4):
        H, W, T = 32 * sbin // (2 if sbin > 1 else 1), 48, 100
        segs = [_video(T, H, W, seed=s) for s in range(3)]
        Lyb, Lxb = H // sbin, W // sbin
        p = Lyb * Lxb
        ref_f = np.zeros(p, np.float32)
        ref_m = np.zeros(p, np.float32)
        dev = cuda_device
        acc_f = torch.zeros(p, dtype=torch.float32, device=dev)
        acc_m = torch.zeros(p, dtype=torch.float32, device=dev)
        for fr in segs:
            b = orc.bin_frames(fr, sbin, Lyb, Lxb)
            ref_f += b.mean(axis=0)
            ref_m += np.abs(np.diff(b, axis=0)).mean(axis=0)
            tsb = torch.zeros(p, dtype=torch.int64, device=dev)
            tsa = torch.zeros(p, dtype=torch.int64, device=dev)
            K.bin_motion(torch.from_numpy(fr).to(dev), sbin, tsum_bin=tsb, tsum_adiff=tsa)
            K.mean_update(acc_f, tsb, sbin, T)
            K.mean_update(acc_m, tsa, sbin, T - 1)
        ref_f /= float(len(segs))
        ref_m /= float(len(segs))
        K.mean_update(acc_f, None, sbin, 1, finalize_div=len(segs))
        K.mean_update(acc_m, None, sbin, 1, finalize_div=len(segs))
        torch.cuda.synchronize()
        assert np.array_equal(acc_f.cpu().numpy(), ref_f), f"avgframe sbin={sbin}"
        assert np.array_equal(acc_m.cpu().numpy(), ref_m), f"avgmotion sbin={sbin}"


@pytest.mark.parametrize("T,H,W,sbin", CASES)
@pytest.mark.parametrize("with_carry,lean", [(False, False), (True, False), (True, True)])
def test_k1_painted_equals_k1_on_painted_frames(cuda_device, T, H, W, sbin, with_carry, lean):
    """fm_bin_motion_painted(frames, ormask) == fm_bin_motion(frames | ormask) bit for bit (every output),
    i.e. the fused form of process.py:543, 567 painting the chunk buffer before it is binned."""
    from facemap_b200 import kernels as K
    from facemap_b200._lib import Rois

    dev = cuda_device
    frames = _video(T, H, W, seed=3 * T + sbin)
    rng = np.random.RandomState(7)
    m = np.zeros((H, W), np.uint8)
    m[H // 5:H // 2, W // 6:2 * W // 3] = np.where(rng.rand(H // 2 - H // 5, 2 * W // 3 - W // 6) < 0.6, 255, 0)
    Lyb, Lxb = H // sbin, W // sbin
    p = Lyb * Lxb
    avgmot = torch.from_numpy((rng.rand(p) * 3).astype(np.float32)).to(dev)
    avgfrm = torch.from_numpy((rng.rand(p) * 200).astype(np.float32)).to(dev)
    rects = [(1, max(2, Lyb // 2), 0, max(1, Lxb // 2))]
    prev = torch.from_numpy(rng.randint(0, 255 * sbin * sbin, size=p).astype(np.int32)).to(dev) if with_carry else None
    rows = T if with_carry else T - 1

    def run(fr, ormask):
        out = {"xm": K.alloc_matrix(rows, p, dev, zero=True), "xf": K.alloc_matrix(rows, p, dev, zero=True),
               "e": torch.zeros(rows, dtype=torch.int64, device=dev),
               "re": torch.zeros((1, rows), dtype=torch.int64, device=dev),
               "last": torch.zeros(p, dtype=torch.int32, device=dev),
               "tsb": torch.zeros(p, dtype=torch.int64, device=dev), "tsa": torch.zeros(p, dtype=torch.int64, device=dev)}
        if lean:      # the motion-only instantiation (no movie operand, no ROI energies)
            K.bin_motion(fr, sbin, prev_sum=prev, avgmotion=avgmot, x_mot=out["xm"], energy=out["e"],
                         last_sum=out["last"], ormask=ormask)
        else:
            K.bin_motion(fr, sbin, prev_sum=prev, avgmotion=avgmot, avgframe=avgfrm, x_mot=out["xm"], x_mov=out["xf"],
                         energy=out["e"], rois=Rois.from_rects(rects), roi_energy=out["re"], last_sum=out["last"],
                         tsum_bin=out["tsb"], tsum_adiff=out["tsa"], ormask=ormask)
        return out

    a = run(torch.from_numpy(frames).to(dev), torch.from_numpy(m).to(dev))
    b = run(torch.from_numpy(frames | m).to(dev), None)
    for key in a:
        assert torch.equal(a[key], b[key]), key


@pytest.mark.parametrize("T,H,W,sbin", [(9, 61, 67, 3), (7, 144, 192, 12), (12, 140, 196, 7), (6, 50, 70, 5),
                                        (5, 300, 310, 1), (4, 24, 32, 1), (5, 96, 128, 4), (3, 400, 600, 1),
                                        (4, 251, 263, 1), (3, 5, 7, 1)])
def test_motion_ref_matches_numpy_bitwise(cuda_device, T, H, W, sbin):
    """fm_motion_ref == the reference's `np.abs(np.diff(spatial_bin(im))).sum(-1)` (process.py:34-43, 455-460),
    bit for bit, with and without a carried previous frame and a paint mask"""
    from facemap_b200 import kernels as K

    dev = cuda_device
    frames = _video(T + 1, H, W, seed=11 * T + sbin)
    if H * W > 60000:         # white noise: float32 sums beyond 2^24, where the summation order shows
        frames = np.random.RandomState(9).randint(0, 256, size=(T + 1, H, W)).astype(np.uint8)
    m = np.where(np.random.RandomState(3).rand(H, W) < 0.2, 255, 0).astype(np.uint8)
    Lyb, Lxb = H // sbin, W // sbin
    plan = K.pairwise_plan(Lyb * Lxb, dev)
    blocks = 3
    scratch = K.motion_ref_scratch(H, W, sbin, plan[0].numel(), blocks, dev)
    for mask in (None, m):
        painted = frames if mask is None else (frames | mask)
        b = orc.bin_frames(painted, sbin, Lyb, Lxb)
        ref = np.abs(np.diff(b, axis=0)).sum(axis=-1)            # float64 (sbin > 1) or float32 (sbin == 1)
        d_mask = None if mask is None else torch.from_numpy(mask).to(dev)
        out = torch.zeros(T, dtype=torch.float64, device=dev)
        rows = K.motion_ref(torch.from_numpy(frames).to(dev), sbin, plan, scratch, blocks, out, ormask=d_mask)
        assert rows == T
        assert np.array_equal(out.cpu().numpy(), ref.astype(np.float64))
        out2 = torch.zeros(T, dtype=torch.float64, device=dev)
        rows = K.motion_ref(torch.from_numpy(frames[1:]).to(dev), sbin, plan, scratch, blocks, out2,
                            prev_frame=torch.from_numpy(frames[0]).to(dev), ormask=d_mask)
        assert rows == T and np.array_equal(out2.cpu().numpy(), ref.astype(np.float64))


@pytest.mark.parametrize("T,H,W,col0,masked", [(6, 300, 310, 0, False), (5, 251, 263, 24, True), (4, 1024, 1280, 0, False),
                                               (3, 40, 56, 7, True)])
def test_motion_ref_planes_is_k1_plus_motion_ref(cuda_device, T, H, W, col0, masked):
    """fm_motion_ref_planes (sbin 1: one pass for the reference-arithmetic sums, the operand plane and the integer
    energy) against the two kernels it replaces in stage 3, fm_bin_motion_planes and fm_motion_ref, bit for bit — with
    and without a carried previous frame, a paint mask, an unaligned column offset"""
    from facemap_b200 import kernels as K

    dev = cuda_device
    frames = np.random.RandomState(T * H).randint(0, 256, size=(T + 1, H, W)).astype(np.uint8)
    mask = torch.from_numpy(np.where(np.random.RandomState(3).rand(H, W) < 0.2, 255, 0).astype(np.uint8)).to(dev) if masked else None
    npix = H * W
    ldq = K.round_up(col0 + npix, 16)
    plan = K.pairwise_plan(npix, dev)
    gplan = K.pairwise_plan(npix, dev, groups=True)          # + the exact-subtree fold plan (kernels.exact_groups)
    assert len(gplan) == 6 and int(gplan[4].sum()) == plan[0].numel()
    blocks = 5
    scratch = K.motion_ref_scratch(H, W, 1, plan[0].numel(), blocks, dev)
    fr = torch.from_numpy(frames).to(dev)
    for carried in (False, True):
        src = fr[1:] if carried else fr
        prev = fr[0].contiguous() if carried else None
        rows = T
        prev_sum = None
        if carried:
            prev_sum = torch.zeros(npix, dtype=torch.int32, device=dev)
            K.bin_motion(fr[:1].contiguous(), 1, last_sum=prev_sum, ormask=mask)
        # the two-kernel route
        want_plane = torch.full((rows, ldq), 7, dtype=torch.uint8, device=dev)
        want_e = torch.zeros(rows, dtype=torch.int64, device=dev)
        K.bin_motion_planes(src.contiguous(), 1, prev_sum=prev_sum, q_mot=(None, want_plane[:, :col0 + npix]), col0=col0,
                            energy=want_e, ormask=mask)
        want_ref = torch.zeros(rows, dtype=torch.float64, device=dev)
        K.motion_ref(src.contiguous(), 1, plan, scratch, blocks, want_ref, prev_frame=prev, ormask=mask)
        # the fused kernel
        got_plane = torch.full((rows, ldq), 7, dtype=torch.uint8, device=dev)
        got_e = torch.zeros(rows, dtype=torch.int64, device=dev)
        got_ref = torch.zeros(rows, dtype=torch.float64, device=dev)
        assert K.motion_ref_planes(src.contiguous(), plan, scratch, blocks, got_ref, got_plane[:, :col0 + npix], col0=col0,
                                   energy=got_e, prev_frame=prev, ormask=mask) == rows
        torch.cuda.synchronize()
        assert torch.equal(got_plane, want_plane) and torch.equal(got_e, want_e) and torch.equal(got_ref, want_ref)
        # the same with the leaves folded by exact groups instead of one by one
        got_ref2 = torch.zeros(rows, dtype=torch.float64, device=dev)
        got_plane.fill_(9)
        got_e.zero_()
        K.motion_ref_planes(src.contiguous(), gplan, scratch, blocks, got_ref2, got_plane[:, :col0 + npix], col0=col0,
                            energy=got_e, prev_frame=prev, ormask=mask)
        torch.cuda.synchronize()
        assert torch.equal(got_ref2, want_ref) and torch.equal(got_e, want_e)
        assert torch.equal(got_plane[:, col0:col0 + npix], want_plane[:, col0:col0 + npix])


def test_motion_ref_full_resolution_view_with_budgeted_blocks(cuda_device):
    """a 1280x1024 view at sbin 1 (BASELINE configs[3]) with the block count svd.motion_ref_blocks picks under its
    1 GiB scratch budget: still the reference's float32 sums, bit for bit"""
    from facemap_b200 import kernels as K, svd

    dev = cuda_device
    T, H, W = 4, 1024, 1280
    frames = np.random.RandomState(5).randint(0, 256, size=(T + 1, H, W)).astype(np.uint8)
    plan = K.pairwise_plan(H * W, dev)
    blocks = svd.motion_ref_blocks(H * W, plan[0].numel(), 1)
    assert 8 <= blocks <= svd.MOTION_REF_BLOCKS
    scratch = K.motion_ref_scratch(H, W, 1, plan[0].numel(), blocks, dev)
    assert scratch.numel() <= svd.MOTION_REF_SCRATCH_MAX
    ref = np.abs(np.diff(orc.bin_frames(frames, 1, H, W), axis=0)).sum(axis=-1)
    out = torch.zeros(T, dtype=torch.float64, device=dev)
    assert K.motion_ref(torch.from_numpy(frames).to(dev), 1, plan, scratch, blocks, out) == T
    assert np.array_equal(out.cpu().numpy(), ref.astype(np.float64))
