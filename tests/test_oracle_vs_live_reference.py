"""The oracle against the UNMODIFIED reference run live (oracle/ref_harness.py) on configurations the committed golden
fixtures do not hold: odd frame sizes, sequential blocks, a non-power-of-two bin, frame counts around the chunk
boundaries.  Integer stages bit for bit with the stock reference; the whole proc dict (singular values, masks,
projections) bit for bit with the reference whose `utils.svdecon` is the exact float64 one — the same two oracles as
SURVEY §8(c).  Runs where /root/reference exists (the build container); each case takes a few seconds."""
import numpy as np
import pytest

from facemap_b200.synth import SynthVideo
from oracle import facemap_oracle as orc
from oracle import ref_harness

from .helpers import motion_roi

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference (build container)")

# (views [(H, W, seed)], frames per sequential block, sbin, motSVD, movSVD[, motion ROIs, fullSVD])
CASES = {
    "odd_size_s2": ([(46, 70, 21)], [1203], 2, True, True),
    "two_blocks_s3": ([(45, 63, 22)], [700, 650], 3, True, False),
    "two_views_s5": ([(50, 60, 23), (40, 75, 24)], [1100], 5, True, True),
    "below_one_chunk_s1": ([(20, 28, 25)], [640], 1, True, False),
    "tiny_two_chunks_s4": ([(40, 40, 26)], [2100], 4, True, True),          # 100 binned pixels < 250 chunk components
    "tiny_three_views_s2": ([(16, 20, 27), (12, 30, 28), (20, 14, 29)], [1500], 2, True, False),
    # movie SVD without the motion SVD: the reference's merge loop runs over len(U_mot) (process.py:264), so the movie
    # targets beyond it are never merged — all of them without ROIs, the last ROI with ROIs
    "mov_only_s2": ([(36, 44, 4)], [2200], 2, False, True),
    "mov_only_two_rois_s2": ([(36, 44, 5)], [2100], 2, False, True, [motion_roi(0, 4, 16, 4, 20), motion_roi(0, 20, 34, 10, 40)], True),
    "mov_only_rois_no_fullsvd_s2": ([(36, 44, 6)], [2100], 2, False, True, [motion_roi(0, 4, 16, 4, 20), motion_roi(0, 20, 34, 10, 40)], False),
    "roi_single_chunk_s2": ([(40, 56, 1)], [1100], 2, True, True, [motion_roi(0, 4, 20, 6, 30)], True),
    "sixty_frames_s2": ([(30, 40, 6)], [60], 2, True, True),
    "three_frames_s2": ([(30, 40, 8)], [3], 2, True, False),
    "both_svds_off_roi_s2": ([(30, 40, 11)], [1200], 2, False, False, [motion_roi(0, 2, 20, 4, 30)], True),
    "mov_only_single_chunk_roi_s2": ([(30, 40, 10)], [1500], 2, False, True, [motion_roi(0, 2, 20, 4, 30)], True),
    "roi_s6_blocks": ([(48, 60, 14)], [1100, 1000], 6, True, True, [motion_roi(0, 5, 40, 7, 50)], True),
}


def _videos(views, blocks):
    total = int(sum(blocks))
    full = [SynthVideo(H, W, total, seed=seed).all_frames() for (H, W, seed) in views]
    edges = np.concatenate([[0], np.cumsum(blocks)])
    return full, [[v[edges[b]:edges[b + 1]] for v in full] for b in range(len(blocks))]


def _equal(a, b, key):
    if isinstance(a, (list, tuple)):
        assert len(a) == len(b), key
        for i, (x, y) in enumerate(zip(a, b)):
            _equal(x, y, f"{key}[{i}]")
    else:
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, (key, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b), key


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_the_live_reference(name):
    views, blocks, sbin, mot, mov, *rest = CASES[name]
    rois, fullSVD = rest if rest else (None, True)
    full, per_block = _videos(views, blocks)
    kw = dict(sbin=sbin, motSVD=mot, movSVD=mov, rois=rois, fullSVD=fullSVD)
    stock = ref_harness.run_reference(per_block, **kw)
    exact = ref_harness.run_reference(per_block, exact=True, **kw)
    got = orc.run_oracle(full, svd="exact", block_frames=blocks, **kw)
    for key in ("avgframe", "avgmotion", "motion", "avgframe_reshape", "avgmotion_reshape"):
        _equal(got[key], stock[key], key)                     # integer-derived stages: stock reference, bit for bit
    for key in ("motSv", "movSv", "motMask", "movMask", "motMask_reshape", "movMask_reshape", "motSVD", "movSVD"):
        _equal(got[key], exact[key], key)                     # SVD outputs: reference with the exact svdecon
    for key in ("Ly", "Lx", "Lybin", "Lxbin", "sybin", "sxbin", "LYbin", "LXbin"):
        assert np.array_equal(np.asarray(got[key]), np.asarray(stock[key])), key


def _same_tree(a, b, path=""):
    if isinstance(a, dict):
        assert sorted(a) == sorted(b), path
        for k in a:
            _same_tree(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same_tree(x, y, f"{path}[{i}]")
    else:
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, (path, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), path


def _fullres_cases():
    from .helpers import blink_roi, eye_video, pupil_roi, running_roi

    return {
        # ROIs touching the frame border; 640 frames cross one 500-frame seam of process_ROIs
        "border_rois": ([eye_video(50, 64, 640, seed=7)], 2, True, False,
                        [pupil_roi(0, 0, 30, 0, 40, 75.5), blink_roi(0, 20, 50, 30, 64, 150.0), running_roi(0, 36, 50, 0, 64)]),
        # no SVD at all, blink before pupil (paint order), integer blink saturation
        "no_svd_sweep": ([eye_video(48, 60, 530, seed=8)], 1, False, False,
                         [blink_roi(0, 4, 30, 6, 50, 99), pupil_roi(0, 5, 35, 8, 50, 110.2, sigma=3.0)]),
        # two views, processors on different views next to a motion ROI, motion + movie SVD
        "two_views_mixed": ([eye_video(48, 60, 1010, seed=9), eye_video(40, 52, 1010, seed=10, wheel=False)], 2, True, True,
                            [running_roi(0, 34, 48, 2, 58), pupil_roi(1, 4, 34, 6, 46, 85.0), motion_roi(1, 2, 30, 4, 40)]),
    }


@pytest.mark.parametrize("name", ["border_rois", "no_svd_sweep", "two_views_mixed"])
@pytest.mark.filterwarnings("ignore:Mean of empty slice", "ignore:invalid value encountered")
def test_roi_processor_oracle_reproduces_the_live_reference(name):
    """pupil / blink / running (oracle/roi_oracle.py) beyond the three committed fixtures, bit for bit — including what
    the painted frames do to the motion energy and the projections of the same sweep"""
    import copy

    views, sbin, mot, mov, rois = _fullres_cases()[name]
    ref = ref_harness.run_reference([views], sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), exact=True)
    got = orc.run_oracle(views, sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), svd="exact")
    for key in ("pupil", "blink", "running", "motion", "avgframe", "avgmotion", "motSVD", "movSVD", "motMask", "movMask"):
        _same_tree(got[key], ref[key], key)
