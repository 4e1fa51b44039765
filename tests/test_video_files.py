"""VideoFileSource (the cv2 seam replacing utils.get_frames / get_frame_details, facemap/utils.py:522-591):
decoding, cross-file stitching of sequential blocks, range clamping — CPU part; and the real drop-in call
`process.run([[...avi...]])` on the GPU against the oracle on the same frames."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from facemap_b200.synth import SynthVideo  # noqa: E402
from oracle import facemap_oracle as orc  # noqa: E402


def write_avi(path, frames):
    """lossless gray AVI (FFV1 round-trips bit-exactly through VideoCapture + BGR2GRAY)."""
    T, H, W = frames.shape
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (W, H), isColor=False)
    assert vw.isOpened(), "FFV1 writer unavailable"
    for f in frames:
        vw.write(f)
    vw.release()


@pytest.fixture(scope="module")
def avi_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("avis")
    a = SynthVideo(48, 64, 1300, seed=21).all_frames()
    b = SynthVideo(48, 64, 900, seed=22).all_frames()
    c = SynthVideo(40, 56, 1300, seed=23).all_frames()
    paths = {k: str(d / f"{k}.avi") for k in "abc"}
    for k, arr in zip("abc", (a, b, c)):
        write_avi(paths[k], arr)
    return paths, {"a": a, "b": b, "c": c}


def test_video_file_source_reads_and_stitches(avi_files):
    from facemap_b200.frames import VideoFileSource

    paths, arr = avi_files
    src = VideoFileSource([[paths["a"]], [paths["b"]]])                 # two sequential blocks, one view
    try:
        assert src.nframes == 2200 and src.Ly == [48] and src.Lx == [64] and src.block_frames == [1300, 900]
        whole = np.concatenate((arr["a"], arr["b"]))
        assert np.array_equal(src._read(0, 0, 40), whole[:40])
        assert np.array_equal(src._read(0, 1280, 1330), whole[1280:1330])   # spans the file boundary
        assert np.array_equal(src._read(0, 2150, 2200), whole[2150:2200])
        assert np.array_equal(src._read(0, 100, 130), whole[100:130])       # seek backwards
        assert src.clamp(2100, 2600) == (2100, 2200)
    finally:
        src.close()
    src = VideoFileSource([[paths["a"], paths["c"]]])                   # two simultaneous views
    try:
        assert src.nframes == 1300 and src.Ly == [48, 40] and src.Lx == [64, 56]
        assert np.array_equal(src._read(1, 500, 520), arr["c"][500:520])
    finally:
        src.close()


@pytest.mark.gpu
def test_run_on_video_files_matches_oracle(cuda_device, avi_files, tmp_path):
    """process.run(filenames, ...) with real video files == the oracle on the decoded frames."""
    from facemap_b200 import process

    paths, arr = avi_files
    # (1) sequential blocks: frames concatenated in time (utils.get_frame_details cumframes)
    name = process.run([[paths["a"]], [paths["b"]]], sbin=2, motSVD=True, movSVD=False, savepath=str(tmp_path))
    assert name == os.path.join(str(tmp_path), "a_proc.npy")
    d = np.load(name, allow_pickle=True).item()
    whole = np.concatenate((arr["a"], arr["b"]))
    o = orc.run_oracle([whole], sbin=2, motSVD=True, movSVD=False, svd="exact", block_frames=[1300, 900])
    assert d["filenames"] == [[paths["a"]], [paths["b"]]]
    assert np.array_equal(d["avgframe"][0], o["avgframe"][0]) and np.array_equal(d["avgmotion"][0], o["avgmotion"][0])
    assert np.array_equal(d["motion"][0], o["motion"][0])
    S, Sr = d["motSv"].astype(np.float64), o["motSv"].astype(np.float64)
    assert (np.abs(S - Sr) / Sr).max() <= 1e-4                      # all 500 singular values
    U, Ur = d["motMask"][0][:, :4].astype(np.float64), o["motMask"][0][:, :4].astype(np.float64)
    assert np.abs(np.abs((U * Ur).sum(axis=0)) - 1).max() < 1e-5
    assert np.abs(d["motSVD"][0][:, :4] - o["motSVD"][0][:, :4]).max() <= 1e-3 * np.abs(o["motSVD"][0][:, :4]).max()
    # (2) two simultaneous views of different size
    name = process.run([[paths["a"], paths["c"]]], sbin=4, motSVD=True, movSVD=True, savepath=str(tmp_path))
    d = np.load(name, allow_pickle=True).item()
    o = orc.run_oracle([arr["a"], arr["c"]], sbin=4, motSVD=True, movSVD=True, svd="exact")
    for key in ("avgframe", "avgmotion", "motion"):
        for x, y in zip(d[key], o[key]):
            assert np.array_equal(x, y), key
    assert np.array_equal(d["motMask_reshape"][0].shape, o["motMask_reshape"][0].shape)
    assert np.array_equal(np.asarray(d["sybin"]), o["sybin"]) and np.array_equal(np.asarray(d["sxbin"]), o["sxbin"])
    # 1300 frames = one SVD chunk: the reference returns U*S as the masks and zero singular values (quirk Q6)
    for sk, mk in (("motSv", "motMask"), ("movSv", "movMask")):
        assert not np.any(o[sk]) and not np.any(d[sk]), sk
        U, Ur = d[mk][0].astype(np.float64), o[mk][0].astype(np.float64)
        assert U.shape == Ur.shape
        nrm, nrm_r = np.linalg.norm(U, axis=0), np.linalg.norm(Ur, axis=0)        # = singular values of the chunk
        # every component; the last one of a centred 192-pixel view is numerically zero, and a zero singular value
        # comes out of a float64 Gram matrix as sqrt(n * 2^-53) ~ 1e-6 of the largest
        assert np.all(np.abs(nrm - nrm_r) <= np.maximum(1e-4 * nrm_r, 2e-6 * nrm_r[0])), mk
        cos = np.abs((U[:, :4] * Ur[:, :4]).sum(axis=0)) / (nrm[:4] * nrm_r[:4])
        assert np.abs(cos - 1).max() < 1e-5, mk


@pytest.mark.gpu
def test_decode_cache_equals_on_demand_decoding(cuda_device, avi_files):
    """the decode-once HBM cache (threaded decoders, priority pieces first) hands out the same frames as the
    on-demand path, for ranges in any order, across a file boundary, and end to end"""
    import torch

    from facemap_b200 import process
    from facemap_b200.frames import VideoFileSource

    paths, arr = avi_files
    whole = np.concatenate((arr["a"], arr["b"]))
    src = VideoFileSource([[paths["a"]], [paths["b"]]], decode_workers=5)
    try:
        src.plan_access([(2000, 2100), (0, 100), (1250, 1350)], (0, src.nframes), cuda_device)
        assert src._cache is not None
        for a, b in ((1250, 1350), (0, 100), (2100, 2200), (255, 258), (1299, 1301), (0, 2200)):
            got = src.fetch(a, b, cuda_device)[0]
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), whole[a:b]), (a, b)
        assert src._cache.decoded_frames == src.nframes            # every frame decoded exactly once
    finally:
        src.close()
    outs = []
    for cached in (True, False):
        src = VideoFileSource([[paths["a"], paths["c"]]], cache_on_device=cached)
        try:
            outs.append(process.process_source(src, sbin=2, motSVD=True, movSVD=True))
            assert (src._cache is not None) == cached
        finally:
            src.close()
    for key in ("avgframe", "avgmotion", "motion", "motMask", "movMask", "motSVD", "movSVD"):
        for x, y in zip(outs[0][key], outs[1][key]):
            assert np.array_equal(x, y), key


def test_decode_cache_run_plan():
    """sample ranges first (in access order, consecutive pieces joined into one run = one seek), then the rest of
    the range as long spans in frame order; every piece exactly once"""
    from facemap_b200.frames import _DecodeCache

    runs = _DecodeCache.plan_runs([(2000, 2100), (0, 100), (1250, 1350), (300, 700)], (0, 10_000), 256, 1, 4)
    assert runs[:4] == [(7, 9), (0, 1), (4, 6), (1, 3)]
    pieces = [pc for a, b in runs for pc in range(a, b)]
    assert sorted(pieces) == list(range(40)) and len(pieces) == 40
    sweep = runs[4:]
    assert sweep == sorted(sweep) and len(sweep) <= 4 + 3                # <= workers spans, cut at the 3 holes
    assert max(b - a for a, b in sweep) == 9                             # ceil(33 pieces / 4 threads)
    # two views share the threads: spans half as many per view
    runs2 = _DecodeCache.plan_runs([], (500, 3000), 256, 2, 4)
    assert runs2 == [(0, 5), (5, 10)]


class _CoarseSeekCapture:
    """cv2.VideoCapture duck type of an inter-frame codec whose seeks land on the previous `gop`-th frame while the
    capture REPORTS the requested position: only decoding shows it."""

    def __init__(self, arr, gop=8):
        self.arr, self.pos, self.said, self.gop = arr, 0, 0, gop

    def isOpened(self):
        return True

    def get(self, prop):
        return {cv2.CAP_PROP_POS_FRAMES: self.said, cv2.CAP_PROP_FRAME_COUNT: self.arr.shape[0],
                cv2.CAP_PROP_FRAME_HEIGHT: self.arr.shape[1], cv2.CAP_PROP_FRAME_WIDTH: self.arr.shape[2]}.get(prop, 0.0)

    def set(self, prop, value):
        if prop == cv2.CAP_PROP_POS_FRAMES:
            self.said = int(value)
            self.pos = int(value) // self.gop * self.gop
        return True

    def read(self):
        if self.pos >= self.arr.shape[0]:
            return False, None
        f = self.arr[self.pos]
        self.pos += 1
        self.said += 1
        return True, f

    def release(self):
        pass


@pytest.mark.gpu
@pytest.mark.parametrize("gop", [1, 7])      # 7: no piece boundary (multiples of 256) is a key frame
def test_sweep_never_uses_inexact_seeks(cuda_device, gop):
    """a capture whose seeks land up to gop-1 frames early (and hide it): the decode-once cache notices at the run
    seams and stage 3's sweep is decoded sequentially instead; with exact seeks (gop 1) the cache serves the sweep"""
    import torch

    from facemap_b200.frames import VideoFileSource

    rs = np.random.RandomState(8)
    arr = rs.randint(0, 256, size=(1500, 12, 16)).astype(np.uint8)

    class Src(VideoFileSource):
        def _open(self):
            return [[_CoarseSeekCapture(arr, gop)]]

    src = Src([["a.avi"]], decode_workers=3)
    try:
        src.plan_access([(1000, 1100), (300, 420)], (0, 1500), cuda_device)
        assert src._cache is not None and len(src._cache.runs) >= 4
        src.begin_sweep()
        assert src.seek_inexact == (gop > 1) and (src._cache is None) == (gop > 1)
        got = torch.cat([src.fetch(t, min(1500, t + 333), cuda_device)[0] for t in range(0, 1500, 333)])
        assert np.array_equal(got.cpu().numpy(), arr)                   # the sweep is the sequential decode
    finally:
        src.close()


def write_mp4v(path, frames):
    T, H, W = frames.shape
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (W, H), isColor=False)
    if not vw.isOpened():
        return False
    for f in frames:
        vw.write(f)
    vw.release()
    return True


@pytest.mark.gpu
def test_decode_cache_on_an_inter_frame_codec(cuda_device, tmp_path):
    """MPEG-4 part 2 (inter-frame, lossy): what the sweep gets after begin_sweep equals ONE sequential decode of the
    file, whether or not OpenCV's seeks into it are frame-accurate"""
    import torch

    from facemap_b200.frames import VideoFileSource

    frames = SynthVideo(64, 96, 1400, seed=31).all_frames()
    path = str(tmp_path / "inter.mp4")
    if not write_mp4v(path, frames):
        pytest.skip("mp4v writer unavailable")
    cap = cv2.VideoCapture(path)
    seq = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        seq.append(f if f.ndim == 2 else cv2.cvtColor(f, cv2.COLOR_BGR2GRAY))
    cap.release()
    seq = np.stack(seq)
    src = VideoFileSource([[path]], decode_workers=4)
    try:
        assert src.nframes == seq.shape[0]
        src.plan_access([(900, 1000), (200, 300)], (0, src.nframes), cuda_device)
        src.begin_sweep()
        got = torch.cat([src.fetch(t, min(src.nframes, t + 500), cuda_device)[0] for t in range(0, src.nframes, 500)])
        assert np.array_equal(got.cpu().numpy(), seq), f"seek_inexact={src.seek_inexact}"
    finally:
        src.close()


class _FlakyCapture:
    """cv2.VideoCapture duck type over an array whose decoder fails from frame `fail_from` on (a truncated file:
    the header promises more frames than can be decoded)."""

    def __init__(self, arr, fail_from=None):
        self.arr, self.pos, self.fail_from = arr, 0, arr.shape[0] if fail_from is None else fail_from

    def isOpened(self):
        return True

    def get(self, prop):
        return {cv2.CAP_PROP_POS_FRAMES: self.pos, cv2.CAP_PROP_FRAME_COUNT: self.arr.shape[0],
                cv2.CAP_PROP_FRAME_HEIGHT: self.arr.shape[1], cv2.CAP_PROP_FRAME_WIDTH: self.arr.shape[2]}.get(prop, 0.0)

    def set(self, prop, value):
        if prop == cv2.CAP_PROP_POS_FRAMES:
            self.pos = int(value)
        return True

    def read(self):
        if self.pos >= self.fail_from:
            return False, None
        f = self.arr[self.pos]
        self.pos += 1
        return True, np.repeat(f[:, :, None], 3, axis=2)

    def release(self):
        pass


def _flaky_blocks():
    rs = np.random.RandomState(4)
    a = rs.randint(1, 256, size=(60, 12, 16)).astype(np.uint8)         # no zero pixels: a zero frame is a hole
    b = rs.randint(1, 256, size=(50, 12, 16)).astype(np.uint8)
    return [(a, 37), (b, None)]                                        # block 0 stops decoding at frame 37 of 60


_FLAKY_REQUESTS = [(0, 30), (20, 30), (36, 10), (37, 10), (45, 10), (30, 60), (55, 20), (100, 30), (-5, 20)]


def test_decode_failure_semantics_match_get_frames():
    """a frame that fails to decode repeats its predecessor once, the rest of that file's range stays zero, the next
    sequential file is read normally (facemap/utils.py:536-553) — against the oracle's restatement of get_frames"""
    from facemap_b200.frames import VideoFileSource

    blocks = _flaky_blocks()

    class Src(VideoFileSource):
        def _open(self):
            return [[_FlakyCapture(arr, ff)] for arr, ff in blocks]

    src = Src([["a.avi"], ["b.avi"]])
    try:
        assert src.nframes == 110 and src.block_frames == [60, 50]
        stage = np.full((64, 12, 16), 7, np.uint8)                      # a reused (dirty) staging buffer
        for t0, n in _FLAKY_REQUESTS:
            want = orc.read_frames([[_FlakyCapture(arr, ff)] for arr, ff in blocks], src.cumframes, [(12, 16)], t0, n)[0]
            a, b = src.clamp(t0, t0 + n)
            got = src._read(0, a, b)
            assert got.shape == want.shape and np.array_equal(got, want), (t0, n)
            got2 = src._read(0, a, b, out=stage[:b - a])
            assert np.array_equal(got2, want), (t0, n)
        whole = src._read(0, 0, 110)
        assert np.array_equal(whole[:37], blocks[0][0][:37]) and np.array_equal(whole[37], blocks[0][0][36])
        assert not whole[38:60].any() and np.array_equal(whole[60:], blocks[1][0])
    finally:
        src.close()


def test_oracle_read_frames_is_the_reference_get_frames():
    """pins oracle.read_frames to the unmodified utils.get_frames (only where /root/reference exists)"""
    from oracle import ref_harness

    if not ref_harness.reference_available():
        pytest.skip("reference tree not present on this machine")
    _, ref_utils = ref_harness._import_reference()
    blocks = _flaky_blocks()
    cum = np.array([0, 60, 110])
    import contextlib
    import io

    for t0, n in _FLAKY_REQUESTS:
        imall = [np.zeros((n, 12, 16), np.uint8)]
        caps = [[_FlakyCapture(arr, ff)] for arr, ff in blocks]
        with contextlib.redirect_stdout(io.StringIO()):
            ref_utils.get_frames(imall, caps, np.arange(t0, t0 + n), cum)
        want = orc.read_frames([[_FlakyCapture(arr, ff)] for arr, ff in blocks], cum, [(12, 16)], t0, n)[0]
        assert imall[0].shape == want.shape and np.array_equal(imall[0], want), (t0, n)
