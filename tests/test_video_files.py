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
    big = Sr >= 1e-2 * Sr[0]
    assert (np.abs(S - Sr)[big] / Sr[big]).max() <= 1e-4
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
        big = nrm_r >= 1e-2 * nrm_r[0]
        assert (np.abs(nrm - nrm_r)[big] / nrm_r[big]).max() <= 1e-4, mk
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
