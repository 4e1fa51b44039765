"""BASELINE configs[1] at FULL size (1 camera, 640x480, 100 000 frames, sbin 4, motSVD, 500 components)
checked through size-independent properties, since the oracle cannot finish that size in seconds:

  * stage 3 is a pure function of (frames, avgmotion, masks): a float64 recomputation of the projection and
    of the motion energy on windows of frames (start, chunk seams, end) must match what the full run wrote;
  * reference row conventions at scale: motSVD[0] == motSVD[1] (Q2), motion[499] == 0 and the one-frame shift
    of the first chunk (Q1);
  * masks orthonormal, singular values positive and sorted, mask columns have zero mean (PCA centring);
  * integer stage: avgmotion equals the oracle's stage-1 means computed from the 1000 sampled frames only."""
import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc

pytestmark = pytest.mark.gpu

H, W, N, SBIN = 480, 640, 100_000, 4


@pytest.fixture(scope="module")
def full_run(cuda_device):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    synth = SynthVideo(H, W, N, seed=0)
    video = torch.empty((N, H, W), dtype=torch.uint8, device=cuda_device)
    for t in range(0, N, 500):
        video[t:t + 500] = synth.frames_torch(t, min(N, t + 500), cuda_device)
    del synth._dev
    torch.cuda.empty_cache()
    proc = process.process_source(DeviceArraySource([video]), sbin=SBIN, motSVD=True, movSVD=False)
    return proc, video


def _window_reference(video, t0, t1, avgmotion, U):
    """float64 projection and energy for frames t0..t1-1 (needs frame t0-1)."""
    fr = video[t0 - 1:t1].cpu().numpy()
    b = orc.bin_frames(fr, SBIN, H // SBIN, W // SBIN)
    d = np.abs(np.diff(b, axis=0))
    X = (d - avgmotion).astype(np.float32).astype(np.float64)
    return X @ U.astype(np.float64), d.sum(axis=1).astype(np.float32)


def test_shapes_and_conventions(full_run):
    proc, _ = full_run
    p = (H // SBIN) * (W // SBIN)
    assert proc["motMask"][0].shape == (p, 500) and proc["motSVD"][0].shape == (N, 500)
    assert proc["motSv"].shape == (500,) and proc["motion"][0].shape == (N,)
    V, M = proc["motSVD"][0], proc["motion"][0]
    assert np.array_equal(V[0], V[1])                               # Q2
    assert M[499] == 0 and np.all(M[:499] > 0) and np.all(M[500:] > 0)   # Q1
    S = proc["motSv"]
    assert np.all(S > 0) and np.all(np.diff(S) <= 0)


def test_masks_are_a_centred_orthonormal_basis(full_run):
    proc, _ = full_run
    U = proc["motMask"][0].astype(np.float64)
    G = U.T @ U
    assert np.abs(G - np.eye(500))[:100, :100].max() < 1e-4
    assert np.abs(G - np.eye(500)).max() < 5e-3
    assert np.abs(U.mean(axis=0)).max() < 1e-6                      # columns of Y were centred (sklearn PCA)


def test_projection_and_energy_recomputed_on_windows(full_run):
    proc, video = full_run
    U, am = proc["motMask"][0], proc["avgmotion"][0]
    V, M = proc["motSVD"][0], proc["motion"][0]
    for (t0, t1) in [(1, 40), (480, 520), (18900, 18990), (50000, 50032), (N - 40, N)]:
        Vw, Ew = _window_reference(video, t0, t1, am, U)
        got = V[t0:t1].astype(np.float64)
        assert np.abs(got - Vw).max() <= 1e-5 * np.abs(Vw).max(), (t0, t1)
        # motion: exact (integer energy / 16 -> float32), with the first-chunk shift of Q1
        for i, t in enumerate(range(t0, t1)):
            idx = t - 1 if t < 500 else t
            assert M[idx] == Ew[i], (t, M[idx], Ew[i])


def test_stage1_means_bit_exact_at_full_size(full_run):
    proc, video = full_run
    nt0, tf = orc.stage1_sample_times(N, [N])
    avgframe = np.zeros((H // SBIN) * (W // SBIN), np.float32)
    avgmotion = np.zeros_like(avgframe)
    for t in tf:
        b = orc.bin_frames(video[int(t):int(t) + nt0].cpu().numpy(), SBIN, H // SBIN, W // SBIN)
        avgframe += b.mean(axis=0)
        avgmotion += np.abs(np.diff(b, axis=0)).mean(axis=0)
    avgframe /= float(len(tf))
    avgmotion /= float(len(tf))
    assert np.array_equal(proc["avgframe"][0], avgframe)
    assert np.array_equal(proc["avgmotion"][0], avgmotion)


def test_singular_values_consistent_with_projection_energy(full_run):
    """Rayleigh check tying stage 2 to stage 3: for the 15 x 1000 sampled frames the variance captured by
    component j of the masks, ||X u_j||, must reproduce the spectrum's ordering and scale (leading
    components within 5 %: the masks come from the chunk-compressed data, the check uses raw frames)."""
    proc, video = full_run
    U, am, S = proc["motMask"][0].astype(np.float64), proc["avgmotion"][0], proc["motSv"].astype(np.float64)
    nt0, nsegs, nc, tf = orc.stage2_plan(N)
    acc = np.zeros(8)
    for t in tf:
        b = orc.bin_frames(video[int(t):int(t) + nt0].cpu().numpy(), SBIN, H // SBIN, W // SBIN)
        X = (np.abs(np.diff(b, axis=0)) - am).astype(np.float64)
        X -= X.mean(axis=1, keepdims=True)                           # svdecon's PCA centring per frame
        acc += ((X @ U[:, :8]) ** 2).sum(axis=0)
    assert np.allclose(np.sqrt(acc), S[:8], rtol=5e-2)
