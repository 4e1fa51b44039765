"""Pins the oracle (oracle/facemap_oracle.py) to the UNMODIFIED reference: tests/golden/*.npz hold
outputs of /root/reference/facemap/process.py run in the build container by scripts/make_golden.py
(O1 = stock, O2 = exact-svdecon variant).  Integer-stage outputs must match bit-for-bit; SVD
outputs to tight tolerances (they are bit-identical on the machine that generated them, but BLAS
threading may differ elsewhere)."""
import os

import numpy as np
import pytest

from oracle import facemap_oracle as orc
from tests.helpers import CASES, KEEP, LARGE_CASES, case_rois, case_videos, frame_window


@pytest.mark.parametrize("mode", ["exact", "stock"])
@pytest.mark.parametrize("name", list(CASES) + ["c1_s4"])          # c1_s4 = BASELINE configs[0] at full size
def test_oracle_reproduces_reference(golden_dir, name, mode):
    views, N, sbin, mot, mov, _ = {**CASES, **LARGE_CASES}[name]
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    assert int(z["N"]) == N and int(z["sbin"]) == sbin
    o = orc.run_oracle(case_videos(name), sbin=sbin, motSVD=mot, movSVD=mov, rois=case_rois(name), svd=mode)
    pre = mode + "/"
    for key in ("Lybin", "Lxbin", "sybin", "sxbin"):
        assert np.array_equal(np.asarray(o[key]), z[pre + key]), key
    assert int(o["LYbin"]) == int(z[pre + "LYbin"]) and int(o["LXbin"]) == int(z[pre + "LXbin"])
    for key in ("avgframe", "avgmotion"):
        for i, a in enumerate(o[key]):
            assert np.array_equal(a, z[f"{pre}{key}_{i}"]), (key, i)
        assert np.array_equal(o[key + "_reshape"], z[pre + key + "_reshape"])
    for i, m in enumerate(o["motion"]):
        assert np.array_equal(m, z[f"{pre}motion_{i}"]), f"motion[{i}]"
    w = frame_window(N)
    tol = dict(rtol=2e-3, atol=0) if mode == "stock" else dict(rtol=1e-5, atol=0)
    for sk in ("motSv", "movSv"):
        assert np.allclose(np.asarray(o[sk]), z[pre + sk], **tol), sk
    for mk, vk in (("motMask", "motSVD"), ("movMask", "movSVD")):
        assert len(o[mk]) == int(z[f"{pre}{mk}_n"])
        for i, U in enumerate(o[mk]):
            assert tuple(U.shape) == tuple(z[f"{pre}{mk}_{i}_shape"])
            if U.shape[0] == 0 or len(o[vk]) <= i:
                continue
            Ur = z[f"{pre}{mk}_{i}"]
            lead = min(4, Ur.shape[1])                     # leading components are well separated
            scale = np.abs(Ur[:, :lead]).max()
            assert np.abs(U[:, :lead] - Ur[:, :lead]).max() <= 2e-3 * scale, (mk, i)
            V, Vr = o[vk][i], z[f"{pre}{vk}_{i}"]
            if V.shape[0] == N:
                assert np.abs(V[w][:, :lead] - Vr[:, :lead]).max() <= 2e-3 * np.abs(Vr[:, :lead]).max(), (vk, i)
    if {**CASES, **LARGE_CASES}[name][5] is not None:
        k = 0
        for i, r in enumerate(o["rois"]):
            if r["rind"] == 1:
                assert np.array_equal(r["yrange_bin"], z[f"{pre}roi{i}_yrange_bin"])
                assert np.array_equal(r["xrange_bin"], z[f"{pre}roi{i}_xrange_bin"])
                k += 1


def test_gram_backend_equals_exact_backend():
    """The Gram + eigh formulation the CUDA path uses is the same decomposition as the exact SVD."""
    rng = np.random.RandomState(0)
    X = (rng.randn(400, 60) * np.linspace(5, 0.5, 60)).astype(np.float32)
    for k in (10, 59):
        u1, s1, v1 = orc.svd_centered(X, k, "exact")
        u2, s2, v2 = orc.svd_centered(X, k, "gram")
        assert np.allclose(s1, s2, rtol=1e-5)
        assert np.abs(np.abs(u1[:, :10]) - np.abs(u2[:, :10])).max() < 1e-4
        assert np.array_equal(np.sign(v1[np.arange(10), np.argmax(np.abs(v1[:10]), 1)]), np.ones(10))


def test_stage1_means_at_the_c2_parity_size(golden_dir):
    """BASELINE configs[1] at SURVEY §8(d)'s parity size (640x480, 20 000 frames, sbin 4): the oracle's avgframe /
    avgmotion from the 1 000 sampled frames equal the unmodified reference's, bit for bit (the whole-run comparison
    ran when the fixture was made, scripts/make_golden.py; replaying all of it takes minutes)."""
    from facemap_b200.synth import SynthVideo

    views, N, sbin, *_ = LARGE_CASES["c2_20k_s4"]
    (H, W, seed), = views
    z = np.load(os.path.join(golden_dir, "c2_20k_s4.npz"))
    synth = SynthVideo(H, W, N, seed=seed)

    def get(t0, n):
        a = max(0, min(N - 1, t0))
        b = max(0, min(N - 1, t0 + n - 1))
        return [synth.frames(a, b + 1)]

    import types

    lazy = types.SimpleNamespace(N=N, Ly=[H], Lx=[W], block_frames=[N], get=get)
    avgframe, avgmotion = orc.stage1_means(lazy, sbin)
    assert np.array_equal(avgframe[0], z["stock/avgframe_0"]) and np.array_equal(avgmotion[0], z["stock/avgmotion_0"])
    assert np.array_equal(z["stock/avgmotion_0"], z["exact/avgmotion_0"])
    # what the reference itself resolves at this size: its randomized svdecon agrees with the exact one to 2e-5 on
    # the first 100 singular values and is off by percents in the tail
    S1, S2 = z["stock/motSv"].astype(np.float64), z["exact/motSv"].astype(np.float64)
    rel = np.abs(S1 - S2) / S2
    assert rel[:100].max() < 5e-5 and rel[250:].max() > 1e-2


def test_large_fixture_comparison_on_the_oracle(golden_dir):
    """the comparison the opt-in GPU test applies to the large fixtures (tests/test_large_golden_gpu.py), run here on
    the oracle's Gram back-end for BASELINE configs[0]: keeps that code path exercised on every CPU run"""
    from tests.test_large_golden_gpu import compare_with_fixture

    views, N, sbin, mot, mov, _ = LARGE_CASES["c1_s4"]
    o = orc.run_oracle(case_videos("c1_s4"), sbin=sbin, motSVD=mot, movSVD=mov, svd="gram")
    compare_with_fixture(o, np.load(os.path.join(golden_dir, "c1_s4.npz")), "c1_s4")
