"""Whole path on the GPU (facemap_b200.process.process_source) against
  (a) the committed golden fixtures = outputs of the UNMODIFIED reference run in the build
      container (scripts/make_golden.py; "stock" = O1, "exact" = O2 of SURVEY.md §8c), and
  (b) the oracle restatement run live on the same bytes.

Gates (north_star): geometry exact; avgframe / avgmotion / motion bit-exact (ROI motion of a
non-power-of-two sbin: <= 1 ulp); singular values <= 1e-4 relative; sign-aligned masks with principal-angle
sine <= 1e-3; projected traces <= 1e-3 relative — the SVD gates on the spectrally resolved
leading components against O2, and on the components O1 itself resolves against O1 (F3)."""
import os

import numpy as np
import pytest
import torch

from oracle import facemap_oracle as orc
from tests.helpers import CASES, KEEP, case_rois, case_videos, frame_window, gap_resolved, sign_align, subspace_sin

pytestmark = pytest.mark.gpu

S_RTOL = 1e-4
ANGLE_TOL = 1e-3
TRACE_RTOL = 1e-3


def _run_cuda(name, device, source_kind="device", **kw):
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource, HostArraySource

    views, N, sbin, mot, mov, _ = CASES[name]
    vids = case_videos(name)
    if source_kind == "device":
        src = DeviceArraySource([torch.from_numpy(v).to(device) for v in vids])
    else:
        src = HostArraySource(vids)
    return process.process_source(src, sbin=sbin, motSVD=mot, movSVD=mov, rois=case_rois(name), **kw), vids


def svd_metrics(S_ref, U_ref, V_ref, S, U, V, kmax):
    """relative S error, per-prefix subspace sine and sign-aligned trace error for the first kmax
    components (U: [p, >=kmax] masks, V: [frames, >=kmax] traces)."""
    kmax = min(kmax, U_ref.shape[1], U.shape[1])
    out = {"S_rel": None, "angle": [], "trace_rel": [], "sign_agree": 0}
    if S_ref is not None and len(S_ref) >= kmax and np.any(S_ref):
        out["S_rel"] = np.abs(S[:kmax].astype(np.float64) - S_ref[:kmax]) / S_ref[:kmax]
    Ua, sg = sign_align(U[:, :kmax], U_ref[:, :kmax])
    out["sign_agree"] = int((sg > 0).sum())
    for k in range(1, kmax + 1):
        out["angle"].append(subspace_sin(U[:, :k], U_ref[:, :k]))
    Va = V[:, :kmax] * sg[None, :]
    num = np.linalg.norm(Va.astype(np.float64) - V_ref[:, :kmax], axis=0)
    den = np.linalg.norm(V_ref[:, :kmax].astype(np.float64), axis=0) + 1e-30
    out["trace_rel"] = num / den
    return out


def leading_resolved(S, angle_budget=ANGLE_TOL, eps=2e-7):
    """How many leading components a float32-grade Gram can pin to `angle_budget`: the subspace of
    the first k components moves by ~ eps * S0^2 / (S_k^2 - S_{k+1}^2)."""
    S = np.asarray(S, np.float64)
    k_ok = 0
    for k in range(1, len(S)):
        gap = S[k - 1] ** 2 - S[k] ** 2
        if gap <= 0 or eps * S[0] ** 2 / gap > angle_budget / 4:
            break
        k_ok = k
    return k_ok


@pytest.mark.parametrize("name", list(CASES))
def test_integer_stages_match_reference_golden(cuda_device, golden_dir, name):
    """geometry, avgframe, avgmotion, motion vs the unmodified reference (both oracles agree here)."""
    proc, _ = _run_cuda(name, cuda_device)
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sbin = CASES[name][2]
    pow2 = sbin & (sbin - 1) == 0
    for key in ("Lybin", "Lxbin", "sybin", "sxbin"):
        assert np.array_equal(np.asarray(proc[key]), z["stock/" + key]), key
    assert int(proc["LYbin"]) == int(z["stock/LYbin"]) and int(proc["LXbin"]) == int(z["stock/LXbin"])
    for key in ("avgframe", "avgmotion"):
        for i, a in enumerate(proc[key]):
            ref = z[f"stock/{key}_{i}"]
            assert a.shape == ref.shape, (key, i, a.shape, ref.shape)      # quirk Q4 included
            # bit-exact for every sbin: integer time sums for powers of two, the reference's float64
            # arithmetic replayed (fm_mean_ref) otherwise
            assert np.array_equal(a, ref), (key, i)
        ref = z[f"stock/{key}_reshape"]
        got = proc[f"{key}_reshape"]
        assert got.shape == ref.shape and np.array_equal(got, ref)
    for i, m in enumerate(proc["motion"]):
        ref = z[f"stock/motion_{i}"]
        assert m.shape == ref.shape
        if pow2 or i == 0:                                                  # full frame: fm_motion_ref for other sbin
            assert np.array_equal(m, ref), f"motion[{i}]"                   # quirk Q1 included
        else:                                                               # ROI energies of odd bins: <= 1 ulp
            assert np.allclose(m, ref, rtol=2e-7, atol=0), f"motion[{i}]"


@pytest.mark.parametrize("name", list(CASES))
def test_svd_outputs_match_exact_oracle_golden(cuda_device, golden_dir, name):
    """S, masks and traces vs O2 (reference with an exact float64 svdecon), leading components."""
    proc, _ = _run_cuda(name, cuda_device)
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    N = CASES[name][1]
    w = frame_window(N)
    for mk, vk, sk in (("motMask", "motSVD", "motSv"), ("movMask", "movSVD", "movSv")):
        n_ref = int(z[f"exact/{mk}_n"])
        assert len(proc[mk]) == n_ref, mk
        for i in range(n_ref):
            shape = tuple(z[f"exact/{mk}_{i}_shape"])
            assert tuple(proc[mk][i].shape) == shape, (mk, i, proc[mk][i].shape, shape)
            if shape[0] == 0 or not CASES[name][3 if mk == "motMask" else 4]:
                continue
            U_ref, V_ref = z[f"exact/{mk}_{i}"], z[f"exact/{vk}_{i}"]
            assert tuple(proc[vk][i].shape) == tuple(z[f"exact/{vk}_{i}_shape"])
            S_ref = z[f"exact/{sk}"] if i == n_ref - 1 else None          # quirk Q3: last target wins
            S = np.asarray(proc[sk])
            if S_ref is not None:
                assert S.shape == S_ref.shape
            # every stored component (KEEP leading ones) whose subspace / direction the data determine
            # (tests/helpers.gap_resolved); without the target's own spectrum (quirk Q3) a short prefix
            kmax = min(KEEP, U_ref.shape[1])
            met = svd_metrics(S_ref, U_ref, V_ref, S, proc[mk][i], proc[vk][i][w], kmax)
            if S_ref is not None and np.any(S_ref):
                comp, pre = gap_resolved(S_ref, ANGLE_TOL)
                comp, pre = comp[:kmax], pre[:kmax]
            else:
                comp = pre = np.arange(kmax) < 6
            assert pre.sum() >= min(3, kmax), (mk, i, "fewer than 3 resolved prefixes")
            if met["S_rel"] is not None:
                assert met["S_rel"].max() <= S_RTOL, (mk, i, met["S_rel"])
            assert np.asarray(met["angle"])[pre].max() <= ANGLE_TOL, (mk, i, met["angle"])
            assert met["trace_rel"][comp].max() <= TRACE_RTOL, (mk, i, met["trace_rel"])
    # all singular values, not only the leading ones: every non-zero one within the north-star 1e-4
    for sk in ("motSv", "movSv"):
        S_ref, S = z["exact/" + sk].astype(np.float64), np.asarray(proc[sk]).astype(np.float64)
        if np.any(S_ref):
            nz = S_ref > 0
            rel = np.abs(S - S_ref)[nz] / S_ref[nz]
            print(name, sk, "S rel err over all", int(nz.sum()), "components:", float(rel.max()))
            assert rel.max() <= S_RTOL, (sk, float(rel.max()), int(np.argmax(rel)))


@pytest.mark.parametrize("name", ["single_s2", "multi_roi_s4", "crop_s3"])
def test_svd_outputs_match_stock_reference_on_resolved_components(cuda_device, golden_dir, name):
    """Against O1 = the UNMODIFIED reference (sklearn randomized svdecon): on the leading components that
    the stock solver itself resolves (K_res: stock and exact reference agree to the north-star tolerances,
    SURVEY F3 / section 8c) the CUDA path must agree with the stock outputs to the same tolerances (x2: both
    sides carry their own error against the exact decomposition)."""
    proc, _ = _run_cuda(name, cuda_device)
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    N = CASES[name][1]
    w = frame_window(N)
    checked = 0
    for mk, vk, sk, on in (("motMask", "motSVD", "motSv", CASES[name][3]), ("movMask", "movSVD", "movSv", CASES[name][4])):
        if not on:
            continue
        n_ref = int(z[f"stock/{mk}_n"])
        for i in range(n_ref):
            if tuple(z[f"stock/{mk}_{i}_shape"])[0] == 0:
                continue
            Us, Ue = z[f"stock/{mk}_{i}"], z[f"exact/{mk}_{i}"]
            Vs = z[f"stock/{vk}_{i}"]
            # K_res: longest prefix on which the stock and the exact reference span the same subspace
            kres = 0
            for k in range(1, Us.shape[1] + 1):
                if subspace_sin(Us[:, :k], Ue[:, :k]) > ANGLE_TOL:
                    break
                kres = k
            if i == n_ref - 1 and np.any(z["stock/" + sk]):
                Ss, Se = z["stock/" + sk].astype(np.float64), z["exact/" + sk].astype(np.float64)
                ks = 0     # resolved by the stock solver itself
                while ks < len(Ss) and abs(Ss[ks] - Se[ks]) / Se[ks] <= S_RTOL:
                    ks += 1
                S = np.asarray(proc[sk], np.float64)
                assert ks >= 3 and (np.abs(S - Ss)[:ks] / Ss[:ks]).max() <= 2 * S_RTOL, (sk, ks)
            assert kres >= 3, (name, mk, i, "stock reference resolves fewer than 3 components")
            met = svd_metrics(None, Us, Vs, None, proc[mk][i], proc[vk][i][w], kres)
            assert max(met["angle"]) <= 2 * ANGLE_TOL, (mk, i, kres, met["angle"])
            assert met["trace_rel"].max() <= 2 * TRACE_RTOL, (mk, i, kres, met["trace_rel"])
            checked += 1
    assert checked >= 1


def test_against_live_oracle_all_components(cuda_device):
    """All 500 components vs the oracle run live (exact float64 svdecon): S everywhere above the
    float32 floor, mask subspace of the resolved prefix, traces; orthonormality of the masks."""
    from facemap_b200 import process
    from facemap_b200.frames import DeviceArraySource
    from facemap_b200.synth import SynthVideo

    H, W, N, sbin = 96, 128, 2600, 2
    v = SynthVideo(H, W, N, seed=5).all_frames()
    o = orc.run_oracle([v], sbin=sbin, motSVD=True, movSVD=False, svd="exact")
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(cuda_device)]), sbin=sbin)
    assert np.array_equal(p["avgmotion"][0], o["avgmotion"][0])
    assert np.array_equal(p["motion"][0], o["motion"][0])
    S_ref, S = o["motSv"].astype(np.float64), p["motSv"].astype(np.float64)
    assert S.shape == S_ref.shape == (500,)
    assert (np.abs(S - S_ref) / S_ref).max() <= S_RTOL                 # all 500
    U, U_ref = p["motMask"][0], o["motMask"][0]
    assert U.shape == U_ref.shape
    gram = U.astype(np.float64).T @ U.astype(np.float64)
    assert np.abs(gram - np.eye(500))[:100, :100].max() < 1e-4        # resolved components: orthonormal
    assert np.abs(gram - np.eye(500)).max() < 5e-3                     # float32-Gram floor on the tail
    kres = max(3, leading_resolved(S_ref))
    met = svd_metrics(S_ref, U_ref, o["motSVD"][0], S, U, p["motSVD"][0], kres)
    assert max(met["angle"]) <= ANGLE_TOL, met["angle"]
    assert met["trace_rel"].max() <= TRACE_RTOL, met["trace_rel"]
    # the projection is self-consistent for EVERY component: V == X @ U recomputed in float64
    # on a window of frames (stage 3 does not depend on which basis of a near-degenerate pair won)
    Lyb, Lxb = H // sbin, W // sbin
    b = orc.bin_frames(v[999:1031], sbin, Lyb, Lxb)
    X = (np.abs(np.diff(b, axis=0)) - o["avgmotion"][0]).astype(np.float32).astype(np.float64)
    Vw = X @ U.astype(np.float64)
    got = p["motSVD"][0][1000:1031].astype(np.float64)
    # in-kernel chains of 128 K (48 truncating accumulations x 5.4e-8) + float32 rounding of the register sums
    assert np.abs(got - Vw).max() <= 1e-5 * np.abs(Vw).max()


def test_host_source_equals_device_source(cuda_device):
    """Frames streamed from host memory give bit-identical outputs to HBM-resident frames."""
    a, _ = _run_cuda("short_s2", cuda_device, "device")
    b, _ = _run_cuda("short_s2", cuda_device, "host")
    for key in ("avgframe", "avgmotion", "motion", "motMask", "motSVD", "movMask", "movSVD"):
        for x, y in zip(a[key], b[key]):
            assert np.array_equal(x, y), key
    assert np.array_equal(a["motSv"], b["motSv"])


def test_result_buffers_are_pinned_and_reused_across_calls(cuda_device):
    """The page-locked result buffers come from a pool that persists across calls (process._PinnedPool): a call whose
    predecessor's results were dropped locks no new memory, a call whose predecessor's results are still held never
    writes into them, and what is handed out really is page-locked (asynchronous device-to-host copies)."""
    import gc

    from facemap_b200 import process

    assert process._HostSink.alloc((1 << 20,)).is_pinned()
    a, _ = _run_cuda("short_s2", cuda_device, "device")
    keep = {k: [np.array(x, copy=True) for x in a[k]] for k in ("motMask", "motSVD")}
    b, _ = _run_cuda("short_s2", cuda_device, "device")            # `a` is alive: nothing of it may be reused
    for k in keep:
        for x, y, z in zip(a[k], keep[k], b[k]):
            assert np.array_equal(x, y) and np.array_equal(x, z), k
            assert not np.shares_memory(x, z), k
    del a, b
    gc.collect()
    before = dict(process._POOL.stats)
    c, _ = _run_cuda("short_s2", cuda_device, "device")
    assert process._POOL.stats["allocated"] == before["allocated"], (before, process._POOL.stats)
    assert process._POOL.stats["reused"] > before["reused"]
    for k in keep:
        for x, y in zip(c[k], keep[k]):
            assert np.array_equal(x, y), k


def test_streaming_host_source_equals_cached(cuda_device):
    """Videos that do not fit HBM stream through pinned buffers chunk by chunk (prefetch on a copy stream);
    forcing that path must give bit-identical outputs to the upload-once cache."""
    from facemap_b200 import process
    from facemap_b200.frames import HostArraySource

    views, N, sbin, mot, mov, _ = CASES["single_s2"]
    vids = case_videos("single_s2")
    outs = []
    for cache in (True, False):
        src = HostArraySource(vids, cache_on_device=cache)
        outs.append(process.process_source(src, sbin=sbin, motSVD=mot, movSVD=mov))
        assert src.h2d > 0
        src.close()
    for key in ("avgframe", "avgmotion", "motion", "motMask", "motSVD", "movMask", "movSVD"):
        for x, y in zip(outs[0][key], outs[1][key]):
            assert np.array_equal(x, y), key


def test_run_with_proc_settings_and_mat_output(cuda_device, tmp_path):
    """`proc=` overrides the keyword arguments (process.py:701-708) and save_mat writes the flattened .mat
    (process.py:615-634)."""
    import scipy.io

    from facemap_b200 import process
    from tests.helpers import motion_roi

    vids = case_videos("short_s2")
    settings = {"sbin": 2, "fullSVD": True, "save_mat": True, "rois": [motion_roi(0, 4, 30, 6, 40)], "sy": 3, "sx": 5,
                "savepath": str(tmp_path)}
    name = process.run([torch.from_numpy(v) for v in vids], sbin=7, motSVD=True, movSVD=False, proc=settings)
    d = np.load(name, allow_pickle=True).item()
    assert d["sbin"] == 2 and d["sy"] == 3 and d["sx"] == 5 and d["save_mat"] is True
    assert len(d["motMask"]) == 2 and len(d["motSVD"]) == 2 and len(d["motion"]) == 2
    assert "yrange_bin" in d["rois"][0] and d["avgmotion"][0].ndim == 2          # quirks: rois mutated, Q4
    mat = scipy.io.loadmat(os.path.join(str(tmp_path), "array_input_proc.mat"))
    for key in ("motSVD_0", "motSVD_1", "motMask_0", "motion_0", "avgframe_0", "motSv"):
        assert key in mat, key
    assert mat["motSVD_0"].shape == d["motSVD"][0].shape


@pytest.mark.parametrize("case,mot,mov", [("single_s2", True, False), ("short_s2", True, True), ("multi_roi_s4", False, True),
                                          ("crop_s3", True, False), ("tiny_onechunk_s2", True, True)])
def test_run_writes_the_file_while_the_pass_runs(cuda_device, tmp_path, case, mot, mov):
    """run() streams the big arrays into the file during the pass (process._EarlyFile).  Whatever went to disk early,
    the file must hold exactly what process_source returns for the same input — every key, bit for bit.  (One and
    two views, motion / movie / both, a non-power-of-two bin factor, the single-chunk regime; no ROIs.)"""
    from facemap_b200 import process

    vids = [torch.from_numpy(v) for v in case_videos(case)]
    sbin = CASES[case][2]
    name = process.run(vids, sbin=sbin, motSVD=mot, movSVD=mov, savepath=str(tmp_path))
    split = dict(process.LAST_RUN)
    assert split["written_early"] is True and split["early_bytes"] > 0, split
    assert sorted(os.listdir(tmp_path)) == ["array_input_proc.npy"]
    d = np.load(name, allow_pickle=True).item()
    want = process.process_source(vids, sbin=sbin, motSVD=mot, movSVD=mov)
    for key, val in want.items():
        if isinstance(val, list):
            assert len(d[key]) == len(val), key
            for a, b in zip(d[key], val):
                assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b), key
        elif isinstance(val, np.ndarray):
            assert d[key].dtype == val.dtype and np.array_equal(d[key], val), key
    # and it is the file the plain writer produces for the same dict
    plain = str(tmp_path / "plain.npy")
    process.save_npy(plain, d)
    assert os.path.getsize(plain) == os.path.getsize(name)


def test_run_with_motion_rois_falls_back_to_the_plain_writer(cuda_device, tmp_path):
    from facemap_b200 import process
    from tests.helpers import motion_roi

    vids = [torch.from_numpy(v) for v in case_videos("short_s2")]
    settings = {"sbin": 2, "fullSVD": True, "save_mat": False, "rois": [motion_roi(0, 4, 30, 6, 40)], "sy": 0, "sx": 0,
                "savepath": str(tmp_path)}
    name = process.run(vids, proc=settings)
    assert process.LAST_RUN["written_early"] is False
    assert sorted(os.listdir(tmp_path)) == ["array_input_proc.npy"]
    assert len(np.load(name, allow_pickle=True).item()["motSVD"]) == 2


def test_run_writes_reference_proc_file(cuda_device, tmp_path):
    """process.run: same signature / file naming / keys as facemap.process.run (process.py:605-647, 859-892)."""
    from facemap_b200 import process

    vids = case_videos("short_s2")
    name = process.run([torch.from_numpy(v) for v in vids], sbin=2, motSVD=True, movSVD=True, savepath=str(tmp_path))
    assert name == os.path.join(str(tmp_path), "array_input_proc.npy")
    d = np.load(name, allow_pickle=True).item()
    expect = {"filenames", "save_path", "Ly", "Lx", "sbin", "fullSVD", "save_mat", "Lybin", "Lxbin", "sybin", "sxbin",
              "LYbin", "LXbin", "avgframe", "avgmotion", "avgframe_reshape", "avgmotion_reshape", "motion", "motSv",
              "movSv", "motMask", "movMask", "motMask_reshape", "movMask_reshape", "motSVD", "movSVD", "pupil",
              "running", "blink", "rois", "sy", "sx"}
    assert set(d.keys()) == expect
    assert d["motSVD"][0].dtype == np.float32 and d["motSVD"][0].shape[0] == CASES["short_s2"][1]
    assert np.array_equal(d["motSVD"][0][0], d["motSVD"][0][1])            # quirk Q2
