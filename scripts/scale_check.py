"""GPU scale check of the other BASELINE configs at reduced frame counts (functional + timing):
  C3-like: 2 cameras 480x640, mot+mov, 3 motion ROIs, sbin 4
  C4-like: 1 camera 1024x1280, sbin 1 (npix = 1.31 M)
Checks: runs to completion, shapes, projection self-consistency on a window, reference row conventions."""
import os, sys, time, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facemap_b200 import process
from facemap_b200.frames import DeviceArraySource
from facemap_b200.synth import SynthVideo
from oracle import facemap_oracle as orc
from tests.helpers import motion_roi

dev = torch.device("cuda", 0)

def gen(H, W, N, seed):
    s = SynthVideo(H, W, N, seed=seed)
    v = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
    step = max(50, min(500, (1 << 27) // (H * W)))
    for t in range(0, N, step):
        v[t:t + step] = s.frames_torch(t, min(N, t + step), dev)
    del s._dev; torch.cuda.empty_cache()
    return v

def window_check(v, sbin, am, U, V, t0, t1, col0=0):
    H, W = v.shape[1:]
    b = orc.bin_frames(v[t0 - 1:t1].cpu().numpy(), sbin, H // sbin, W // sbin)
    return np.abs(np.diff(b, axis=0)) 

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    N = 20000
    vids = [gen(480, 640, N, 0), gen(480, 640, N, 1)]
    rois = [motion_roi(0, 40, 200, 80, 320), motion_roi(1, 32, 160, 64, 300), motion_roi(0, 240, 440, 400, 600)]
    for rep in range(2):
        tm = {}
        torch.cuda.synchronize(); t = time.perf_counter()
        p = process.process_source(DeviceArraySource(vids), sbin=4, motSVD=True, movSVD=True, rois=copy.deepcopy(rois), timing=tm)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"C3-like N={N}: {dt:.3f} s -> {N/dt:.0f} frames/s (2 views)", flush=True)
    print({k: round(x, 1) for k, x in sorted(tm["device_ms"].items()) if x > 2})
    print("shapes motMask", [u.shape for u in p["motMask"]], "movMask", [u.shape for u in p["movMask"]])
    print("motSVD", [u.shape for u in p["motSVD"]], "motion", [m.shape for m in p["motion"]], "motSv", p["motSv"].shape)
    # self-consistency of the full-frame motion projection on a window (both views concatenated)
    t0, t1 = 7000, 7032
    ds = []
    for v in vids:
        b = orc.bin_frames(v[t0 - 1:t1].cpu().numpy(), 4, 120, 160)
        ds.append(np.abs(np.diff(b, axis=0)))
    am = np.concatenate([a.reshape(-1) for a in p["avgmotion"]])
    X = (np.concatenate(ds, axis=1) - am).astype(np.float32).astype(np.float64)
    Vw = X @ p["motMask"][0].astype(np.float64)
    err = np.abs(p["motSVD"][0][t0:t1] - Vw).max() / np.abs(Vw).max()
    e_ref = np.float32(np.float32(ds[0].sum(axis=1)).astype(np.float64) + ds[1].sum(axis=1))
    print(f"projection self-consistency rel err {err:.2e}; motion exact: {np.array_equal(p['motion'][0][t0:t1], e_ref)}")
    # ROI 0 projection
    y, x = p["rois"][0]["yrange_bin"], p["rois"][0]["xrange_bin"]
    sub = ds[0].reshape(-1, 120, 160)[:, y[0]:y[-1] + 1, x[0]:x[-1] + 1]
    amr = p["avgmotion"][0].reshape(120, 160)[y[0]:y[-1] + 1, x[0]:x[-1] + 1]
    Xr = (sub - amr).astype(np.float32).astype(np.float64).reshape(sub.shape[0], -1)
    Vr = Xr @ p["motMask"][1].astype(np.float64)
    print(f"ROI0 projection rel err {np.abs(p['motSVD'][1][t0:t1] - Vr).max() / np.abs(Vr).max():.2e}; ROI motion exact: "
          f"{np.array_equal(p['motion'][1][t0:t1], sub.sum(axis=(1, 2)).astype(np.float32))}")
else:
    N = 3000
    v = gen(1024, 1280, N, 0)
    for rep in range(2):
        tm = {}
        torch.cuda.synchronize(); t = time.perf_counter()
        p = process.process_source(DeviceArraySource([v]), sbin=1, motSVD=True, movSVD=False, timing=tm)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"C4-like N={N}: {dt:.3f} s -> {N/dt:.0f} frames/s", flush=True)
    print({k: round(x, 1) for k, x in sorted(tm["device_ms"].items()) if x > 2})
    print("motMask", p["motMask"][0].shape, "motSVD", p["motSVD"][0].shape)
    t0, t1 = 1500, 1516
    fr = v[t0 - 1:t1].cpu().numpy().astype(np.float32).reshape(t1 - t0 + 1, -1)
    d = np.abs(np.diff(fr, axis=0))
    X = (d - p["avgmotion"][0]).astype(np.float64)
    Vw = X @ p["motMask"][0].astype(np.float64)
    print(f"projection self-consistency rel err {np.abs(p['motSVD'][0][t0:t1] - Vw).max() / np.abs(Vw).max():.2e}")
    U = p["motMask"][0].astype(np.float64)
    print("orth err (first 50)", np.abs(U[:, :50].T @ U[:, :50] - np.eye(50)).max())
