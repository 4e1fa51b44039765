"""BASELINE configs[4]: frame-count / resolution sweep on one GPU (frames resident in HBM), motSVD, 500 components.
Writes gpurun_out/sweep.json; one line per (H, W, N, sbin)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facemap_b200 import process
from facemap_b200.frames import DeviceArraySource
from facemap_b200.synth import SynthVideo

dev = torch.device("cuda", 0)
CONFIGS = [(128, 128, 10_000, 4), (128, 128, 100_000, 4), (128, 128, 1_000_000, 4),
           (256, 256, 10_000, 4), (256, 256, 100_000, 4), (256, 256, 1_000_000, 4),
           (480, 640, 10_000, 4), (480, 640, 100_000, 4), (480, 640, 300_000, 4),
           (1024, 1280, 10_000, 4), (1024, 1280, 50_000, 4), (1024, 1280, 5_000, 1)]
out = []
for (H, W, N, sbin) in CONFIGS:
    try:
        # a short synthetic clip tiled in time with a per-repeat brightness offset keeps generation cheap
        base_n = min(N, 5000)
        s = SynthVideo(H, W, base_n, seed=0)
        v = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        step = max(50, min(1000, (1 << 27) // (H * W)))
        clip = torch.empty((base_n, H, W), dtype=torch.uint8, device=dev)
        for t in range(0, base_n, step):
            clip[t:t + step] = s.frames_torch(t, min(base_n, t + step), dev)
        for r, t in enumerate(range(0, N, base_n)):
            n = min(base_n, N - t)
            v[t:t + n] = torch.clamp(clip[:n].to(torch.int16) + (r % 7), 0, 255).to(torch.uint8)
        del s._dev, clip
        torch.cuda.empty_cache()
        src = DeviceArraySource([v])
        for _ in range(2):
            process.process_source(src, sbin=sbin, motSVD=True, movSVD=False, keep_on_device=True)
        torch.cuda.synchronize()
        ts = []
        tm = {}
        for _ in range(3):
            t0 = time.perf_counter()
            tm = {}
            process.process_source(src, sbin=sbin, motSVD=True, movSVD=False, keep_on_device=True, timing=tm)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        eig = sum(x for k, x in tm["device_ms"].items() if k.endswith("syevd"))
        row = {"H": H, "W": W, "N": N, "sbin": sbin, "npix": (H // sbin) * (W // sbin), "seconds": round(dt, 4),
               "frames_per_s": round(N / dt), "cusolver_ms": round(eig, 1),
               "proj_ms": round(tm["device_ms"].get("stage3.project.mma", 0.0), 2),
               "k1_ms": round(tm["device_ms"].get("stage3.bin", 0.0), 2)}
    except Exception as e:
        row = {"H": H, "W": W, "N": N, "sbin": sbin, "error": str(e)[:200]}
    print(json.dumps(row), flush=True)
    out.append(row)
    del v
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
