"""File ingest measurement: `process.run` on a video FILE, decode included.

Writes a synthetic 640x480 gray MJPG video of N frames to a scratch folder, then times
  (a) decoding alone on one thread (what the reference's get_frames does, utils.py:522-553),
  (b) process.run through the on-demand path (every stage decodes the ranges it needs, one thread per view),
  (c) process.run through the decode-once HBM cache (threaded decoders).

    python scripts/bench_ingest.py [--frames 10000] [--workers 8]   -> one JSON line
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--cache-only", action="store_true", help="skip the one-thread decode and the on-demand path")
    a = ap.parse_args()
    import cv2
    import torch

    from facemap_b200 import process
    from facemap_b200.frames import VideoFileSource

    H, W, N = 480, 640, a.frames
    d = tempfile.mkdtemp(prefix="ingest_")
    path = os.path.join(d, "cam.avi")
    t = time.perf_counter()
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (W, H), isColor=False)
    # cheap content: a smooth texture drifting under a fixed vignette plus a little noise (the generator must
    # not dominate the box time; the codec work per frame is what matters here)
    rs = np.random.RandomState(0)
    tex = cv2.GaussianBlur(rs.randint(0, 256, size=(H + 128, W + 128)).astype(np.float32), (0, 0), 6.0)
    tex = ((tex - tex.min()) / (tex.max() - tex.min()) * 200 + 20).astype(np.uint8)
    pos = np.cumsum(rs.randint(-2, 3, size=(N, 2)), axis=0) % 128
    for i in range(N):
        f = tex[pos[i, 0]:pos[i, 0] + H, pos[i, 1]:pos[i, 1] + W].copy()
        f[::7, ::5] += rs.randint(0, 8, size=f[::7, ::5].shape).astype(np.uint8)
        vw.write(f)
    vw.release()
    out = {"frames": N, "H": H, "W": W, "codec": "MJPG", "file_MB": os.path.getsize(path) / 1e6,
           "write_s": time.perf_counter() - t, "host_cores": os.cpu_count()}
    # (a) one-thread decode of the whole file
    if not a.cache_only:
        cap = cv2.VideoCapture(path)
        t = time.perf_counter()
        n = 0
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            if fr.ndim == 3:
                fr = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY)
            n += 1
        dt = time.perf_counter() - t
        cap.release()
        out["decode_one_thread_fps"] = n / dt
    torch.zeros(1, device="cuda")
    for label, cached in ((("decode_once_cache", True),) if a.cache_only else (("on_demand", False), ("decode_once_cache", True))):
        best = None
        for rep in range(2 if cached else 1):
            src = VideoFileSource([[path]], cache_on_device=cached, decode_workers=a.workers)
            try:
                torch.cuda.synchronize()
                t = time.perf_counter()
                process.process_source(src, sbin=4, motSVD=True, movSVD=False)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t
            finally:
                workers = src.decode_workers
                src.close()
            best = dt if best is None else min(best, dt)
        out[label] = {"seconds": best, "frames_per_s": N / best, "decode_workers": workers if cached else 1}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
