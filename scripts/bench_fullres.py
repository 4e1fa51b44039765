"""Kernel times of the full-resolution ROI processors at the BASELINE frame size (640x480), one stage-3
fetch (2368 frames) per launch, CUDA events on the launching stream, inputs larger than L2 (727 MB).

    python scripts/bench_fullres.py [--frames 2368] [--reps 5]  -> one JSON line
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facemap_b200 import fullres, kernels as K  # noqa: E402
from tests.helpers import ellipse_mask  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2368)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    T, H, W = a.frames, 480, 640
    g = torch.Generator(device=dev).manual_seed(0)
    fr = torch.randint(0, 256, (T, H, W), dtype=torch.uint8, device=dev, generator=g)
    # an eye-like dark disc so the pupil fit has a realistic number of dark pixels
    yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    disc = ((yy - 150) ** 2 + (xx - 250) ** 2) <= 30 ** 2
    fr[:, disc] = (fr[:, disc] // 8)
    fr[:, ~disc] = 128 + fr[:, ~disc] // 2
    out = {"frames": T, "H": H, "W": W, "bytes_per_launch": T * H * W}
    # paint
    om = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    om[100:200, 180:320] = torch.from_numpy(np.where(ellipse_mask(100, 140), 0, 255).astype(np.uint8)).to(dev)
    buf = torch.empty_like(fr)
    ms = timed(lambda: K.paint(fr, om, buf), a.reps)
    out["paint"] = {"ms": ms, "GBps_read_plus_write": 2 * T * H * W / ms / 1e6}
    # blink 100x140
    rect = (100, 200, 180, 320)
    mask = torch.from_numpy((~ellipse_mask(100, 140)).astype(np.uint8)).to(dev)
    term = torch.from_numpy(fullres.blink_terms(140.25)).to(dev)
    o1 = torch.zeros(T, dtype=torch.float64, device=dev)
    ms = timed(lambda: K.blink(fr, rect, mask, term, o1), a.reps)
    out["blink_100x140"] = {"ms": ms, "us_per_frame": 1e3 * ms / T, "roi_GBps": T * 100 * 140 / ms / 1e6}
    # pupil 100x140 (disc of radius 30 -> ~2800 dark pixels)
    o9 = torch.zeros((T, 9), dtype=torch.float64, device=dev)
    out["pupil_100x140"] = {}
    for blocks in (296, 592, 1184):
        scratch = K.pupil_scratch(rect, blocks, dev)
        ms = timed(lambda: K.pupil(fr, rect, mask, np.float32(255.0 - 90.27), 2.0, scratch, blocks, o9), a.reps)
        out["pupil_100x140"][f"blocks_{blocks}"] = {"ms": ms, "us_per_frame": 1e3 * ms / T,
                                                    "nan_frames": int(torch.isnan(o9[:, 2]).sum())}
    # running 120x400
    rrect = (340, 460, 100, 500)
    ormask = torch.zeros((120, 400), dtype=torch.uint8, device=dev)
    g2 = torch.from_numpy(fullres.gaussian_fourier_sq(120, 400)).to(dev)
    means = torch.empty(T, dtype=torch.float32, device=dev)
    o2 = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    ms = timed(lambda: K.running(fr, rrect, ormask, g2, means, o2), a.reps)
    out["running_120x400"] = {"ms": ms, "us_per_frame": 1e3 * ms / T, "roi_GBps": T * 120 * 400 / ms / 1e6}
    # the reference-arithmetic motion / mean kernels of the non-power-of-two bins (sbin 3 and 7) and K1 beside them
    from facemap_b200 import svd
    for sbin in (3, 7):
        npix = (H // sbin) * (W // sbin)
        plan = K.pairwise_plan(npix, dev)
        scr = K.motion_ref_scratch(H, W, sbin, plan[0].numel(), svd.MOTION_REF_BLOCKS, dev)
        o64 = torch.zeros(T, dtype=torch.float64, device=dev)
        ms = timed(lambda: K.motion_ref(fr, sbin, plan, scr, svd.MOTION_REF_BLOCKS, o64), a.reps)
        X = K.alloc_matrix(T - 1, npix, dev)
        avg = torch.zeros(npix, dtype=torch.float32, device=dev)
        e = torch.zeros(T - 1, dtype=torch.int64, device=dev)
        ms_k1 = timed(lambda: K.bin_motion(fr, sbin, avgmotion=avg, x_mot=X, energy=e), a.reps)
        mb, ma = torch.zeros(npix, dtype=torch.float64, device=dev), torch.zeros(npix, dtype=torch.float64, device=dev)
        ms_mean = timed(lambda: K.mean_ref(fr[:100], sbin, mb, ma), a.reps)
        out[f"sbin{sbin}"] = {"motion_ref_ms": ms, "motion_ref_GBps": T * H * W / ms / 1e6, "k1_generic_ms": ms_k1,
                              "k1_generic_GBps": T * H * W / ms_k1 / 1e6, "mean_ref_100_frames_ms": ms_mean}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
