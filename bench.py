#!/usr/bin/env python
"""Benchmark of the motion-SVD path (BASELINE.json metric: motion-SVD frames/sec, bin + diff +
cov + SVD + proj) on the configuration the metric is quoted on:

    1 camera, 640x480 uint8, 100 000 frames, sbin = 4, motSVD, 500 components   (configs[1])

    python bench.py --gpus N --steps K --warmup W            # this library (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A step = one pass of the whole path (stage 1 means, stage 2 two-level SVD, stage 3 projection of
every frame) over the synthetic video.  `value` = frames processed by all ranks / device time with
the frames resident in HBM; `e2e` = the same through the public API with the frames in pinned HOST
memory (H2D of every frame read, D2H of every output inside the timed region).
N > 1: weak scaling — every rank owns 100 000 frames of one N x 100 000-frame video; the SVD chunks
are dealt across ranks and all-gathered (facemap_b200/parallel.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, SBIN, NCOMP = 480, 640, 4, 500
FRAMES_PER_GPU = 100_000
WORKLOAD = "1-camera 640x480 uint8, 100k frames per GPU, sbin=4, motSVD 500 components (BASELINE configs[1])"


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region.  NVML from a thread
    (pynvml; its calls release the GIL and cost microseconds) — a looping `nvidia-smi` subprocess stalls
    the CUDA driver for 100-200 ms at a time on this platform, which showed up as outlier steps."""

    def __init__(self, gpu_index, period_s=0.4):
        self.idx, self.period = gpu_index, period_s
        self.samples, self.stop_flag, self.thread, self.nvml = [], False, None, None
        self.call_ms = (0.0, 0.0)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                t0 = time.perf_counter()
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                t1 = time.perf_counter()
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                t2 = time.perf_counter()
                self.samples.append((t2, sm, rs))
                self.call_ms = (max(self.call_ms[0], 1e3 * (t1 - t0)), max(self.call_ms[1], 1e3 * (t2 - t1)))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.thread.join(1.0)
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(s[1] for s in self.samples if t0 <= s[0] <= t1)
        reasons = set()
        for ts, _, rs in self.samples:
            if t0 <= ts <= t1:
                for k, bit in names.items():
                    if rs & bit:
                        reasons.add(k)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.smax),
                "reasons": sorted(reasons), "samples": len(sm), "how": f"NVML every {self.period} s, rank-0 GPU",
                "nvml_call_ms_max": {"clock": round(self.call_ms[0], 2), "reasons": round(self.call_ms[1], 2)}}


# ---------------------------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------------------------
def make_sources(rank, world, device, want_host):
    """Synthetic video of world x FRAMES_PER_GPU frames; this rank keeps frames
    [rank*F, (rank+1)*F) resident (HBM, and pinned host for e2e) and generates any other range it is
    asked for (the SVD sample chunks of stage 1/2) on first use — the warm-up steps fill that cache."""
    import torch

    from facemap_b200.frames import DeviceFrameCache, FrameSource, _H2DPipe
    from facemap_b200.synth import SynthVideo

    total = world * FRAMES_PER_GPU
    synth = SynthVideo(H, W, total, seed=0)
    lo, hi = rank * FRAMES_PER_GPU, (rank + 1) * FRAMES_PER_GPU
    resident = torch.empty((hi - lo, H, W), dtype=torch.uint8, device=device)
    for t in range(lo, hi, 500):
        resident[t - lo:min(hi, t + 500) - lo] = synth.frames_torch(t, min(hi, t + 500), device)
    torch.cuda.synchronize()
    if hasattr(synth, "_dev"):
        del synth._dev
    torch.cuda.empty_cache()

    class ShardSource(FrameSource):
        def __init__(self, on_host):
            self.nframes, self.Ly, self.Lx, self.block_frames = total, [H], [W], [total]
            self.on_host = on_host
            self.cache = {}
            self.h2d = 0
            self.pipe = _H2DPipe()
            self.dcache = DeviceFrameCache()
            self.base = resident
            if on_host:
                self.base = torch.empty(tuple(resident.shape), dtype=torch.uint8, pin_memory=True)
                self.base.copy_(resident)

        def _range(self, a, b):
            if a >= lo and b <= hi:
                return self.base[a - lo:b - lo]
            key = (a, b)
            if key not in self.cache:                    # warm-up only
                fr = synth.frames_torch(a, b, device)
                if hasattr(synth, "_dev"):
                    del synth._dev
                self.cache[key] = fr.cpu().pin_memory() if self.on_host else fr
            return self.cache[key]

        def plan_access(self, priority, home, dev):
            if self.on_host:
                self.dcache.drop()
                a, b = max(home[0], lo), min(home[1], hi)       # the part of `home` this rank holds in host memory
                for (x, y) in priority:                        # sample ranges outside the shard go first: the H2D
                    x, y = self.clamp(x, y)                    # engine serves copies in submission order
                    if not (x >= a and y <= b):
                        self.pipe.issue((x, y), [self._range(x, y)], dev)
                before = self.dcache.bytes
                self.dcache.start(lambda v, x, y: self.base[x - lo:y - lo], [(H, W)], priority, (a, b), dev)
                self.h2d += self.dcache.bytes - before

        def resident(self, t0, t1):
            if self.on_host:
                return self.dcache.covers(t0, t1)
            return t0 >= lo and t1 <= hi

        def prefetch(self, t0, t1, dev):
            if self.on_host:
                a, b = self.clamp(t0, t1)
                if not self.dcache.covers(a, b):
                    self.pipe.issue((a, b), [self._range(a, b)], dev)

        def fetch(self, t0, t1, dev):
            a, b = self.clamp(t0, t1)
            if not self.on_host:
                return [self._range(a, b)]
            if self.dcache.covers(a, b):
                return self.dcache.get(a, b)
            if (a, b) not in self.pipe.pending:
                self.pipe.issue((a, b), [self._range(a, b)], dev)
            self.h2d += (b - a) * H * W
            return self.pipe.take((a, b))

        def h2d_bytes(self, nframes):
            return int(nframes) * H * W if self.on_host else 0

    # the pinned host copy is made lazily, AFTER the device-resident measurement: page-locking and filling
    # 30.7 GB per rank keeps the host busy (and jittery) for a while
    return ShardSource(False), ((lambda: ShardSource(True)) if want_host else None)


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port; /root/reference is absent on the GPU box)
# ---------------------------------------------------------------------------------------------
def cpu_reference_pass(nframes, seed=0):
    """One pass of the reference algorithm (stock sklearn randomized svdecon) over an in-memory
    synthetic video of `nframes` 640x480 frames.  Returns seconds."""
    from facemap_b200.synth import SynthVideo
    from oracle import facemap_oracle as orc

    v = SynthVideo(H, W, nframes, seed=seed).all_frames()
    t0 = time.perf_counter()
    orc.run_oracle([v], sbin=SBIN, motSVD=True, movSVD=False, svd="stock")
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_steps = max(1, args.steps + args.warmup)
    budget = 200.0 / total_steps                       # seconds per step so the whole run stays a few minutes
    # ~7 ms per frame on the host; >= 2000 frames keeps the two-level SVD (2 chunks + merge) in the sample,
    # shorter budgets fall back to a single-chunk sample
    nframes = int(min(4000, budget / 0.007))
    nframes = nframes if nframes >= 2000 else max(1100, min(1900, nframes))
    for _ in range(args.warmup):
        cpu_reference_pass(nframes)
    times = [cpu_reference_pass(nframes) for _ in range(args.steps)]
    dt = sum(times)
    value = nframes * args.steps / dt
    cores = os.cpu_count()
    sample = (f"{nframes} frames of the same 640x480 generator per step, whole path (means, two-level randomized SVD "
              f"with sklearn PCA like facemap/utils.py:773-775, projection), frames in host memory (no decode)")
    line = {
        "impl": "reference", "metric": "motion-SVD frames/sec (bin+diff+cov+SVD+proj)", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_frames_per_step": nframes},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=2000)
    ap.add_argument("--with-rois", action="store_true",
                    help="side measurement (not the headline workload): add a pupil, a blink and a running ROI")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    os.environ["NCCL_DEBUG"] = os.environ.get("FM_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist

    from facemap_b200 import _lib, kernels, process

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _lib.require_device(local)
    shard = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from facemap_b200.parallel import Shard
        shard = Shard()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    want_host = not args.no_e2e
    e2e_skip = None
    if want_host:
        # every rank pins its 30.7 GB shard: only if the host really has the room (never risk the box)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 0
        need = FRAMES_PER_GPU * H * W * 1.25 * world
        if avail < need:
            want_host = False
            e2e_skip = f"host memory: {avail / 2**30:.0f} GiB available < {need / 2**30:.0f} GiB needed to pin the frames of {world} ranks"
    dev_src, host_factory = make_sources(rank, world, device, want_host)
    total_frames = world * FRAMES_PER_GPU

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bench_rois():
        """pupil + blink (100x140 ellipses) and a running band (120x400) on the one view: SURVEY.md §8f row 4"""
        if not args.with_rois:
            return None
        import numpy as np

        yy, xx = np.meshgrid(np.arange(100), np.arange(140), indexing="ij")
        ell = ((yy - 50) ** 2 / 50 ** 2 + (xx - 70) ** 2 / 70 ** 2) <= 1
        base = {"iROI": 0, "ivid": 0, "color": (0, 0, 0)}
        return [dict(base, rind=0, rtype="Pupil", yrange=np.arange(100, 200), xrange=np.arange(180, 320),
                     saturation=90.27, pupil_sigma=2.0, ellipse=ell),
                dict(base, rind=2, rtype="blink", yrange=np.arange(60, 160), xrange=np.arange(400, 540),
                     saturation=140.25, ellipse=ell),
                dict(base, rind=3, rtype="running", yrange=np.arange(340, 460), xrange=np.arange(100, 500), saturation=0)]

    def one_step(src, timing=None, keep=True, overlap=True):
        # every rank keeps the projections of its own frame range (north_star: "each rank projects its own
        # frames with no further communication"); the gather for the file is process.run's business
        return process.process_source(src, sbin=SBIN, motSVD=True, movSVD=False, dist=shard, timing=timing,
                                      keep_on_device=keep, overlap_stage3=overlap, gather_outputs=False,
                                      rois=bench_rois())

    nwarm = max(args.warmup, 5)      # allocator pools, cuSOLVER workspaces and NCCL channels settle in ~4 passes
    clocks = ClockSampler(local)
    for i in range(nwarm):
        if i == nwarm - 1 and rank == 0:
            clocks.start()                       # one sampler (rank 0's GPU), spawned outside the timed region
        one_step(dev_src, timing={})             # same code path as the timed steps (CUDA-event timers on)
    barrier()
    timings = []
    import gc
    gc.collect()
    gc.disable()                                 # no collector pauses inside the timed region
    n0 = kernels.launch_count()
    mem0 = torch.cuda.memory_stats(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        tm = {}
        ts = time.perf_counter()
        one_step(dev_src, timing=tm)
        tm["host_ms"] = 1e3 * (time.perf_counter() - ts)
        timings.append(tm)
    ev1.record()
    barrier()
    w1 = time.perf_counter()
    gc.enable()
    launches = kernels.launch_count() - n0
    mem1 = torch.cuda.memory_stats(device)
    alloc_diag = {k: int(mem1.get(k, 0) - mem0.get(k, 0)) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    clk = clocks.stop(w0, w1) if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = total_frames / (ms_step / 1e3)

    # per-kernel device time (ms per step, this rank)
    agg = {}
    for tm in timings:
        for k, v in tm["device_ms"].items():
            agg[k] = agg.get(k, 0.0) + v / len(timings)
    # K1 runs on a side stream underneath the eigensolves in the timed steps; its own duration is taken from
    # two extra passes with that overlap switched off (same kernels, same shapes, CUDA events)
    k1_clean = []
    for _ in range(2):
        tm = {}
        one_step(dev_src, timing=tm, overlap=False)
        k1_clean.append(tm["device_ms"].get("stage3.bin", 0.0))
    barrier()
    p = (H // SBIN) * (W // SBIN)
    f0, f1 = (0, total_frames) if shard is None else __import__("facemap_b200.parallel", fromlist=["x"]).frame_ranges(
        total_frames, world)[rank]
    rows3 = f1 - f0
    k1_ms = min(k1_clean) if k1_clean else agg.get("stage3.bin", 0.0)
    k1_bytes = rows3 * (H * W + 4 * p)                       # algorithmic: read u8 frame once + write fp32 X
    k1 = {"kernel": "bin_motion_vec_kernel<4> (stage 3)", "bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms else None,
          "peak": hbm_peak, "unit": "GB/s", "ms_per_step": k1_ms, "algorithmic_bytes_per_frame": H * W + 4 * p,
          "ms_per_step_overlapped_with_stage2": agg.get("stage3.bin", 0.0)}
    if k1["achieved"]:
        k1["frac"] = k1["achieved"] / hbm_peak
    mma_ms = agg.get("stage3.project.mma", 0.0)
    flops = 2.0 * rows3 * p * NCOMP                          # algorithmic (a 3xTF32 split issues 3x this)
    mm = {"kernel": "gemm_tf32x3_chain_kernel<16,1> (stage 3 projection)", "bound": "tensor",
          "achieved": flops / (mma_ms * 1e-3) / 1e12 if mma_ms else None, "peak": tc_peak, "unit": "TFLOP/s",
          "ms_per_step": mma_ms, "algorithmic_flop_per_frame": 2.0 * p * NCOMP,
          "note": "peak = measured dense bf16; the kernel runs kind::tf32 (half the bf16 rate) and issues 3 MMAs per "
                  "algorithmic product, so frac <= 1/6 by construction; issued TF32 rate = 3 x achieved"}
    if mm["achieved"]:
        mm["frac"] = mm["achieved"] / tc_peak
        mm["issued_tf32_tflops"] = 3 * mm["achieved"]
    proj_i8 = os.environ.get("FM_PROJ_MODE") == "i8"
    if proj_i8:     # experimental integer projection: byte planes out of K1, kind::i8 products (6 per 32 K at sbin 4)
        k1.update({"kernel": "bin_motion_vec_kernel<4, planes> (stage 3)", "algorithmic_bytes_per_frame": H * W + 2 * p})
        if k1_ms:
            k1["achieved"] = rows3 * (H * W + 2 * p) / (k1_ms * 1e-3) / 1e9
            k1["frac"] = k1["achieved"] / hbm_peak
        mm.update({"kernel": "project_i8_kernel<2> (stage 3 projection)",
                   "note": "peak = measured dense bf16; kind::i8 runs at twice the bf16 rate and 6 plane products form "
                           "one algorithmic product (2 frame planes x 3 mask digits), so frac <= 1/3 by construction"})
        mm.pop("issued_tf32_tflops", None)
    dominant, other = (mm, k1) if mma_ms >= k1_ms else (k1, mm)
    traffic = None
    try:    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["gemm_tf32x3_projection" if dominant is mm else "bin_motion_vec_kernel"]["bytes_per_launch"]
        traffic = None if proj_i8 else traffic               # no ncu capture of the experimental kernels yet
    except Exception:
        pass
    mm["algorithmic_bytes_per_launch"] = 4 * (18944 * p + NCOMP * p + 18944 * NCOMP)
    roofline = {"bound": dominant["bound"], "achieved": dominant["achieved"], "peak": dominant["peak"],
                "unit": dominant["unit"], "frac": dominant.get("frac"), "traffic": traffic, "kernel": dominant["kernel"],
                "peak_source": peak_src, "detail": dominant, "other_kernel": other}
    solver_ms = sum(v for k, v in agg.items() if k.endswith(".syevd"))

    # ---- e2e: host-resident frames through the public API ----
    e2e = None
    if want_host:
        host_src = host_factory()
        for _ in range(2):                                   # warm-up (host cache of sample chunks, pinned pools)
            one_step(host_src, timing={}, keep=False)
        barrier()
        host_src.h2d = 0
        d2h = 0
        t0 = time.perf_counter()
        e2e_tm = {}
        for _ in range(args.e2e_steps):
            e2e_tm = {"timeline": True}
            out = one_step(host_src, timing=e2e_tm, keep=False)
            d2h = sum(a.nbytes for key in ("avgframe", "avgmotion", "motion", "motMask", "motSVD") for a in out[key]) \
                + out["motSv"].nbytes
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        io = torch.tensor([host_src.h2d // args.e2e_steps, int(d2h)], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(io, op=dist.ReduceOp.SUM)          # bytes of the whole job, all ranks
        e2e = {"value": total_frames * args.e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
               "steps": args.e2e_steps, "api": "facemap_b200.process.process_source(HostArraySource-like pinned frames)",
               "upload_ms": round(host_src.dcache.upload_ms(), 1), "wall_s_last_step": round(e2e_tm.get("wall_s", 0.0), 4),
               "assemble_wall_s": round(e2e_tm.get("assemble_wall_s", 0.0), 4),
               "stage_ms": {k: round(v, 2) for k, v in sorted(e2e_tm.get("device_ms", {}).items()) if v > 1.0},
               "timeline_ms": {k: round(v, 1) for k, v in sorted(e2e_tm.get("timeline_ms", {}).items(), key=lambda kv: kv[1])
                               if k.split(".")[-1] in ("fetch", "syevd", "mma", "bin", "d2h")}}

    # ---- CPU baseline (rank 0, single GPU run only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nfr = args.cpu_frames
        secs = cpu_reference_pass(nfr)
        cpu = {"value": nfr / secs, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{nfr} frames of the same 640x480 generator, whole path with the reference's stock randomized "
                         f"svdecon (oracle/facemap_oracle.py), frames in host memory (no decode); {secs:.1f} s"}

    diag = None
    if os.environ.get("FM_BENCH_DIAG"):
        tm = {"sync": True}
        one_step(dev_src, timing=tm)
        diag = {"rank": rank, "wall_ms": tm.get("wall_ms")}
        print("DIAG " + json.dumps(diag), file=sys.stderr, flush=True)
    if rank == 0:
        line = {
            "metric": "motion-SVD frames/sec (bin+diff+cov+SVD+proj)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": nwarm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "dtype_detail": "u8 frames, integer block sums / energies, f32 operands as 3xTF32 tensor-core products with f64 chain reduction, f64 eigensolve",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_total": total_frames, "frame_bytes": H * W,
                       "l2": "inputs (30.7 GB per GPU) exceed L2; no flush needed", "parallelism": f"frame-range x{world}",
                       **({"gram_mode": os.environ["FM_GRAM_MODE"]} if os.environ.get("FM_GRAM_MODE") else {}),
                       **({"proj_mode": os.environ["FM_PROJ_MODE"]} if os.environ.get("FM_PROJ_MODE") else {}),
                       **({"proj_i8_impl": os.environ["FM_PROJ_I8_IMPL"]} if os.environ.get("FM_PROJ_I8_IMPL") else {}),
                       **({"rois": "pupil 100x140 + blink 100x140 + running 120x400 (side measurement)"} if args.with_rois else {})},
            "clocks": clk, "e2e": e2e if e2e is not None else ({"value": None, "skipped": e2e_skip} if e2e_skip else None),
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "stage_ms": {k: round(v, 3) for k, v in sorted(agg.items())},
            "cusolver_ms_per_step": solver_ms,
            "host_ms_per_step": [round(tm["host_ms"], 1) for tm in timings],
            "allocator_in_timed_region": alloc_diag,
            "report_ms_each_step": [(tm.get("final_sync_ms"), tm.get("report_ms")) for tm in timings],
            "host_marks_ms_slowest_step": max(timings, key=lambda tm: tm["host_ms"]).get("host_marks_ms"),
            "host_marks_ms_fastest_step": min(timings, key=lambda tm: tm["host_ms"]).get("host_marks_ms"),
            "cusolver_ms_each_step": [round(sum(v for k, v in tm["device_ms"].items() if k.endswith(".syevd")), 1)
                                      for tm in timings],
            "kernels_only_frames_per_s": total_frames / ((ms_step - solver_ms) / 1e3) if ms_step > solver_ms else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
