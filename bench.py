#!/usr/bin/env python
"""Benchmark of the motion-SVD path (BASELINE.json metric: motion-SVD frames/sec, bin + diff +
cov + SVD + proj) on the configuration the metric is quoted on:

    1 camera, 640x480 uint8, 100 000 frames, sbin = 4, motSVD, 500 components   (configs[1])

    python bench.py --gpus N --steps K --warmup W            # this library (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A step = one pass of the whole path (stage 1 means, stage 2 two-level SVD, stage 3 projection of
every frame) over the synthetic video.  `value` = frames processed by all ranks / device time with
the frames resident in HBM; `e2e` = the same through the public API with the frames in pinned HOST
memory (H2D of every frame read, D2H of every output inside the timed region; at N = 1 through
`process.run` on a `HostArraySource`, the `_proc.npy` file write included).
N > 1: weak scaling — every rank owns 100 000 frames of one N x 100 000-frame video; the SVD chunks
are dealt across ranks and all-gathered (facemap_b200/parallel.py).

Side measurements on the same JSON line (`extras`, none of them touches the headline fields):
  strong_scaling_100k   ONE 100 000-frame video cut over the N ranks (what north_star's target names)
  configs2              BASELINE configs[2]: two 640x480 cameras, motion + movie SVD, three motion ROIs
  configs3              BASELINE configs[3]: 1280x1024 at sbin 1 (1.31 M binned pixels), frames made on the device

The reference arm (`--impl reference`) costs the reference's algorithm on the SAME 100 000-frame
configuration from its four parts, each timed on the host cores on a sample
(facemap/process.py:53-102, 134-141, 191-259, 264-295, 394-515):

    T(100k) = T_mean(1000 frames) + 15 x T_chunk(1000 frames: bin + svdecon k=250)
              + T_merge(svdecon of 19200 x 3750, k=500) + 100000 x t_proj(bin + diff + sgemm per frame)
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, SBIN, NCOMP = 480, 640, 4, 500
FRAMES_PER_GPU = 100_000
METRIC = "motion-SVD frames/sec (bin+diff+cov+SVD+proj)"
WORKLOAD = "1-camera 640x480 uint8, 100k frames per GPU, sbin=4, motSVD 500 components (BASELINE configs[1])"


def workload_config(world):
    """`config` of the JSON line — the same dict for this library's arm and the reference arm."""
    cfg = {"workload": WORKLOAD, "frames_total": world * FRAMES_PER_GPU, "frame_bytes": H * W,
           "l2": "inputs (30.7 GB per GPU) exceed L2; no flush needed", "parallelism": f"frame-range x{world}"}
    for env, key in (("FM_GRAM_MODE", "gram_mode"), ("FM_PROJ_MODE", "proj_mode"), ("FM_PROJ_I8_IMPL", "proj_i8_impl")):
        if os.environ.get(env):
            cfg[key] = os.environ[env]
    return cfg


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region, through NVML (pynvml), from the main
    thread at step boundaries (every other step, ~1 ms per sample, inside the timed region).  A sampling THREAD is not
    an option on this path: every NVML query contends for the driver lock that cuSOLVER's host-driven eigensolve
    takes for each of its ~10^4 launches — queries were seen to wait 46-149 ms and the step they landed in ran 30-60 ms
    longer (profiles/r02_summary.md); a looping `nvidia-smi` subprocess is worse still (100-200 ms driver stalls)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples, self.nvml = [], None
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

    def sample(self):
        nv = self.nvml
        if nv is None:
            return
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

    def stop(self, t0, t1):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
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
                "reasons": sorted(reasons), "samples": len(sm),
                "how": "NVML from the main thread between timed steps (after 1/4, 1/2, 3/4 of them), rank-0 GPU",
                "nvml_call_ms_max": {"clock": round(self.call_ms[0], 2), "reasons": round(self.call_ms[1], 2)}}


# ---------------------------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------------------------
def make_shard_source_class():
    import torch

    from facemap_b200.frames import DeviceFrameCache, FrameSource, _H2DPipe

    class ShardSource(FrameSource):
        """One rank's window on a synthetic multi-view video of `total` frames: frames [lo, hi) are held
        (`held[v]`: HBM tensors, or pinned host tensors with on_host), any other range it is asked for (the
        SVD sample chunks of stage 1 / 2 that belong to other ranks' ranges) is generated on first use by
        `gen(view, a, b)` and cached — the warm-up steps fill that cache."""

        def __init__(self, held, total, lo, hi, gen, device, on_host=False):
            self.held = held
            self.nframes, self.block_frames = int(total), [int(total)]
            self.Ly = [int(t.shape[1]) for t in held]
            self.Lx = [int(t.shape[2]) for t in held]
            self.lo, self.hi, self.gen, self.device, self.on_host = int(lo), int(hi), gen, device, on_host
            self.cache = {}
            self.h2d = 0
            self.pipe = _H2DPipe()
            self.dcache = DeviceFrameCache()

        def _range(self, a, b):
            if a >= self.lo and b <= self.hi:
                return [t[a - self.lo:b - self.lo] for t in self.held]
            key = (a, b)
            if key not in self.cache:                    # warm-up only
                fr = [self.gen(v, a, b) for v in range(len(self.held))]
                self.cache[key] = [f.cpu().pin_memory() for f in fr] if self.on_host else fr
            return self.cache[key]

        def plan_access(self, priority, home, dev):
            if self.on_host:
                self.dcache.drop()
                a, b = max(home[0], self.lo), min(home[1], self.hi)   # the part of `home` this rank holds in host memory
                for (x, y) in priority:                        # sample ranges outside the shard go first: the H2D
                    x, y = self.clamp(x, y)                    # engine serves copies in submission order
                    if not (x >= a and y <= b):
                        self.pipe.issue((x, y), self._range(x, y), dev)
                before = self.dcache.bytes
                self.dcache.start(lambda v, x, y: self.held[v][x - self.lo:y - self.lo], list(zip(self.Ly, self.Lx)),
                                  priority, (a, b), dev)
                self.h2d += self.dcache.bytes - before

        def resident(self, t0, t1):
            if self.on_host:
                return self.dcache.covers(t0, t1)
            return t0 >= self.lo and t1 <= self.hi

        def prefetch(self, t0, t1, dev):
            if self.on_host:
                a, b = self.clamp(t0, t1)
                if not self.dcache.covers(a, b):
                    self.pipe.issue((a, b), self._range(a, b), dev)

        def fetch(self, t0, t1, dev):
            a, b = self.clamp(t0, t1)
            if not self.on_host:
                return self._range(a, b)
            if self.dcache.covers(a, b):
                return self.dcache.get(a, b)
            if (a, b) not in self.pipe.pending:
                self.pipe.issue((a, b), self._range(a, b), dev)
            self.h2d += (b - a) * sum(h * w for h, w in zip(self.Ly, self.Lx))
            return self.pipe.take((a, b))

        def h2d_bytes(self, nframes):
            return int(nframes) * sum(h * w for h, w in zip(self.Ly, self.Lx)) if self.on_host else 0

    return ShardSource


def generate_resident(synths, lo, hi, device, piece=500):
    """frames [lo, hi) of every view, made on the device (same bytes as the NumPy generator)"""
    import torch

    out = []
    for s in synths:
        piece_v = max(16, int(piece * (480 * 640) / (s.H * s.W)))
        t = torch.empty((hi - lo, s.H, s.W), dtype=torch.uint8, device=device)
        for a in range(lo, hi, piece_v):
            b = min(hi, a + piece_v)
            t[a - lo:b - lo] = s.frames_torch(a, b, device)
        out.append(t)
        if hasattr(s, "_dev"):
            del s._dev
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


def synth_gen(synths, device, index_map=None):
    """gen(view, a, b) for ShardSource: frames [a, b) of the video, through `index_map` (a list of
    (start, stop, source_start) pieces) when the video is a re-indexing of the generator's frames"""
    import torch

    def gen(v, a, b):
        s = synths[v]
        if index_map is None:
            out = s.frames_torch(a, b, device)
        else:
            parts = []
            for (p0, p1, src0) in index_map:
                x, y = max(a, p0), min(b, p1)
                if x < y:
                    parts.append(s.frames_torch(src0 + x - p0, src0 + y - p0, device))
            out = torch.cat(parts, dim=0)
        if hasattr(s, "_dev"):
            del s._dev
        return out

    return gen


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port; /root/reference is absent on the GPU box)
# ---------------------------------------------------------------------------------------------
def cpu_reference_model(rounds=1, warm=0, heavy_rounds=2, proj_frames=500, log=None):
    """Cost of ONE pass of the reference algorithm over the 100 000-frame configuration, from its four parts timed
    on the host cores on samples of the same synthetic video (oracle port, stock sklearn randomized svdecon like
    facemap/utils.py:773-775; frames in host memory, no decode):

        mean    stage 1 on 200 of its 1 000 frames (2 of 10 segments)                     x 5
        chunk   one 1 000-frame SVD chunk: bin + |diff| - avg, svdecon(19 200 x 999, 250) x 15
        merge   svdecon(19 200 x 3 750, k = 500) of 15 stacked chunk blocks               x 1
        proj    stage 3 on `proj_frames` frames: bin + diff + energy + X @ U               x 100 000 / proj_frames

    mean and proj are timed in every round, chunk and merge in the first `heavy_rounds` timed rounds (they are
    deterministic in cost); the parts are averaged over the timed rounds.  Returns (seconds for 100k frames, parts)."""
    import numpy as np

    from facemap_b200.synth import SynthVideo
    from oracle import facemap_oracle as orc

    say = (lambda *a: None) if log is None else log
    t0 = time.perf_counter()
    v = SynthVideo(H, W, 1000, seed=0).all_frames()
    say(f"sample video: 1000 frames generated in {time.perf_counter() - t0:.1f} s")
    Lyb, Lxb, cols = orc.binned_geometry([H], [W], SBIN)
    parts = {"mean": [], "chunk_bin": [], "chunk_svd": [], "merge": [], "proj": []}
    U = Ycat = None
    af = am = None
    for r in range(warm + rounds):
        timed = r >= warm
        t = time.perf_counter()
        af, am = orc.stage1_means(orc.ArrayVideo([v[:200]]), SBIN)
        dt_mean = time.perf_counter() - t
        heavy = (r - warm) < heavy_rounds if timed else (U is None)
        if heavy:
            t = time.perf_counter()
            X, _, _, _ = orc.chunk_matrices([v], SBIN, Lyb, Lxb, cols, af, am, True, False)
            dt_bin = time.perf_counter() - t
            t = time.perf_counter()
            u, s, _ = orc.svd_centered(X.T, 250, "stock")
            dt_svd = time.perf_counter() - t
            if Ycat is None:
                # 15 stacked chunk blocks: this chunk's U S, each copy rescaled and perturbed at its noise floor (the
                # randomized solver's cost does not depend on the values: fixed 4 power iterations)
                rs = np.random.RandomState(0)
                Y = (u * s).astype(np.float32)
                Ycat = np.ascontiguousarray(np.concatenate(
                    [Y * rs.uniform(0.8, 1.2, size=(1, 250)).astype(np.float32)
                     + (rs.standard_normal(Y.shape) * (s[-1] / np.sqrt(Y.shape[0]))).astype(np.float32) for _ in range(15)],
                    axis=1))
            t = time.perf_counter()
            U, _, _ = orc.svd_centered(Ycat, NCOMP, "stock")
            dt_merge = time.perf_counter() - t
            if timed:
                parts["chunk_bin"].append(dt_bin)
                parts["chunk_svd"].append(dt_svd)
                parts["merge"].append(dt_merge)
        t = time.perf_counter()
        orc.stage3_project(orc.ArrayVideo([v[:proj_frames]]), SBIN, [a.copy() for a in af], [a.copy() for a in am],
                           [U], [], True, False)
        dt_proj = time.perf_counter() - t
        if timed:
            parts["mean"].append(dt_mean)
            parts["proj"].append(dt_proj)
        say(f"round {r}{'' if timed else ' (warm-up)'}: mean {dt_mean:.2f} s, proj {dt_proj:.2f} s" + (f", chunk {dt_bin:.2f} + {dt_svd:.2f} s, merge {dt_merge:.2f} s" if heavy else ""))
    avg = {k: float(np.mean(x)) for k, x in parts.items()}
    total = 5 * avg["mean"] + 15 * (avg["chunk_bin"] + avg["chunk_svd"]) + avg["merge"] + FRAMES_PER_GPU * avg["proj"] / proj_frames
    detail = {"T_mean_s": 5 * avg["mean"], "T_chunk_s": avg["chunk_bin"] + avg["chunk_svd"], "T_chunk_bin_s": avg["chunk_bin"],
              "T_chunk_svdecon_s": avg["chunk_svd"], "T_merge_s": avg["merge"], "t_proj_ms_per_frame": 1e3 * avg["proj"] / proj_frames,
              "T_100k_s": total, "proj_sample_frames": proj_frames, "rounds": rounds, "heavy_rounds": len(parts["merge"]),
              "formula": "T_mean + 15 T_chunk + T_merge + 100000 t_proj (facemap/process.py:134-141, 394-395)"}
    return total, detail


def cpu_sample_text(detail, cores):
    return (f"reference algorithm (oracle port, stock sklearn randomized svdecon) costed on the 100 000-frame configuration "
            f"from samples of the same 640x480 generator on {cores} host cores, frames in host memory (no decode): "
            f"T_mean {detail['T_mean_s']:.2f} s (200 of 1000 frames x5) + 15 x T_chunk {detail['T_chunk_s']:.2f} s (1000 frames: bin "
            f"{detail['T_chunk_bin_s']:.2f} + svdecon 19200x999 k=250 {detail['T_chunk_svdecon_s']:.2f}) + T_merge "
            f"{detail['T_merge_s']:.2f} s (svdecon 19200x3750 k=500) + 100000 x t_proj {detail['t_proj_ms_per_frame']:.3f} ms "
            f"({detail['proj_sample_frames']}-frame sample) = {detail['T_100k_s']:.1f} s")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    total, detail = cpu_reference_model(rounds=max(1, args.steps), warm=min(args.warmup, 1), heavy_rounds=min(max(1, args.steps), 3),
                                        log=lambda s: print("[reference arm] " + s, file=sys.stderr, flush=True))
    value = FRAMES_PER_GPU / total
    sample = cpu_sample_text(detail, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(1, int(os.environ.get("WORLD_SIZE", "1")))),
        "model": "one CPU process runs the reference algorithm on ONE 100 000-frame video whatever N is (the reference "
                 "has no multi-GPU or multi-process path); a step = the modelled pass over 100 000 frames",
        "reference_model": detail,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# measured peaks next to the driver's bf16 figure
# ---------------------------------------------------------------------------------------------
def measure_matmul_peaks(device):
    """Dense TF32 and int8 tensor-core throughput of library GEMMs in this process (8192^3, best of 8, CUDA events):
    the denominators the 3xTF32 and kind::i8 kernels are really up against.  Plumbing only (torch / cuBLASLt)."""
    import torch

    out = {}
    n = 8192
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def best(fn, reps=8):
        fn()
        torch.cuda.synchronize()
        b = 1e9
        for _ in range(reps):
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            b = min(b, ev0.elapsed_time(ev1))
        return b

    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((n, n), device=device, dtype=torch.float32)
        b = torch.randn((n, n), device=device, dtype=torch.float32)
        out["tf32_tflops"] = 2.0 * n ** 3 / (best(lambda: torch.matmul(a, b)) * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
    except Exception as e:      # noqa: BLE001
        out["tf32_error"] = str(e)[:120]
    try:
        a = torch.randint(-100, 100, (n, n), device=device, dtype=torch.int8)
        b = torch.randint(-100, 100, (n, n), device=device, dtype=torch.int8)
        out["int8_tops"] = 2.0 * n ** 3 / (best(lambda: torch._int_mm(a, b)) * 1e-3) / 1e12
        del a, b
    except Exception as e:      # noqa: BLE001
        out["int8_error"] = str(e)[:120]
    try:
        a = torch.randn((n, n), device=device, dtype=torch.bfloat16)
        b = torch.randn((n, n), device=device, dtype=torch.bfloat16)
        out["bf16_tflops_this_process"] = 2.0 * n ** 3 / (best(lambda: torch.matmul(a, b)) * 1e-3) / 1e12
        del a, b
    except Exception as e:      # noqa: BLE001
        out["bf16_error"] = str(e)[:120]
    out["how"] = "torch.matmul / torch._int_mm 8192^3, best of 8, CUDA events, this process"
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--c2-frames", type=int, default=20_000, help="frames per GPU of the configs[2] side measurement")
    ap.add_argument("--c3-frames", type=int, default=50_000, help="frames per GPU of the configs[3] side measurement")
    ap.add_argument("--with-rois", action="store_true",
                    help="side measurement (not the headline workload): add a pupil, a blink and a running ROI")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from facemap_b200 import _lib, kernels, process
    from facemap_b200.synth import SynthVideo

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
        from facemap_b200.parallel import Shard, frame_ranges
        shard = Shard()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    mm_peaks = measure_matmul_peaks(device) if rank == 0 else {}

    ShardSource = make_shard_source_class()
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
    total_frames = world * FRAMES_PER_GPU
    synth = SynthVideo(H, W, total_frames, seed=0)
    lo, hi = rank * FRAMES_PER_GPU, (rank + 1) * FRAMES_PER_GPU
    resident = generate_resident([synth], lo, hi, device)
    dev_src = ShardSource(resident, total_frames, lo, hi, synth_gen([synth], device), device)

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

    def one_step(src, timing=None, keep=True, overlap=True, **kw):
        # every rank keeps the projections of its own frame range (north_star: "each rank projects its own
        # frames with no further communication"); the gather for the file is process.run's business
        opts = dict(sbin=SBIN, motSVD=True, movSVD=False, dist=shard, timing=timing, keep_on_device=keep,
                    overlap_stage3=overlap, gather_outputs=False, rois=bench_rois())
        opts.update(kw)
        return process.process_source(src, **opts)

    nwarm = max(args.warmup, 5)      # allocator pools, cuSOLVER workspaces and NCCL channels settle in ~4 passes
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                           # NVML handle of rank 0's GPU, opened outside the timed region
    for i in range(nwarm):
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
    torch.cuda.cudart().cudaProfilerStart()      # no-op outside a profiler; `ncu --profile-from-start off` lists the timed region only
    ev0.record()
    # NVML is read three times inside the timed region (after a quarter, half, three quarters of the steps): a query
    # normally takes 2 ms, but now and then the driver answers after 25 - 100 ms and the step that follows runs
    # 20 - 50 ms long (r02_logs/bench_r2t.json), so the sampling is kept sparse
    sample_after = {max(0, args.steps // 4 - 1), max(0, args.steps // 2 - 1), max(0, (3 * args.steps) // 4 - 1)}
    for step in range(args.steps):
        tm = {}
        ts = time.perf_counter()
        one_step(dev_src, timing=tm)
        tm["host_ms"] = 1e3 * (time.perf_counter() - ts)
        timings.append(tm)
        if rank == 0 and step in sample_after:
            clocks.sample()                      # inside the timed region, between two steps
    ev1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
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
    f0, f1 = (0, total_frames) if shard is None else frame_ranges(total_frames, world)[rank]
    rows3 = f1 - f0
    proj_i8 = os.environ.get("FM_PROJ_MODE", "i8") == "i8"
    xbytes = 2 if proj_i8 else 4                              # K1 writes two u8 planes (i8 mode) or one f32 per binned pixel
    k1_ms = min(k1_clean) if k1_clean else agg.get("stage3.bin", 0.0)
    k1_bytes = rows3 * (H * W + xbytes * p)                  # algorithmic: read u8 frame once + write the X operand
    k1 = {"kernel": ("bin_motion_vec_kernel<4, planes> (stage 3)" if proj_i8 else "bin_motion_vec_kernel<4> (stage 3)"),
          "bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms else None,
          "peak": hbm_peak, "unit": "GB/s", "ms_per_step": k1_ms, "algorithmic_bytes_per_frame": H * W + xbytes * p,
          "ms_per_step_overlapped_with_stage2": agg.get("stage3.bin", 0.0)}
    if k1["achieved"]:
        k1["frac"] = k1["achieved"] / hbm_peak
    mma_ms = agg.get("stage3.project.mma", 0.0)
    flops = 2.0 * rows3 * p * NCOMP                          # algorithmic
    if proj_i8:
        mm = {"kernel": "project_i8_kernel<2> (stage 3 projection)", "bound": "tensor",
              "note": "peak = measured dense bf16 (MEASURED_PEAKS.json); the kernel runs kind::i8 (twice the bf16 rate) and "
                      "forms one algorithmic product from 6 plane products (2 frame planes x 3 mask digits), so frac <= 1/3 "
                      "by construction; issued int8 rate = 6 x achieved"}
    else:
        mm = {"kernel": "gemm_tf32x3_chain_kernel<16,1> (stage 3 projection)", "bound": "tensor",
              "note": "peak = measured dense bf16; the kernel runs kind::tf32 (half the bf16 rate) and issues 3 MMAs per "
                      "algorithmic product, so frac <= 1/6 by construction; issued TF32 rate = 3 x achieved"}
    mm.update({"achieved": flops / (mma_ms * 1e-3) / 1e12 if mma_ms else None, "peak": tc_peak, "unit": "TFLOP/s",
               "ms_per_step": mma_ms, "algorithmic_flop_per_frame": 2.0 * p * NCOMP})
    if mm["achieved"]:
        mm["frac"] = mm["achieved"] / tc_peak
        if proj_i8:
            mm["issued_int8_tops"] = 6 * mm["achieved"]
            if mm_peaks.get("int8_tops"):
                mm["frac_of_measured_int8_peak_issued"] = 6 * mm["achieved"] / mm_peaks["int8_tops"]
        else:
            mm["issued_tf32_tflops"] = 3 * mm["achieved"]
            if mm_peaks.get("tf32_tflops"):
                mm["frac_of_measured_tf32_peak_issued"] = 3 * mm["achieved"] / mm_peaks["tf32_tflops"]
    dominant, other = (mm, k1) if mma_ms >= k1_ms else (k1, mm)
    traffic = None
    try:    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = {("mm", True): "project_i8", ("mm", False): "gemm_tf32x3_projection",
               ("k1", True): "bin_motion_planes", ("k1", False): "bin_motion_vec_kernel"}[("mm" if dominant is mm else "k1", proj_i8)]
        traffic = tj[key]["bytes_per_launch"]
        dominant["traffic_launch"] = tj[key].get("launch")
        dominant["algorithmic_bytes_per_launch"] = tj[key].get("algorithmic_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": dominant["bound"], "achieved": dominant["achieved"], "peak": dominant["peak"],
                "unit": dominant["unit"], "frac": dominant.get("frac"), "traffic": traffic, "kernel": dominant["kernel"],
                "peak_source": peak_src, "detail": dominant, "other_kernel": other, "matmul_peaks_this_process": mm_peaks}
    solver_ms = sum(v for k, v in agg.items() if k.endswith(".syevd"))

    # ---- side measurement: strong scaling on ONE 100 000-frame video (north_star's target) ----
    extras = {}

    def timed_passes(fn, n_warm, n_timed):
        for _ in range(n_warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_timed):
            fn()
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / n_timed

    if not args.no_extras and world > 1:
        try:
            F = FRAMES_PER_GPU
            ranges = frame_ranges(F, world)
            # strong-video frame t of rank r's range <-> frame r*F + (t - f0_r) of the generator: every rank already
            # holds its piece at the front of its resident tensor
            imap = [(a, b, r * F) for r, (a, b) in enumerate(ranges)]
            a, b = ranges[rank]
            strong_src = ShardSource([resident[0][:b - a]], F, a, b, synth_gen([synth], device, imap), device)
            tms = {}
            ms = timed_passes(lambda: one_step(strong_src, timing=tms), 3, max(3, min(args.steps, 5)))
            extras["strong_scaling_100k"] = {
                "workload": "ONE 640x480 100 000-frame video, sbin 4, motSVD, cut into frame ranges over the ranks",
                "frames_total": F, "n_gpus": world, "ms_per_step": ms, "frames_per_s": F / (ms * 1e-3),
                "cusolver_ms_per_step_rank0": sum(v for k, v in tms.get("device_ms", {}).items() if k.endswith(".syevd")),
                "stage_ms_rank0": {k: round(v, 2) for k, v in sorted(tms.get("device_ms", {}).items()) if v > 0.5}}
            del strong_src
        except Exception as e:      # noqa: BLE001
            extras["strong_scaling_100k"] = {"error": repr(e)[:300]}
    elif not args.no_extras:
        extras["strong_scaling_100k"] = {"n_gpus": 1, "note": "at N = 1 this is the headline measurement",
                                         "ms_per_step": ms_step, "frames_per_s": value}

    # ---- e2e: host-resident frames through the public API ----
    e2e = None
    if want_host:
        t_pin = time.perf_counter()
        host_frames = torch.empty(tuple(resident[0].shape), dtype=torch.uint8, pin_memory=True)
        pin_s = time.perf_counter() - t_pin
        host_frames.copy_(resident[0])
        torch.cuda.synchronize()
        if world == 1:
            # the call a user makes: facemap_b200.process.run on host frames, `_proc.npy` written
            from facemap_b200.frames import HostArraySource
            import shutil
            savedir = tempfile.mkdtemp(prefix="fm_bench_")
            host_src = HostArraySource([host_frames])
            run_s, out_bytes, last_tm, splits = [], 0, {}, []
            for i in range(2 + args.e2e_steps):                      # 2 warm-up passes
                # every call writes into a fresh folder: deleting the previous 277 MB output (what overwriting the same
                # path amounts to: 0.05 - 0.1 s of page-cache and block freeing) is not part of the call being measured
                stepdir = tempfile.mkdtemp(prefix=f"step{i}_", dir=savedir)
                ts = time.perf_counter()
                with contextlib.redirect_stdout(sys.stderr):         # run() prints its run time like the reference
                    fn = process.run(host_src, sbin=SBIN, motSVD=True, movSVD=False, savepath=stepdir)
                torch.cuda.synchronize()
                if i >= 2:
                    run_s.append(time.perf_counter() - ts)
                    splits.append(json.loads(json.dumps(process.LAST_RUN, default=float),
                                             parse_float=lambda x: round(float(x), 4)))
                out_bytes = os.path.getsize(fn)
                shutil.rmtree(stepdir, ignore_errors=True)
            h2d_total = host_src.h2d
            # where the time goes: the same pass without the file (process_source), then the write alone
            ps_s = []
            for _ in range(2):
                last_tm = {"timeline": True}
                ts = time.perf_counter()
                out = process.process_source(host_src, sbin=SBIN, motSVD=True, movSVD=False, timing=last_tm)
                torch.cuda.synchronize()
                ps_s.append(time.perf_counter() - ts)
            d2h = sum(a.nbytes for key in ("avgframe", "avgmotion", "motion", "motMask", "motSVD") for a in out[key]) \
                + out["motSv"].nbytes
            del out
            try:
                import shutil
                shutil.rmtree(savedir)
            except Exception:
                pass
            dt = sum(run_s)
            e2e = {"value": FRAMES_PER_GPU * len(run_s) / dt, "unit": "frames/s",
                   "h2d_bytes_per_step": int(h2d_total // (2 + args.e2e_steps + 0)), "d2h_bytes_per_step": int(d2h),
                   "steps": len(run_s), "api": "facemap_b200.process.run(HostArraySource([pinned uint8 frames]), savepath=<fresh folder>): "
                                               "H2D of every frame, D2H of every output, _proc.npy written, all inside the timed region",
                   "wall_s_each_step": [round(x, 4) for x in run_s], "proc_file_bytes": int(out_bytes), "savepath": savedir,
                   "run_split_s_each_step": splits,
                   "pin_host_frames_s_outside_timed_region": round(pin_s, 3),
                   "process_source_only": {"frames_per_s": FRAMES_PER_GPU / min(ps_s), "wall_s": [round(x, 4) for x in ps_s],
                                           "note": "same pass without assembling the reference's dict layout into a file"},
                   "upload_ms": round(host_src._cache.upload_ms(), 1),
                   "assemble_wall_s": round(last_tm.get("assemble_wall_s", 0.0), 4),
                   "assemble_host_s_after_device_done": round(last_tm.get("assemble_host_s", 0.0), 4),
                   "stage_ms": {k: round(v, 2) for k, v in sorted(last_tm.get("device_ms", {}).items()) if v > 1.0},
                   "timeline_ms": {k: round(v, 1) for k, v in sorted(last_tm.get("timeline_ms", {}).items(), key=lambda kv: kv[1])
                                   if k.split(".")[-1] in ("fetch", "syevd", "mma", "bin", "d2h")}}
            host_src.close()
        else:
            host_src = ShardSource([host_frames], total_frames, lo, hi, synth_gen([synth], device), device, on_host=True)
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
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            io = torch.tensor([host_src.h2d // args.e2e_steps, int(d2h)], dtype=torch.int64, device=device)
            dist.all_reduce(io, op=dist.ReduceOp.SUM)          # bytes of the whole job, all ranks
            e2e = {"value": total_frames * args.e2e_steps / dt, "unit": "frames/s",
                   "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
                   "steps": args.e2e_steps,
                   "api": "facemap_b200.process.process_source(dist=Shard, gather_outputs=False) on every rank's pinned host shard "
                          "(a whole N x 100k-frame host array per rank does not exist under weak scaling, so no single file is written)",
                   "pin_host_frames_s_outside_timed_region": round(pin_s, 3),
                   "upload_ms": round(host_src.dcache.upload_ms(), 1), "wall_s_last_step": round(e2e_tm.get("wall_s", 0.0), 4),
                   "assemble_wall_s": round(e2e_tm.get("assemble_wall_s", 0.0), 4),
                   "assemble_host_s_after_device_done": round(e2e_tm.get("assemble_host_s", 0.0), 4),
                   "stage_ms": {k: round(v, 2) for k, v in sorted(e2e_tm.get("device_ms", {}).items()) if v > 1.0},
                   "timeline_ms": {k: round(v, 1) for k, v in sorted(e2e_tm.get("timeline_ms", {}).items(), key=lambda kv: kv[1])
                                   if k.split(".")[-1] in ("fetch", "syevd", "mma", "bin", "d2h")}}
            del host_src
        del host_frames

    # ---- side measurements: BASELINE configs[2] and configs[3] (the headline frames are released first) ----
    del dev_src, resident
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_extras:
        import numpy as np

        def roi(ivid, y0, y1, x0, x1):
            return {"rind": 1, "rtype": "motion SVD", "iROI": 0, "ivid": ivid, "color": (0, 0, 0),
                    "yrange": np.arange(y0, y1), "xrange": np.arange(x0, x1), "saturation": 0}

        def side(name, views, frames, sbin, mov, rois_fn, what):
            if frames <= 0:
                return
            try:
                n_tot = world * frames
                syn = [SynthVideo(h, w, n_tot, seed=s) for (h, w, s) in views]
                a, b = rank * frames, (rank + 1) * frames
                held = generate_resident(syn, a, b, device)
                src = ShardSource(held, n_tot, a, b, synth_gen(syn, device), device)
                tms = {}
                fn = lambda: process.process_source(src, sbin=sbin, motSVD=True, movSVD=mov, rois=rois_fn(), dist=shard,
                                                    timing=tms, keep_on_device=True, gather_outputs=False)
                ms = timed_passes(fn, 2, 2)
                dm = tms.get("device_ms", {})
                extras[name] = {"workload": what, "frames_per_gpu": frames, "frames_total": n_tot, "n_gpus": world,
                                "scaling": "weak", "ms_per_step": ms, "frames_per_s": n_tot / (ms * 1e-3),
                                "cusolver_ms_per_step_rank0": round(sum(v for k, v in dm.items() if k.endswith(".syevd")), 2),
                                "stage_ms_rank0": {k: round(v, 2) for k, v in sorted(dm.items()) if v > 1.0},
                                "peak_mem_gb_rank0": round(torch.cuda.max_memory_allocated(device) / 1e9, 1)}
                if world > 1:       # every rank's own view of its last pass: the slowest one sets the step
                    mine = {"wall_ms": round(1e3 * tms.get("wall_s", 0.0), 1),
                            "syevd_ms": round(sum(v for k, v in dm.items() if k.endswith(".syevd")), 1),
                            "chunk_eigh_ms": round(dm.get("stage2.chunk.eigh", 0.0), 1),
                            "bcast_wait_ms": round(dm.get("stage2.merge.bcast", 0.0), 1),
                            "stage3_bin_ms": round(dm.get("stage3.bin", 0.0), 1),
                            "host_marks_ms": tms.get("host_marks_ms", {})}
                    every = [None] * world
                    dist.all_gather_object(every, mine)
                    extras[name]["per_rank"] = every
                pp = sum((h // sbin) * (w // sbin) for (h, w, _) in views)
                f_a, f_b = (0, n_tot) if shard is None else frame_ranges(n_tot, world)[rank]
                if dm.get("stage3.project.mma"):
                    extras[name]["projection_algorithmic_tflops_rank0"] = round(
                        2.0 * (f_b - f_a) * pp * NCOMP * (2 if mov else 1) / (dm["stage3.project.mma"] * 1e-3) / 1e12, 1)
                del src, held
            except Exception as e:      # noqa: BLE001
                extras[name] = {"error": repr(e)[:300]}
            gc.collect()
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(device)

        side("configs2", [(H, W, 0), (H, W, 1)], args.c2_frames, 4, True,
             lambda: [roi(0, 40, 200, 80, 320), roi(1, 32, 160, 64, 300), roi(0, 240, 440, 400, 600)],
             "BASELINE configs[2]: two 640x480 cameras, sbin 4, motion + movie SVD, full frame + 3 motion ROIs "
             f"({args.c2_frames} frames per GPU instead of 200k in total)")
        side("configs3", [(1024, 1280, 0)], args.c3_frames, 1, False, lambda: None,
             "BASELINE configs[3]: 1280x1024 at sbin 1 (1 310 720 binned pixels), motSVD 500 components "
             f"({args.c3_frames} frames per GPU, generated on the device, instead of 500k in total)")

    # ---- CPU baseline (rank 0, single GPU run only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        total_s, detail = cpu_reference_model(rounds=1, warm=0, heavy_rounds=1)
        cpu = {"value": FRAMES_PER_GPU / total_s, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": cpu_sample_text(detail, os.cpu_count()), "model": detail}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": nwarm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32",
            "dtype_detail": "u8 frames, integer block sums / energies; Gram matrices and projections as exact s8/u8 digit-plane "
                            "products on tcgen05 kind::i8 (int32 accumulation, float64 assembly); left singular vectors as 3xTF32 "
                            "products with f64-reduced partials; f64 eigensolves (chunks: hand-written cluster solver; merge: block subspace "
                            "iteration on f64 library GEMMs + Cholesky-QR, projected problem on the cluster solver, cuSOLVER syevdx as "
                            "the fallback)",
            "data": "synthetic",
            "config": {**workload_config(world),
                       **({"rois": "pupil 100x140 + blink 100x140 + running 120x400 (side measurement)"} if args.with_rois else {})},
            "clocks": clk, "e2e": e2e if e2e is not None else ({"value": None, "skipped": e2e_skip} if e2e_skip else None),
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "extras": extras,
            "stage_ms": {k: round(v, 3) for k, v in sorted(agg.items())},
            "cusolver_ms_per_step": solver_ms,
            "host_ms_per_step": [round(tm["host_ms"], 1) for tm in timings],
            "allocator_in_timed_region": alloc_diag,
            "host_marks_ms_slowest_step": max(timings, key=lambda tm: tm["host_ms"]).get("host_marks_ms"),
            "host_marks_ms_fastest_step": min(timings, key=lambda tm: tm["host_ms"]).get("host_marks_ms"),
            "cusolver_ms_each_step": [round(sum(v for k, v in tm["device_ms"].items() if k.endswith(".syevd")), 1)
                                      for tm in timings],
            "kernels_only_frames_per_s": total_frames / ((ms_step - solver_ms) / 1e3) if ms_step > solver_ms else None,
        }
        if world > 1:
            time.sleep(1.0)                      # the other ranks are gone by now (they leave right after the barrier)
        sys.stdout.flush()
        print("\n" + json.dumps(line), flush=True)
    # NCCL (with NCCL_DEBUG=INFO, which the driver may set) logs from its library destructors at process exit, after
    # anything Python can print: leave without running them so that the JSON line stays the last line of stdout
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
