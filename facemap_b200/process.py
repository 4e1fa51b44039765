"""Drop-in for `facemap.process.run` (reference: /root/reference/facemap/process.py:638-903) on the
B200 kernels.  Same signature, same config precedence (parent > proc > kwargs, :676-708), same
`*_proc.npy` dict (keys/dtypes/shapes of :859-892) and the same quirks, written by
`save` (:605-635).  The pupil / blink / running ROIs of `rois` (rind 0, 2, 3) are processed on the same frame
sweep (facemap_b200/fullres.py); keypoints and the GUI's own post-processing are not part of this path.

Beyond the reference signature, `filenames` may also be a FrameSource or a list of uint8
[T,H,W] arrays/tensors (one per simultaneous view), and `process_source` returns the dict without
writing it.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import _lib, fullres, svd
from . import kernels as K
from .kernels import EigCheckFailed
from .frames import FrameSource, as_source
from .utils import binned_inds, multivideo_reshape, roi_bin_ranges, video_placement


class _Segments:
    """file-like sink that keeps what a Pickler writes as a list of buffers without copying them"""

    def __init__(self):
        self.segs = []

    def write(self, b):
        self.segs.append(b if isinstance(b, (bytes, memoryview)) else memoryview(b))
        return len(b) if isinstance(b, bytes) else memoryview(b).nbytes


def save_npy(path, obj, threads=8, piece=32 << 20):
    """`np.save(path, obj)` for an object (the proc dict: facemap/process.py:612) — the same .npy container (version 1.0
    header, object array, pickled payload; `np.load(path, allow_pickle=True).item()` gives the dict back) — written
    faster: pickle protocol 5 hands the big arrays over as buffers instead of copying them into the stream, and the
    pieces go to the file in parallel with positioned writes (277 MB of a 100k-frame video: 0.23 s -> ~0.07 s).
    (Creating the file's page-cache pages ahead of time, with zeros written during the GPU pass, was tried and
    dropped: 4x faster on a quiet file system, slower on the B200 boxes' ext4-on-virtio, where the overwrite waits for
    the writeback of the zero pages — profiles/r02_logs/bench_r2u.json.)"""
    segs = _npy_segments(obj)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.ftruncate(fd, segs[-1][0] + segs[-1][1].nbytes)
        _write_segments(fd, [(off, mv) for off, mv, _ in segs], threads, piece)
    finally:
        os.close(fd)
    return path


def _npy_segments(obj):
    """The bytes of `np.save(path, obj)` for an object, as a list of (file offset, memoryview, address): header and
    pickle stream in the order they go to the file; the big arrays are views of the arrays' own memory (address =
    where that memory starts; None for the pickler's own byte strings)."""
    import io
    import pickle

    from numpy.lib import format as npy_format

    arr = np.empty((), dtype=object)
    arr[()] = obj
    head = io.BytesIO()
    npy_format.write_array_header_1_0(head, npy_format.header_data_from_array_1_0(arr))
    sink = _Segments()
    sink.write(head.getvalue())
    pickle.Pickler(sink, protocol=5).dump(arr)
    out, off = [], 0
    for seg in sink.segs:
        mv = memoryview(seg).cast("B")
        if mv.nbytes == 0:
            continue
        addr = None if isinstance(seg, bytes) else int(np.frombuffer(mv, np.uint8).__array_interface__["data"][0])
        out.append((off, mv, addr))
        off += mv.nbytes
    return out


def _write_segments(fd, jobs, threads=8, piece=32 << 20):
    """positioned writes of [(offset, memoryview)]: small ones in order, big ones cut into pieces on a thread pool"""
    from concurrent.futures import ThreadPoolExecutor

    cut = []
    for off, mv in jobs:
        for a in range(0, mv.nbytes, piece):
            cut.append((off + a, mv[a:a + piece]))

    def put(job):
        pos, mv = job
        while mv.nbytes:
            n = os.pwrite(fd, mv, pos)
            pos, mv = pos + n, mv[n:]

    small = [j for j in cut if j[1].nbytes < (1 << 20)]
    big = [j for j in cut if j[1].nbytes >= (1 << 20)]
    for j in small:
        put(j)
    if big:
        with ThreadPoolExecutor(max_workers=max(1, min(threads, len(big)))) as pool:
            list(pool.map(put, big))


class _EarlyFile:
    """Writes the big arrays of the proc file while the GPU pass is still running.

    A 100k-frame proc file is 277 MB, of which 200 MB are the projections; written after the pass it costs 0.13 - 0.19 s
    of a 0.57 s call, while during the pass the host has nothing to do.  The projections land in page-locked host
    buffers batch by batch (`_HostSink.copy`), the masks right after the merge; process_source announces those buffers
    before stage 3 starts, inside a dict that has the layout of its return value (`begin`), and the file layout follows
    from pickling that dict: every big array sits at an offset that depends on shapes only.  A writer thread then waits
    for each copy's event and writes the rows to their place in `<name>.partial`.  When the pass is over the final dict
    is pickled again: if its layout (every segment's length, the buffer behind every array written early) is the
    announced one, only
    what was not written early — or was changed on the host since (`dirty`) — is written, and the file takes its name;
    if anything differs, or anything failed, the caller writes the file the plain way.  What is in the file is always
    the final pickle of the final dict."""

    BIG = 1 << 16           # the pickler hands over buffers from 64 KB on; smaller arrays are copied into its stream
    WRITERS, PIECE = 4, 8 << 20

    def __init__(self, path, build_full):
        import queue

        self.final_path, self.path, self.build_full = path, path + ".partial", build_full
        self.q, self.thread, self.fd = queue.Queue(), None, None
        self.sig, self.where, self.bases = None, {}, []
        self.waiting, self.done, self.redo, self.used = [], [], [], set()
        self.failed, self.split = False, {}

    @staticmethod
    def _signature(segs):
        return [mv.nbytes for _, mv, _ in segs]

    def begin(self, out_like):
        """out_like: a dict with the layout of process_source's return value whose big arrays ARE the buffers the
        results will land in"""
        import threading

        if self.sig is not None:                # a repeated pass (eigensolver fallback): leave it to the plain writer
            self.failed = True
            return
        try:
            segs = _npy_segments(self.build_full(out_like))
            self.sig = self._signature(segs)
            for off, mv, addr in segs:
                if addr is not None and mv.nbytes >= self.BIG:
                    self.where.setdefault(addr, []).append((off, mv.nbytes))     # a mask and its canvas view: two places
            self.bases = sorted(self.where)
            self.fd = os.open(self.path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            os.ftruncate(self.fd, segs[-1][0] + segs[-1][1].nbytes)
        except Exception:       # noqa: BLE001 — an optimisation: the plain writer takes over
            self.failed = True
            return
        # several writers: one thread writes fresh pages at ~1 GB/s and fell 40 - 57 ms behind a 100k-frame pass
        # (profiles/r02_logs/bench_r2y2.json); positioned writes to different places of a file proceed in parallel
        self.thread = [threading.Thread(target=self._work, daemon=True) for _ in range(self.WRITERS)]
        for t in self.thread:
            t.start()
        waiting, self.waiting = self.waiting, []
        for t, event in waiting:
            self.landed(t, event)

    def landed(self, host_tensor, event):
        """a device-to-host copy into `host_tensor` (contiguous) was enqueued; `event` fires when it is done"""
        if self.thread is None:
            self.waiting.append((host_tensor, event))
            return
        rows = host_tensor.shape[0] if host_tensor.dim() > 1 else 0
        per = host_tensor[0].numel() * host_tensor.element_size() if rows else 0
        step = max(1, self.PIECE // per) if per else 0
        if rows and rows > step:                 # big arrays in pieces, so that the writers share them
            for a in range(0, rows, step):
                self.q.put((host_tensor[a:a + step], event))
        else:
            self.q.put((host_tensor, event))

    def dirty(self, host_array):
        """rows changed on the host after they landed: written again at the end"""
        self.redo.append(host_array)

    def _place(self, addr, nbytes):
        import bisect

        i = bisect.bisect_right(self.bases, addr) - 1
        if i < 0:
            return []
        base = self.bases[i]
        return [(off + addr - base, off, base) for off, n in self.where[base] if addr + nbytes <= base + n]

    def _work(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            try:
                t, event = job
                spots = self._place(t.data_ptr(), t.numel() * t.element_size())
                if not spots:
                    continue
                event.synchronize()
                mv = memoryview(t.numpy()).cast("B")
                for pos, seg_off, base in spots:
                    m = mv
                    p0 = pos
                    while m.nbytes:
                        n = os.pwrite(self.fd, m, p0)
                        p0, m = p0 + n, m[n:]
                    self.done.append((pos, pos + mv.nbytes))
                    self.used.add((seg_off, base))
            except Exception:   # noqa: BLE001
                self.failed = True

    def _stop(self):
        if self.thread is not None:
            for _ in self.thread:
                self.q.put(None)
            for t in self.thread:
                t.join()
            self.thread = None

    def abandon(self):
        self._stop()
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None
            try:
                os.remove(self.path)
            except OSError:
                pass

    def finish(self, full):
        """True: the file is complete under its final name.  False: nothing usable was written (plain writer's turn)."""
        t0 = time.perf_counter()
        self._stop()
        self.split = {"join_s": time.perf_counter() - t0}
        if self.failed or self.sig is None or self.fd is None:
            self.abandon()
            return False
        try:
            t1 = time.perf_counter()
            segs = _npy_segments(full)
            self.split["pickle_s"] = time.perf_counter() - t1
            at = {off: addr for off, _, addr in segs}
            if self._signature(segs) != self.sig or any(at.get(off) != base for off, base in self.used):
                self.abandon()          # not the announced layout, or an early-written array is not the saved one
                return False
            covered = sorted(self.done)
            jobs = []
            for off, mv, addr in segs:
                a = off
                for lo, hi in covered:                       # what the writer thread did not cover
                    if hi <= a or lo >= off + mv.nbytes:
                        continue
                    if lo > a:
                        jobs.append((a, mv[a - off:lo - off]))
                    a = max(a, hi)
                if a < off + mv.nbytes:
                    jobs.append((a, mv[a - off:]))
            for arr in self.redo:
                mv = memoryview(np.ascontiguousarray(arr)).cast("B")
                addr = int(arr.__array_interface__["data"][0])
                for pos, _, _ in self._place(addr, mv.nbytes):
                    jobs.append((pos, mv))
            t2 = time.perf_counter()
            _write_segments(self.fd, jobs)
            self.split["late_bytes"] = sum(mv.nbytes for _, mv in jobs)
            self.split["write_s"] = time.perf_counter() - t2
            t3 = time.perf_counter()
            os.close(self.fd)
            self.fd = None
            os.replace(self.path, self.final_path)
            self.split["close_rename_s"] = time.perf_counter() - t3
            return True
        except Exception:       # noqa: BLE001
            self.abandon()
            return False


def _proc_file_name(names, savepath):
    folder, name = os.path.split(names[0][0])
    return (folder if savepath is None else savepath), os.path.splitext(name)[0]


def save(proc, savepath=None, npy_written=False):
    """Write the proc dict next to the first video (or into `savepath`) as `<name>_proc.npy`, plus
    a flattened `.mat` when proc["save_mat"] (reference process.py:605-635).  npy_written: the .npy file is
    already there (run: written while the pass ran)."""
    folder, stem = _proc_file_name(proc["filenames"], savepath)
    savename = os.path.join(folder, f"{stem}_proc.npy")
    if not npy_written:
        save_npy(savename, proc)
    if proc["save_mat"]:
        from scipy import io

        if "save_path" in proc and proc["save_path"] is None:
            proc["save_path"] = folder
        if proc["rois"] is None:
            proc["rois"] = 0
        flat = {}
        for key, val in proc.items():
            if isinstance(val, list) and len(val) > 0 and isinstance(val[0], np.ndarray):
                for i, a in enumerate(val):
                    flat[f"{key}_{i}"] = a
            else:
                flat[key] = val
        io.savemat(os.path.join(folder, f"{stem}_proc.mat"), flat)
    return savename


def _motion_from_energy(E_views, sbin, N, f0=0, ref_views=None):
    """Reference float semantics of `M[.] += imbin_mot.sum(-1)` per view (process.py:460) applied
    to exact integer energies, then the chunk-0 layout (quirk Q1): entries [0, nt1-1) hold the
    diffs of frames 1..nt1-1, entry nt1-1 stays 0.  E_views cover frames [f0, f0 + len) of an
    N-frame video (a rank's shard when the outputs are not gathered).
    ref_views[v]: optional float64 sums in the reference's own arithmetic (fm_motion_ref) for the views whose
    float32 accumulation the integer energy cannot decide (svd.motion_needs_reference)."""
    s2 = float(sbin * sbin)
    n = len(E_views[0])
    acc = np.zeros(n, np.float32)
    for v, e in enumerate(E_views):
        ref = None if ref_views is None else ref_views[v]
        if sbin == 1:                                           # float32 sums in the reference
            acc = acc + (e.astype(np.float32) if ref is None else ref.astype(np.float32))
        else:
            acc = (acc.astype(np.float64) + (e.astype(np.float64) / s2 if ref is None else ref)).astype(np.float32)
    nt1 = min(500, N)
    out = acc.copy()
    if f0 < nt1:                                                # this range holds (part of) the first chunk
        hi = min(nt1 - 1, f0 + n - 1)                           # global frames f0 .. hi-1 take the next frame's energy
        if hi > f0:
            out[:hi - f0] = acc[1:hi - f0 + 1]
        if nt1 - 1 - f0 < n:
            out[nt1 - 1 - f0] = 0
    return out


class _PinnedPool:
    """Page-locked host buffers that outlive a call.  Page-locking is the expensive part of a result buffer (about
    5 GB/s, and the ranks of one box queue behind each other in the kernel for it: 0.18 -> 0.91 s of a pass on 8 GPUs),
    so a buffer goes back to the pool when the LAST reference to what was handed out dies — the tensor itself, or any
    NumPy array made from it (an ndarray from `Tensor.numpy()` keeps its tensor alive) — and the next call takes it
    again.  A result the caller still holds is never reused.  `allocate(nbytes)` is replaceable for tests."""

    GRAIN = 2 << 20           # sizes are rounded up so that calls with slightly different frame counts share buffers
    SMALL = 4096

    def __init__(self, allocate=None, keep_bytes=4 << 30):
        import threading

        self.allocate = allocate or (lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True))
        self.keep_bytes = keep_bytes          # free buffers kept for the next call; beyond that they are dropped
        self.free, self.free_bytes = {}, 0
        self.lock = threading.Lock()
        self.stats = {"allocated": 0, "reused": 0}

    def _size(self, n):
        if n >= self.GRAIN:
            return -(-n // self.GRAIN) * self.GRAIN
        size = self.SMALL
        while size < n:
            size *= 2
        return size

    def _give_back(self, size, base):
        with self.lock:
            if self.free_bytes + size <= self.keep_bytes:
                self.free.setdefault(size, []).append(base)
                self.free_bytes += size

    def take(self, shape, dtype=torch.float32):
        import weakref

        shape = tuple(int(d) for d in shape)
        n = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        if n == 0:
            return torch.empty(shape, dtype=dtype)
        size = self._size(n)
        with self.lock:
            bucket = self.free.get(size)
            base = bucket.pop() if bucket else None
            if base is not None:
                self.free_bytes -= size
                self.stats["reused"] += 1
        if base is None:
            base = self.allocate(size)
            self.stats["allocated"] += 1
        # what is handed out hangs on a lease object: NumPy keeps the provider of an __array_interface__ alive as the
        # base of every array derived from it, torch.from_numpy keeps that array alive in the tensor's storage, and
        # Tensor.numpy() keeps the storage alive in turn — so the lease dies with the last user, whatever its type
        lease = _Lease(base, n)
        weakref.finalize(lease, self._give_back, size, base)
        arr = np.asarray(lease).view(_NP_DTYPE[dtype]).reshape(shape)
        del lease
        return torch.from_numpy(arr)


class _Lease:
    def __init__(self, base, n):
        self.__array_interface__ = {"data": (base.data_ptr(), False), "shape": (n,), "typestr": "|u1", "version": 3}


_NP_DTYPE = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64,
             torch.uint8: np.uint8, torch.int16: np.int16}


_POOL = _PinnedPool()


class _HostSink:
    """Asynchronous device->host copies into page-locked arrays on a private stream: results leave
    the GPU while later batches are still being computed, and land in pinned memory (fast DMA) taken from a pool
    that persists across calls (_PinnedPool)."""

    def __init__(self, device, listener=None):
        self.stream = K.shared_stream(device, "host_sink")
        self.listener = listener            # _EarlyFile: told about every copy, with an event that fires when it is done

    @staticmethod
    def alloc(shape, dtype=torch.float32):
        return _POOL.take(shape, dtype)

    def copy(self, dev_tensor, host_tensor):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            host_tensor.copy_(dev_tensor, non_blocking=True)
            if self.listener is not None and host_tensor.is_contiguous():
                done = torch.cuda.Event()
                done.record(self.stream)
                self.listener.landed(host_tensor, done)
        dev_tensor.record_stream(self.stream)

    def to_host(self, dev_tensor):
        h = self.alloc(dev_tensor.shape, dev_tensor.dtype)
        self.copy(dev_tensor, h)
        return h

    def wait(self):
        self.stream.synchronize()


def process_source(source, sbin=1, motSVD=True, movSVD=False, rois=None, fullSVD=True, device=None,
                   eig_fp32=False, gemm_impl=0, timing=None, dist=None, keep_on_device=False,
                   overlap_stage3=True, gather_outputs=True, gram_mode=None, proj_mode=None, eig_mode=None,
                   stream_out=None):
    """Run the three stages on a FrameSource and return the proc-dict entries this path owns.

    timing: optional dict that receives per-stage device times (ms) and wall times.
    gather_outputs: with `dist`, all-gather projections / energies so every rank returns the whole video
            (what `run` needs for the file); False keeps each rank's frame range only (proc["frame_range"]).
    overlap_stage3: run stage 3's bin/diff pass on a side stream underneath the stage-2 eigensolves
            when the frames are HBM-resident (svd.Stage3.start_early).
    dist:   optional facemap_b200.parallel.Shard (frame-range sharding across ranks).
    gram_mode: "i8" (default: exact integer products on tcgen05 kind::i8), "i8_merge", "tf32x3", "sliced_merge" or
            "sliced" — how the Gram matrices of the two SVD levels are formed (svd.SvdWorkspace, svd.i8_gram,
            svd.sliced_gram); None reads the environment variable FM_GRAM_MODE.
    proj_mode: "i8" (default: stage 3 on byte planes and tcgen05 kind::i8, svd.Stage3) or "tf32x3" (float32 operands,
            3xTF32 products); None reads FM_PROJ_MODE.
    eig_mode: "cluster" (default: the hand-written batched eigensolver for the per-chunk Gram matrices,
            csrc/eigh_cluster.cu) or "cusolver"; None reads FM_EIG.  If the hand-written solver's self-check flags a
            matrix (exactly repeated eigenvalues), the pass is repeated with cuSOLVER — on every rank.
    stream_out: optional _EarlyFile (run): told before stage 3 which host buffers the big results will land in, and
            of every device-to-host copy, so that the proc file is written while the pass runs."""
    call = dict(sbin=sbin, motSVD=motSVD, movSVD=movSVD, rois=rois, fullSVD=fullSVD, device=device, eig_fp32=eig_fp32,
                gemm_impl=gemm_impl, timing=timing, dist=dist, keep_on_device=keep_on_device,
                overlap_stage3=overlap_stage3, gather_outputs=gather_outputs, gram_mode=gram_mode, proj_mode=proj_mode,
                stream_out=stream_out)
    t_enter = time.perf_counter()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _lib.require_device(dev.index or 0)
    torch.cuda.set_device(dev)
    source = as_source(source)
    N = source.nframes
    Ly, Lx = list(source.Ly), list(source.Lx)
    Lybin, Lxbin, iinds = binned_inds(Ly, Lx, sbin)
    LYbin, LXbin, sybin, sxbin = video_placement(Lybin, Lxbin)
    nroi = 0
    if rois is not None:
        for r in rois:
            if r["rind"] == 1:
                r["yrange_bin"], r["xrange_bin"] = roi_bin_ranges(r, sbin)   # mutated in place, like the reference
                nroi += 1
    geo = svd.Geometry(Ly, Lx, sbin, rois)
    has_fullres = any(r["rind"] in (fullres.PUPIL, fullres.BLINK, fullres.RUNNING) for r in (rois or []))
    mods = [m for m, on in (("mot", motSVD), ("mov", movSVD)) if on]
    timer = svd.StageTimer(enabled=timing is not None)
    wall0 = time.perf_counter()
    marks = []

    hmarks = []

    def mark(name):
        """host timestamps per stage (timing["host_marks_ms"]); with timing["sync"] the device is
        synchronised first, which turns them into per-stage wall times (diagnostic)"""
        if timing is not None:
            hmarks.append((name, time.perf_counter()))
            if timing.get("sync"):
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

    mark("start")
    ws = svd.SvdWorkspace(dev, eig_fp32=eig_fp32, gemm_impl=gemm_impl, timer=timer, gram_mode=gram_mode,
                          proj_mode=proj_mode, eig_mode=eig_mode)
    source.plan_access(*svd.access_plan(source, dist), dev)
    mark("setup")

    # ---- stage 1 ----
    if dist is None or dist.world == 1:
        avgframe, avgmotion = svd.stage1_means(source, geo, dev, timer)
    else:
        avgframe, avgmotion = dist.stage1(source, geo, dev, timer)

    mark("stage1")
    # stage 3's bin phase depends on the averages only: start it underneath the stage-2 eigensolves
    s3 = None
    fr = None
    if dist is not None and dist.world > 1:
        from .parallel import frame_ranges
        fr = frame_ranges(N, dist.world)[dist.rank]
    fres = fullres.FullRes(rois, Ly, Lx, N, dev, frame_range=fr) if has_fullres else None
    if mods and (fullSVD or nroi > 0):
        s3 = svd.Stage3(source, geo, avgframe, avgmotion, mods, fullSVD, ws, dev, timer, frame_range=fr, fullres=fres)
        # with cuSOLVER solving the chunks the SMs idle during all of stage 2; the hand-written chunk solver fills
        # them (one cluster per matrix), so the bin pass then starts after the chunk phase, underneath the merge solve
        if overlap_stage3 and ws.eig_mode != "cluster":
            s3.start_early()
    # ---- stage 2 ----
    masks = {}
    if (fullSVD or nroi > 0) and mods:
        if dist is None or dist.world == 1:
            Yt, ks, nsegs = svd.stage2_chunks(source, geo, avgframe, avgmotion, mods, fullSVD, ws, dev, timer)
        else:
            Yt, ks, nsegs = dist.stage2_chunks(source, geo, avgframe, avgmotion, mods, fullSVD, ws, dev, timer)
        mark("stage2.chunks")
        if s3 is not None and overlap_stage3 and ws.eig_mode == "cluster":
            s3.start_early()
        if geo.p > svd.LEFT_BLOCK_PIXELS:
            ws.release()               # large views: the chunk phase's buffers must not add to the merge's
        masks = svd.stage2_merge(Yt, ks, nsegs, geo, mods, fullSVD, ws, timer, nc=svd.stage2_plan(N)[2], mot_on=bool(motSVD),
                                 shard=dist if (dist is not None and dist.world > 1) else None)
        del Yt
        if geo.p > svd.LEFT_BLOCK_PIXELS:
            ws.release()
        mark("stage2.merge")
    nsegs_plan = svd.stage2_plan(N)[1]
    nc_plan = svd.stage2_plan(N)[2]

    # ---- stage 3 ----
    sink = None if keep_on_device else _HostSink(dev, listener=stream_out)
    pending_masks, pending_canvas = {}, {}
    V_early = None
    V, E, E_roi, Eref = {}, None, None, None
    if mods and masks:
        single = dist is None or dist.world == 1 or not gather_outputs
        if sink is not None:
            # masks and singular values leave while stage 3 runs; projections leave batch by batch
            pending_masks = {m: ([sink.to_host(u) for u in masks[m][1]], sink.to_host(masks[m][2])) for m in masks}
            if fullSVD and len(Ly) > 1:
                # several views: the tiled (LYbin, LXbin, ncomp) canvas of utils.multivideo_reshape (utils.py:622-630) is
                # laid out on the device — one strided copy per view — and leaves as one contiguous transfer, instead
                # of a host-side scatter of the 77 MB mask matrix after the GPU is done
                for m in masks:
                    U_dev = masks[m][1][0]
                    canvas = torch.zeros((int(LYbin), int(LXbin), U_dev.shape[1]), dtype=torch.float32, device=dev)
                    for v in range(len(Ly)):
                        c0, n = geo.col0[v], geo.npix_v[v]
                        canvas[int(sybin[v]):int(sybin[v]) + int(Lybin[v]), int(sxbin[v]):int(sxbin[v]) + int(Lxbin[v])] = \
                            U_dev[c0:c0 + n].view(int(Lybin[v]), int(Lxbin[v]), -1)
                    pending_canvas[m] = sink.to_host(canvas)
        if stream_out is not None and sink is not None and single and fullSVD and nroi == 0 and fres is None \
                and (dist is None or dist.world == 1):
            # the simple layout (no ROIs of any kind, one rank): announce where the results will land, in a dict shaped
            # like the one returned below (_EarlyFile.begin; a wrong guess only costs the early writing)
            V_early = {m: [None if Ut is None else sink.alloc((s3.n_out, Ut.shape[0])) for Ut in masks[m][0]]
                       for m in mods}
            like = {"Ly": Ly, "Lx": Lx, "sbin": sbin, "fullSVD": fullSVD, "Lybin": Lybin, "Lxbin": Lxbin, "sybin": sybin,
                    "sxbin": sxbin, "LYbin": LYbin, "LXbin": LXbin, "rois": rois, "pupil": [], "blink": [], "running": [],
                    "avgframe": [np.empty(len(ir), np.float32) for ir in iinds],
                    "avgmotion": [np.empty(len(ir), np.float32) for ir in iinds],
                    "motion": [np.empty(N, np.float32)]}                 # one trace per target: the full frame only
            flat = np.empty((sum(len(ir) for ir in iinds), 1), np.float32)
            for key in ("avgframe_reshape", "avgmotion_reshape"):
                like[key] = np.squeeze(multivideo_reshape(flat, LYbin, LXbin, sybin, sxbin, Lybin, Lxbin, iinds))
            for m in ("mot", "mov"):
                if m in mods:
                    Us, sv = pending_masks[m]
                    like[m + "Mask"] = [u.numpy() for u in Us]
                    like[m + "Sv"] = sv.numpy()
                    like[m + "Mask_reshape"] = [pending_canvas[m].numpy() if m in pending_canvas else
                                                Us[0].numpy().reshape(int(Lybin[0]), int(Lxbin[0]), -1)]
                    like[m + "SVD"] = [v.numpy() for v in V_early[m]]
                else:
                    like[m + "Mask"], like[m + "Mask_reshape"], like[m + "SVD"] = [], [], []
                    like[m + "Sv"] = np.zeros(500, np.float32)
            stream_out.begin(like)
        V, E, E_roi = s3.project(masks, sink=sink if single else None, V_host=V_early)
        Eref = s3.Eref
        if not single:
            V, E, E_roi = dist.gather_stage3(V, E, E_roi, N, mods, timer)
            if Eref is not None:
                Eref = dist.gather_rows(Eref.t().contiguous(), frame_ranges(N, dist.world)).t().contiguous()

    # ---- pupil / blink / running: rode on stage 3's sweep, or get their own when there is no SVD sweep ----
    fres_out = ([], [], [])
    if fres is not None:
        if s3 is None:
            with timer("fullres.sweep"):
                fres.sweep(source)
        fres_out = fres.device_outputs()
        if dist is not None and dist.world > 1 and gather_outputs:
            ranges = frame_ranges(N, dist.world)
            fres_out = tuple([dist.gather_rows(a, ranges) for a in group] for group in fres_out)

    mark("stage3")
    # ---- assemble (device -> host) ----
    f0_out = 0
    if dist is not None and dist.world > 1 and not gather_outputs and s3 is not None:
        f0_out = s3.f0
    t_asm = time.perf_counter()
    # the eigensolves ran asynchronously; their devInfo words / self-check flags are read once, here, where the host
    # waits anyway
    eig_failed = False
    try:
        ws.solver.check()
    except EigCheckFailed:
        eig_failed = True
    t_dev_done = time.perf_counter()       # check() waited for the stream: stage 3's last projection has run
    if dist is not None and dist.world > 1:
        eig_failed = dist.any_rank(eig_failed, dev)
    if eig_failed:
        if ws.eig_mode != "cluster":
            raise _lib.FacemapB200Error("eigensolver self-check failed")
        timer.release()
        return process_source(source, eig_mode="cusolver", **call)
    with timer("assemble.d2h"):
        avgframe_h = avgframe.cpu().numpy()
        avgmotion_h = avgmotion.cpu().numpy()
        E_h = E.cpu().numpy() if E is not None else None
        E_roi_h = E_roi.cpu().numpy() if E_roi is not None else None
        Eref_h = Eref.cpu().numpy() if Eref is not None else None
    avgframe_l = [avgframe_h[ir].copy() for ir in iinds]
    avgmotion_l = [avgmotion_h[ir].copy() for ir in iinds]
    proc = {
        "Ly": Ly, "Lx": Lx, "sbin": sbin, "fullSVD": fullSVD,
        "Lybin": Lybin, "Lxbin": Lxbin, "sybin": sybin, "sxbin": sxbin, "LYbin": LYbin, "LXbin": LXbin,
        "avgframe": avgframe_l, "avgmotion": avgmotion_l,
        "avgframe_reshape": np.squeeze(multivideo_reshape(avgframe_h[:, None], LYbin, LXbin, sybin, sxbin,
                                                          Lybin, Lxbin, iinds)),
        "avgmotion_reshape": np.squeeze(multivideo_reshape(avgmotion_h[:, None], LYbin, LXbin, sybin, sxbin,
                                                           Lybin, Lxbin, iinds)),
        "rois": rois,
    }
    if dist is not None and dist.world > 1 and not gather_outputs and (s3 is not None or fres is not None):
        proc["frame_range"] = tuple(fr)          # motSVD / movSVD / motion / ROI rows cover this range only

    def host(t):
        if keep_on_device:
            return t
        return t.numpy() if not t.is_cuda else sink.to_host(t)       # pinned tensors are unwrapped below

    def unwrap(x):
        return x.numpy() if torch.is_tensor(x) and not x.is_cuda else x

    U_out = {"mot": [], "mov": []}
    U_resh = {"mot": [], "mov": []}
    S_out = {"mot": [], "mov": []}
    V_out = {"mot": [], "mov": []}
    if fullSVD or nroi > 0:
        S_out = {"mot": np.zeros(500, np.float32), "mov": np.zeros(500, np.float32)}
        for m in ("mot", "mov"):
            if m in masks:
                _, Us, sv = masks[m]
                with timer("assemble.d2h"):
                    if m in pending_masks:
                        U_out[m], S_out[m] = pending_masks[m]
                    else:
                        U_out[m] = [host(u) for u in Us]
                        S_out[m] = host(sv)
            elif m not in mods:
                # quirk Q7: ROI placeholders exist for a modality that is switched off
                for (_, y0, y1, x0, x1) in geo.mot_rois:
                    n = (y1 - y0) * (x1 - x0)
                    U_out[m].append(np.zeros((n, nsegs_plan * min(nc_plan, n)), np.float32))
        if sink is not None:
            sink.wait()
            for m in ("mot", "mov"):
                U_out[m] = [unwrap(u) for u in U_out[m]]
                S_out[m] = unwrap(S_out[m])
        for m in ("mot", "mov"):
            U_resh[m] = list(U_out[m])
            if m in mods and not keep_on_device:
                if fullSVD and len(Ly) == 1:
                    # one view: the canvas IS the mask matrix seen as (Lyb, Lxb, ncomp) — no copy
                    U_resh[m][0] = U_out[m][0].reshape(int(Lybin[0]), int(Lxbin[0]), -1)
                elif fullSVD and m in pending_canvas:
                    U_resh[m][0] = unwrap(pending_canvas[m])
                elif fullSVD:
                    U_resh[m][0] = multivideo_reshape(U_out[m][0], LYbin, LXbin, sybin, sxbin, Lybin, Lxbin, iinds)
                for j, (_, y0, y1, x0, x1) in enumerate(geo.mot_rois):
                    U_resh[m][1 + j] = U_out[m][1 + j].copy().reshape(y1 - y0, x1 - x0, -1)
    # projections and motion energy
    M = []
    if mods and masks:
        for m in mods:
            for idx, v in enumerate(V[m]):
                if v is None:
                    V_out[m].append(np.zeros((0, 1), np.float32))
                    continue
                if v.is_cuda and not keep_on_device:
                    with timer("assemble.d2h"):
                        v = sink.to_host(v)
                        sink.wait()
                if torch.is_tensor(v) and not v.is_cuda:
                    v = v.numpy()
                if N > 1 and f0_out == 0 and v.shape[0] > 1:
                    v[0] = v[1]                         # quirk Q2: first projected row duplicated
                    if stream_out is not None:
                        stream_out.dirty(v[0:1])
                V_out[m].append(v)
        if fullSVD:
            refs = None if Eref_h is None else [Eref_h[v] if v in s3.ref_views else None for v in range(len(Ly))]
            M.append(_motion_from_energy([E_h[v] for v in range(len(Ly))], sbin, N, f0_out, refs) if motSVD
                     else np.zeros(E_h.shape[1], np.float32))
        else:
            M.append(np.zeros((0,), np.float32))
        for j in range(nroi):
            M.append(_motion_from_energy([E_roi_h[j]], sbin, N, f0_out) if motSVD
                     else np.zeros(E_roi_h.shape[1], np.float32))
        # quirk Q4: the avg arrays of a view carrying a motion ROI are saved 2-D
        for v in range(len(Ly)):
            if geo.rois_on_view[v]:
                if motSVD:
                    avgmotion_l[v] = avgmotion_l[v].reshape(Lybin[v], Lxbin[v])
                if movSVD:
                    avgframe_l[v] = avgframe_l[v].reshape(Lybin[v], Lxbin[v])
    elif not fullSVD and nroi == 0:
        # process.py:346-349: without any SVD the reference still returns these empty placeholders
        for m in mods:
            V_out[m].append(np.zeros((0, 1), np.float32))
        M.append(np.zeros((0,), np.float32))
    if keep_on_device:
        pups, blinks, runs = fres_out
    else:
        with timer("assemble.d2h"):
            raw = tuple([a.cpu().numpy() for a in group] for group in fres_out)
        # only a whole video can be smoothed (a shard's traces are left raw)
        pups, blinks, runs = fullres.assemble(*raw, smooth="frame_range" not in proc)
    proc.update({
        "pupil": pups, "blink": blinks, "running": runs,
        "motion": M, "motSv": S_out["mot"], "movSv": S_out["mov"],
        "motMask": U_out["mot"], "movMask": U_out["mov"],
        "motMask_reshape": U_resh["mot"], "movMask_reshape": U_resh["mov"],
        "motSVD": V_out["mot"], "movSVD": V_out["mov"],
    })
    mark("assemble")
    if timing is not None:
        if marks:
            timing["wall_ms"] = {b[0]: 1e3 * (b[1] - a[1]) for a, b in zip(marks[:-1], marks[1:])}
        timing["host_marks_ms"] = {b[0]: round(1e3 * (b[1] - a[1]), 2) for a, b in zip(hmarks[:-1], hmarks[1:])}
        timing["host_marks_ms"]["preamble"] = round(1e3 * (hmarks[0][1] - t_enter), 2)
        t_rep = time.perf_counter()
        torch.cuda.synchronize()
        timing["final_sync_ms"] = round(1e3 * (time.perf_counter() - t_rep), 2)
        timing["device_ms"] = timer.report()
        timing["report_ms"] = round(1e3 * (time.perf_counter() - t_rep), 2)
        if timing.get("timeline"):
            timing["timeline_ms"] = timer.timeline()
        timer.release()
        timing["wall_s"] = time.perf_counter() - wall0
        timing["assemble_wall_s"] = time.perf_counter() - t_asm            # includes waiting for the device
        timing["assemble_host_s"] = time.perf_counter() - t_dev_done       # host work after the device was done
    return proc


LAST_RUN = {}       # wall-clock split of the last run() call (diagnostic: bench.py reports it)


def run(filenames, sbin=1, motSVD=True, movSVD=False, GUIobject=None, parent=None, proc=None, savepath=None):
    """Process video files: SVD of the motion and/or raw movie, projections of every frame, motion
    energy.  Arguments, precedence and return value as facemap.process.run (process.py:638-708):

    filenames : 2-D list, one inner list of simultaneously recorded views per sequential block
    sbin      : spatial bin factor;  motSVD / movSVD : which decompositions to compute
    parent    : GUI main window (settings are read from its widgets);  GUIobject : Qt module
    proc      : settings dict (sbin, fullSVD, save_mat, rois, sy, sx, savepath) overriding kwargs
    savepath  : output folder (default: folder of the first video)
    Returns the path of the written `_proc.npy`."""
    start = time.time()
    start_pc = time.perf_counter()
    rois = None
    sy, sx = 0, 0
    if parent is not None:
        from facemap import utils as ref_utils  # only reachable from the reference GUI

        filenames = parent.filenames
        sbin = parent.sbin
        rois = ref_utils.roi_to_dict(parent.ROIs, parent.rROI)
        fullSVD = parent.multivideo_svd_checkbox.isChecked()
        save_mat = parent.save_mat.isChecked()
        sy, sx = parent.sy, parent.sx
        motSVD, movSVD = parent.motSVD_checkbox.isChecked(), parent.movSVD_checkbox.isChecked()
    elif proc is None:
        fullSVD, save_mat = True, False
    else:
        sbin, fullSVD, save_mat, rois = proc["sbin"], proc["fullSVD"], proc["save_mat"], proc["rois"]
        sy, sx = proc["sy"], proc["sx"]
        savepath = proc["savepath"] if savepath is None else savepath

    source = as_source(filenames)
    names = filenames if not isinstance(filenames, FrameSource) and isinstance(filenames[0], (list, tuple)) \
        else [["array_input.avi"]]

    def build_full(out):
        return {
            "filenames": names, "save_path": savepath, "Ly": out["Ly"], "Lx": out["Lx"], "sbin": sbin,
            "fullSVD": fullSVD, "save_mat": save_mat,
            "Lybin": out["Lybin"], "Lxbin": out["Lxbin"], "sybin": out["sybin"], "sxbin": out["sxbin"],
            "LYbin": out["LYbin"], "LXbin": out["LXbin"],
            "avgframe": out["avgframe"], "avgmotion": out["avgmotion"],
            "avgframe_reshape": out["avgframe_reshape"], "avgmotion_reshape": out["avgmotion_reshape"],
            "motion": out["motion"], "motSv": out["motSv"], "movSv": out["movSv"],
            "motMask": out["motMask"], "movMask": out["movMask"],
            "motMask_reshape": out["motMask_reshape"], "movMask_reshape": out["movMask_reshape"],
            "motSVD": out["motSVD"], "movSVD": out["movSVD"],
            "pupil": out["pupil"], "running": out["running"], "blink": out["blink"], "rois": rois, "sy": sy, "sx": sx,
        }

    # the big arrays go to the file while the pass runs (_EarlyFile); anything it cannot do is written afterwards
    early = None
    try:
        folder, stem = _proc_file_name(names, savepath)
        if os.path.isdir(folder or ".") and os.environ.get("FM_EARLY_FILE", "1") != "0":
            early = _EarlyFile(os.path.join(folder, f"{stem}_proc.npy"), build_full)
    except Exception:           # noqa: BLE001
        early = None
    t_pass = time.perf_counter()
    try:
        out = process_source(source, sbin=sbin, motSVD=motSVD, movSVD=movSVD, rois=rois, fullSVD=fullSVD,
                             stream_out=early)
    except BaseException:
        if early is not None:
            early.abandon()
        raise
    finally:
        t_close = time.perf_counter()
        source.close()
    t_save = time.perf_counter()
    if parent is not None:
        parent.update_status_bar("Computed projection")
    if GUIobject is not None:
        GUIobject.QApplication.processEvents()
    full = build_full(out)
    written_early = early is not None and early.finish(full)
    savename = save(full, savepath, npy_written=written_early)
    LAST_RUN.clear()
    LAST_RUN.update({"setup_s": t_pass - start_pc, "pass_s": t_close - t_pass, "close_s": t_save - t_close,
                     "save_s": time.perf_counter() - t_save, "written_early": bool(written_early),
                     "early_bytes": 0 if early is None else sum(b - a for a, b in early.done),
                     "early_finish": {} if early is None else dict(early.split)})
    if parent is not None:
        parent.update_status_bar("Output saved in " + str(savepath))
    if GUIobject is not None:
        GUIobject.QApplication.processEvents()
    print("run time %0.2fs" % (time.time() - start))
    return savename
