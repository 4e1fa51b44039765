"""Frame sources: the seam that replaces the reference's `containers` (lists of cv2.VideoCapture)
and `utils.get_frames` (facemap/utils.py:522-553) / `get_frame_details` (:567-591).

A source hands out uint8 device tensors [T, H, W] per view for a frame range.  Ranges are clamped
to the video like the reference does (utils.py:523-525, 551-553).  Three kinds:

  DeviceArraySource   views already resident in HBM (slices are zero-copy)
  HostArraySource     views in host memory (NumPy / torch CPU); staged through pinned buffers
  VideoFileSource     video files decoded with OpenCV on the host, sequential blocks stitched
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class FrameSource:
    nframes: int
    Ly: list
    Lx: list
    block_frames: list  # frames per sequential block (stage 1 uses the shortest, process.py:63)

    @property
    def nviews(self):
        return len(self.Ly)

    def clamp(self, t0, t1):
        """[t0, t1) -> the contiguous in-range window the reference would read."""
        a = max(0, min(self.nframes - 1, int(t0)))
        b = max(0, min(self.nframes - 1, int(t1) - 1))
        return a, b + 1

    def fetch(self, t0, t1, device):
        """list (one per view) of contiguous uint8 CUDA tensors [t1'-t0', H, W]."""
        raise NotImplementedError

    def resident(self, t0, t1):
        """True if fetch() of any sub-range of [t0, t1) is a zero-copy view of HBM-resident frames."""
        return False

    def begin_sweep(self):
        """Called once before the sequential sweep over the whole frame range (stage 3); file sources use it to
        make sure the sweep sees what a sequential decode delivers."""

    def prefetch(self, t0, t1, device):
        """Hint that [t0, t1) is the next range to be fetched (host sources start the H2D copy on a
        side stream so it overlaps the kernels of the current range)."""

    def plan_access(self, priority, home, device):
        """Announce the access pattern of one pass: `priority` = the (t0, t1) ranges stage 1 and 2
        will read, in order; `home` = the contiguous range stage 3 will sweep.  Host sources use it
        to upload every frame ONCE, sample ranges first, while the kernels already run."""

    def h2d_bytes(self, nframes):
        """bytes that cross PCIe to deliver `nframes` frames (0 for device-resident sources)."""
        return 0

    def close(self):
        pass


class DeviceArraySource(FrameSource):
    def __init__(self, views, block_frames=None):
        self.views = [v if torch.is_tensor(v) else torch.as_tensor(v) for v in views]
        for v in self.views:
            if not (v.is_cuda and v.dtype == torch.uint8 and v.dim() == 3 and v.is_contiguous()):
                raise ValueError("DeviceArraySource needs contiguous CUDA uint8 [T,H,W] views")
        self.nframes = int(self.views[0].shape[0])
        if any(int(v.shape[0]) != self.nframes for v in self.views):
            raise ValueError("simultaneous views must have the same number of frames")
        self.Ly = [int(v.shape[1]) for v in self.views]
        self.Lx = [int(v.shape[2]) for v in self.views]
        self.block_frames = list(block_frames) if block_frames is not None else [self.nframes]

    def fetch(self, t0, t1, device):
        a, b = self.clamp(t0, t1)
        return [v[a:b] for v in self.views]

    def resident(self, t0, t1):
        return True


class _H2DPipe:
    """Asynchronous host->device copies on a private stream, handed over to the consumer stream
    with an event (double buffering comes from the caching allocator + record_stream)."""

    def __init__(self):
        self.stream = None
        self.pending = {}

    def issue(self, key, host_tensors, device):
        if key in self.pending:
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream(device)
        with torch.cuda.stream(self.stream):
            devs = [h.to(device, non_blocking=True) for h in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending[key] = (devs, ev)

    def take(self, key):
        devs, ev = self.pending.pop(key)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for d in devs:
            d.record_stream(cur)
        return devs


class DeviceFrameCache:
    """One-pass upload of a contiguous `home` frame range into HBM in pieces of `piece` frames on
    a private copy stream: the pieces the SVD sample ranges touch go first, the rest follows in
    frame order, and readers wait on per-piece events — so the PCIe transfer of the whole video
    overlaps stages 1-3 and every frame crosses the bus once."""

    def __init__(self, piece=1024, mem_fraction=0.55):
        self.piece = piece
        self.mem_fraction = mem_fraction
        self.stream = None
        self.home = None
        self.bufs = None
        self.events = None
        self.bytes = 0

    def fits(self, nbytes, device):
        free, _ = torch.cuda.mem_get_info(device)
        return nbytes <= self.mem_fraction * free

    def start(self, host_slice, shapes, priority, home, device):
        """host_slice(view, a, b) -> pinned uint8 [b-a, H, W]; shapes = [(H, W)] per view."""
        lo, hi = home
        if hi <= lo:
            return False
        nbytes = (hi - lo) * sum(h * w for h, w in shapes)
        self.bufs = self.events = self.home = None            # drop the previous pass first
        if not self.fits(nbytes, device):
            return False
        if self.stream is None:
            self.stream = torch.cuda.Stream(device)
        bufs = [torch.empty((hi - lo, h, w), dtype=torch.uint8, device=device) for h, w in shapes]
        npieces = -(-(hi - lo) // self.piece)
        order, seen = [], set()
        for a, b in priority:
            a, b = max(a, lo), min(b, hi)
            if a >= b:
                continue
            for pc in range((a - lo) // self.piece, (b - 1 - lo) // self.piece + 1):
                if pc not in seen:
                    seen.add(pc)
                    order.append(pc)
        order += [pc for pc in range(npieces) if pc not in seen]
        events = [None] * npieces
        self.stream.wait_stream(torch.cuda.current_stream())
        self.ev_begin = torch.cuda.Event(enable_timing=True)
        self.ev_end = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            self.ev_begin.record(self.stream)
            for pc in order:
                a, b = lo + pc * self.piece, min(hi, lo + (pc + 1) * self.piece)
                for v in range(len(shapes)):
                    bufs[v][a - lo:b - lo].copy_(host_slice(v, a, b), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                events[pc] = ev
            self.ev_end.record(self.stream)
        self.bufs, self.events, self.home = bufs, events, (lo, hi)
        self.bytes += nbytes
        return True

    def upload_ms(self):
        """device time of the last upload (call after the pass has been synchronised)."""
        return self.ev_begin.elapsed_time(self.ev_end) if self.home is not None else 0.0

    def covers(self, a, b):
        return self.home is not None and a >= self.home[0] and b <= self.home[1]

    def get(self, a, b):
        lo = self.home[0]
        cur = torch.cuda.current_stream()
        for pc in range((a - lo) // self.piece, (b - 1 - lo) // self.piece + 1):
            cur.wait_event(self.events[pc])
        return [buf[a - lo:b - lo] for buf in self.bufs]

    def drop(self):
        self.bufs = self.events = self.home = None


class HostArraySource(FrameSource):
    """Views in host memory.  The arrays are page-locked in place (cudaHostRegister) so that every
    fetch is a direct DMA from the caller's buffer; if registration is refused the copy goes
    through a pinned staging buffer.  `prefetch` overlaps the next copy with the current kernels."""

    def __init__(self, views, block_frames=None, pin=True, cache_on_device=True):
        self.cache_on_device = cache_on_device      # False forces the streaming path (videos larger than HBM)
        self.views = []
        for v in views:
            t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
            if t.dtype != torch.uint8 or t.dim() != 3:
                raise ValueError("HostArraySource needs uint8 [T,H,W] views")
            self.views.append(t.contiguous())
        self.nframes = int(self.views[0].shape[0])
        if any(int(v.shape[0]) != self.nframes for v in self.views):
            raise ValueError("simultaneous views must have the same number of frames")
        self.Ly = [int(v.shape[1]) for v in self.views]
        self.Lx = [int(v.shape[2]) for v in self.views]
        self.block_frames = list(block_frames) if block_frames is not None else [self.nframes]
        self._registered = []
        self._pinned = [v.is_pinned() for v in self.views]
        self._stage = {}
        self._pin = pin
        self._pipe = _H2DPipe()
        self._cache = DeviceFrameCache()
        self.h2d = 0

    def plan_access(self, priority, home, device):
        self._ensure_pinned()
        self._cache.drop()
        if self.cache_on_device and all(self._pinned):
            for (x, y) in priority:      # sample ranges outside `home` (multi-rank) are copied first: the H2D
                x, y = self.clamp(x, y)  # engine serves copies in submission order
                if not (x >= home[0] and y <= home[1]):
                    self._pipe.issue((x, y), self._host_slices(x, y), device)
            before = self._cache.bytes
            self._cache.start(lambda v, a, b: self.views[v][a:b], list(zip(self.Ly, self.Lx)), priority, home, device)
            self.h2d += self._cache.bytes - before

    def _ensure_pinned(self):
        """page-lock the caller's buffers in place (once, lazily: needs a CUDA context)."""
        if not self._pin:
            return
        for i, v in enumerate(self.views):
            if self._pinned[i] or v.numel() == 0:
                continue
            ok = _lib.load().fm_host_register(v.data_ptr(), v.numel()) == 0
            if ok:
                self._registered.append(v.data_ptr())
                self._pinned[i] = True
        self._pin = False

    def h2d_bytes(self, nframes):
        return int(nframes) * sum(h * w for h, w in zip(self.Ly, self.Lx))

    def resident(self, t0, t1):
        return self._cache.covers(t0, t1)

    def _host_slices(self, a, b):
        out = []
        for i, v in enumerate(self.views):
            src = v[a:b]
            if not self._pinned[i]:
                st = torch.empty(tuple(src.shape), dtype=torch.uint8).pin_memory()   # staging (slow path)
                st.copy_(src)
                src = st
            out.append(src)
        return out

    def prefetch(self, t0, t1, device):
        self._ensure_pinned()
        a, b = self.clamp(t0, t1)
        if not self._cache.covers(a, b):
            self._pipe.issue((a, b), self._host_slices(a, b), device)

    def fetch(self, t0, t1, device):
        self._ensure_pinned()
        a, b = self.clamp(t0, t1)
        if self._cache.covers(a, b):
            return self._cache.get(a, b)
        if (a, b) not in self._pipe.pending:
            self._pipe.issue((a, b), self._host_slices(a, b), device)
        self.h2d += self.h2d_bytes(b - a)
        return self._pipe.take((a, b))

    def close(self):
        self._cache.drop()
        for ptr in self._registered:
            _lib.load().fm_host_unregister(ptr)
        self._registered = []
        self._pinned = [v.is_pinned() for v in self.views]
        self._pin = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DecodeCache:
    """Decode-once cache of a file source: every frame of `home` is decoded by `workers` threads with their own
    capture handles (OpenCV releases the GIL while decoding) into pinned staging buffers and copied into
    HBM-resident views on a private stream.  Readers block (host side) until the pieces they need are decoded,
    then wait (device side) on the copy event.

    The work is dealt in RUNS — stretches of consecutive `piece`-frame pieces that one thread decodes sequentially
    after ONE seek: first the runs the SVD sample ranges touch (the reference seeks to those as well,
    utils.get_frames, facemap/utils.py:536-540), then the rest of `home` as one long span per thread, in frame
    order.  The reference reads its stage-3 sweep sequentially, and a seek into an inter-frame codec can land a
    frame off, so every seek is checked: the position the capture reports after it, and the frame at each run
    start against the SAME frame decoded as the continuation of the run that ends there (one extra frame per run).
    `verify()` waits for all decoders and says whether every seam agreed; a source whose seeks are not
    frame-accurate sweeps through the sequential on-demand path instead (VideoFileSource.begin_sweep)."""

    def __init__(self, src, priority, home, device, piece=256, workers=8):
        import queue
        import threading

        self.src, self.device, self.piece = src, device, piece
        self.lo, self.hi = home
        n = self.hi - self.lo
        self.bufs = [torch.empty((n, h, w), dtype=torch.uint8, device=device) for h, w in zip(src.Ly, src.Lx)]
        self.npieces = -(-n // piece)
        nv = len(src.Ly)
        workers = max(1, workers)
        self.runs = runs = self.plan_runs(priority, home, piece, nv, workers)
        self.ready = [[threading.Event() for _ in range(nv)] for _ in range(self.npieces)]
        self.copied = [[None] * nv for _ in range(self.npieces)]
        self.error = None
        self.stop = False
        self.stream = torch.cuda.Stream(device)
        self.jobs = queue.Queue()
        for (p0, p1) in runs:
            for v in range(nv):
                self.jobs.put((p0, p1, v))
        self.decoded_frames = 0
        self.seam_frames = 0
        self.run_head = {}        # (view, frame) -> first frame decoded after the seek that starts a run
        self.run_tail = {}        # (view, frame) -> that frame decoded as the continuation of the run ending there
        self.seek_mismatch = False
        self._lock = threading.Lock()
        self._verdict = None
        self.threads = [threading.Thread(target=self._work, daemon=True) for _ in range(workers)]
        for t in self.threads:
            t.start()

    @staticmethod
    def plan_runs(priority, home, piece, nviews, workers):
        """[(first piece, end piece)] in the order they are dealt: the runs the sample ranges touch, in access
        order, then the rest of `home` cut into about `workers / nviews` long spans per view in frame order."""
        lo, hi = home
        npieces = -(-(hi - lo) // piece)
        order, seen = [], set()
        for a, b in priority:
            a, b = max(a, lo), min(b, hi)
            if a >= b:
                continue
            for pc in range((a - lo) // piece, (b - 1 - lo) // piece + 1):
                if pc not in seen:
                    seen.add(pc)
                    order.append(pc)
        runs = _DecodeCache._consecutive(order)
        rest = [pc for pc in range(npieces) if pc not in seen]
        span = max(1, -(-len(rest) * nviews // max(1, workers)))
        for p0, p1 in _DecodeCache._consecutive(rest):
            for q in range(p0, p1, span):
                runs.append((q, min(p1, q + span)))
        return runs

    @staticmethod
    def _consecutive(pieces):
        """[p, p+1, ..., q-1, r, r+1, ...] -> [(p, q), (r, ...)] keeping the order of first appearance"""
        out = []
        for pc in pieces:
            if out and out[-1][1] == pc:
                out[-1] = (out[-1][0], pc + 1)
            else:
                out.append((pc, pc + 1))
        return out

    def _work(self):
        import queue

        caps = None
        try:
            torch.cuda.set_device(self.device)
            caps = self.src._open()
            staging = {}
            while not self.stop:
                try:
                    p0, p1, v = self.jobs.get_nowait()
                except queue.Empty:
                    break
                if v not in staging:
                    staging[v] = torch.empty((self.piece, self.src.Ly[v], self.src.Lx[v]), dtype=torch.uint8,
                                             pin_memory=True)
                for pc in range(p0, p1):
                    if self.stop:
                        break
                    a = self.lo + pc * self.piece
                    b = min(self.hi, a + self.piece)
                    host = staging[v][:b - a]
                    seeks = []
                    self.src._read(v, a, b, out=host.numpy(), caps=caps, seeks=seeks)
                    if any(not ok for ok in seeks):
                        self.seek_mismatch = True
                    if pc == p0:
                        with self._lock:
                            self.run_head[(v, a)] = host[0].numpy().copy()
                    with torch.cuda.stream(self.stream):
                        self.bufs[v][a - self.lo:b - self.lo].copy_(host, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.stream)
                    self.copied[pc][v] = ev
                    self.ready[pc][v].set()
                    with self._lock:
                        self.decoded_frames += b - a
                    ev.synchronize()                      # the staging buffer is free again
                # the frame after the run, decoded as its continuation: what the next run's seek must reproduce
                end = min(self.hi, self.lo + p1 * self.piece)
                if end < self.hi and not self.stop:
                    tail = self.src._read(v, end, end + 1, caps=caps)
                    with self._lock:
                        self.run_tail[(v, end)] = tail[0]
                        self.seam_frames += 1
        except Exception as e:                        # surface in the reader instead of hanging it
            self.error = e
            for row in self.ready:
                for r in row:
                    r.set()
        finally:
            if caps is not None:
                self.src._release(caps)

    def verify(self):
        """Wait for every decoder thread; True when each seek landed on the frame asked for: the capture reported
        the requested position after every seek, and every run starts with the frame that the run before it
        decodes there sequentially."""
        if self._verdict is None:
            for t in self.threads:
                t.join()
            if self.error is not None:
                raise self.error
            ok = not self.seek_mismatch
            for key, head in self.run_head.items():
                tail = self.run_tail.get(key)
                if tail is not None and not np.array_equal(head, tail):
                    ok = False
            self._verdict = ok
        return self._verdict

    def covers(self, a, b):
        return a >= self.lo and b <= self.hi

    def get(self, a, b):
        cur = torch.cuda.current_stream()
        for pc in range((a - self.lo) // self.piece, (b - 1 - self.lo) // self.piece + 1):
            for v in range(len(self.bufs)):
                self.ready[pc][v].wait()
                if self.error is not None:
                    raise self.error
                cur.wait_event(self.copied[pc][v])
        return [buf[a - self.lo:b - self.lo] for buf in self.bufs]

    def close(self):
        self.stop = True
        for t in self.threads:
            t.join()
        self.bufs = None


class VideoFileSource(FrameSource):
    """filenames: 2-D list [[view0, view1, ...] per sequential block] (facemap/process.py:652-654).
    Decoding stays on the host (OpenCV); a decode failure repeats the previous frame once and leaves the rest
    of that file's range zero like utils.get_frames (facemap/utils.py:541-548; see `_read`).

    When the frames a pass needs fit in HBM (`plan_access`), every frame is decoded ONCE by a pool of decoder
    threads straight into HBM-resident views (`_DecodeCache`); the reference decodes each frame up to three
    times (stage-1 segments, stage-2 chunks, stage-3 sweep) on one thread.  Otherwise ranges are decoded on
    demand: one decode thread per view into pinned buffers, `prefetch` starting the next range while the GPU
    works on the current one, asynchronous H2D on a side stream."""

    def __init__(self, filenames, decode_workers=None, cache_on_device=True, mem_fraction=0.55):
        import cv2
        from concurrent.futures import ThreadPoolExecutor

        self._cv2 = cv2
        self.filenames = filenames
        self.caps = self._open()
        cum = [0]
        self.Ly, self.Lx = [], []
        for b, caps in enumerate(self.caps):
            if b == 0:
                for cap in caps:
                    self.Ly.append(int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
                    self.Lx.append(int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)))
            # utils.py:583-590: the frame count of a block is the one its LAST view reports
            cum.append(cum[-1] + int(caps[-1].get(cv2.CAP_PROP_FRAME_COUNT)))
        self.cumframes = np.array(cum).astype(int)
        self.nframes = int(cum[-1])
        self.block_frames = [int(x) for x in np.diff(self.cumframes)]
        self._workers = [ThreadPoolExecutor(max_workers=1) for _ in self.Ly]   # on-demand path: one decoder per view
        self._jobs = {}
        self._pipe = _H2DPipe()
        self.h2d = 0
        import os

        # decoder threads of the decode-once cache: the host cores this process may use (its affinity mask),
        # shared between the ranks of one box, two left for the main and copy threads
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 4
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        self.decode_workers = decode_workers if decode_workers is not None else max(2, min(32, (cores - 2) // ranks))
        self.cache_on_device = cache_on_device
        self.mem_fraction = mem_fraction
        self._cache = None
        self.seek_inexact = False

    def _open(self):
        """a private set of capture handles [block][view] (a handle is not thread safe)"""
        out = []
        for block in self.filenames:
            caps = []
            for f in block:
                cap = self._cv2.VideoCapture(f)
                if not cap.isOpened():
                    raise IOError(f"cannot open video {f}")
                caps.append(cap)
            out.append(caps)
        return out

    @staticmethod
    def _release(caps):
        for block in caps:
            for c in block:
                c.release()

    def h2d_bytes(self, nframes):
        return int(nframes) * sum(h * w for h, w in zip(self.Ly, self.Lx))

    def plan_access(self, priority, home, device):
        if self._cache is not None:
            self._cache.close()
            self._cache = None
        lo, hi = home
        nbytes = self.h2d_bytes(hi - lo)
        if not self.cache_on_device or hi <= lo:
            return
        free, _ = torch.cuda.mem_get_info(device)
        if nbytes > self.mem_fraction * free:
            return
        self._cache = _DecodeCache(self, priority, (lo, hi), device, workers=self.decode_workers)
        self.h2d += nbytes

    def _read(self, view, a, b, out=None, caps=None, seeks=None):
        """frames [a, b) of one view (utils.get_frames, facemap/utils.py:522-553).  `seeks`: optional list that
        receives, for every seek made, whether the capture then reported the requested position."""
        cv2 = self._cv2
        caps = self.caps if caps is None else caps
        if out is None:
            out = np.zeros((b - a, self.Ly[view], self.Lx[view]), np.uint8)
        k = 0
        for blk in range(len(caps)):
            lo, hi = max(a, self.cumframes[blk]), min(b, self.cumframes[blk + 1])
            if lo >= hi:
                continue
            cap = caps[blk][view]
            first = int(lo - self.cumframes[blk])
            if int(cap.get(cv2.CAP_PROP_POS_FRAMES)) != first:
                cap.set(cv2.CAP_PROP_POS_FRAMES, first)
                if seeks is not None:
                    seeks.append(int(cap.get(cv2.CAP_PROP_POS_FRAMES)) == first)
            end = k + (hi - lo)
            while k < end:
                ok, frame = cap.read()
                if not ok:
                    # utils.py:541-548: the frame that fails to decode repeats its predecessor (a zero frame when
                    # it is the first of the request), reading of this file then stops and the rest of its range
                    # keeps the zeros of the freshly allocated chunk buffer (process.py:46-50)
                    print("img load failed, replacing with prev..")
                    out[k] = out[k - 1] if k > 0 else 0
                    out[k + 1:end] = 0
                    k = end
                    break
                out[k] = frame if frame.ndim == 2 else cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
                k += 1
        return out

    def begin_sweep(self):
        """Called before the sequential sweep over all frames (stage 3, facemap/process.py:394-515): the reference
        decodes that sweep without seeking, so the decode-once cache serves it only if every seek it made proved
        frame-accurate (_DecodeCache.verify); otherwise the sweep decodes on demand, sequentially."""
        if self._cache is not None and not self._cache.verify():
            print("facemap_b200: seeks into this video are not frame-accurate; stage 3 decodes sequentially")
            self.seek_inexact = True
            self._cache.close()
            self._cache = None

    def _decode_pinned(self, view, a, b):
        buf = torch.empty((b - a, self.Ly[view], self.Lx[view]), dtype=torch.uint8, pin_memory=True)
        self._read(view, a, b, out=buf.numpy())
        return buf

    def _submit(self, a, b):
        if (a, b) not in self._jobs:
            self._jobs[(a, b)] = [w.submit(self._decode_pinned, v, a, b) for v, w in enumerate(self._workers)]

    def prefetch(self, t0, t1, device):
        a, b = self.clamp(t0, t1)
        if self._cache is not None and self._cache.covers(a, b):
            return
        self._submit(a, b)

    def fetch(self, t0, t1, device):
        a, b = self.clamp(t0, t1)
        if self._cache is not None and self._cache.covers(a, b):
            return self._cache.get(a, b)
        self._submit(a, b)
        host = [f.result() for f in self._jobs.pop((a, b))]
        self.h2d += self.h2d_bytes(b - a)
        self._pipe.issue((a, b), host, device)
        return self._pipe.take((a, b))

    def close(self):
        if self._cache is not None:
            self._cache.close()
            self._cache = None
        for w in self._workers:
            w.shutdown(wait=True)
        self._release(self.caps)


def as_source(obj):
    """filenames (2-D list of paths) | FrameSource | list of arrays/tensors -> FrameSource."""
    if isinstance(obj, FrameSource):
        return obj
    if isinstance(obj, (list, tuple)) and len(obj) and isinstance(obj[0], (list, tuple)) \
            and len(obj[0]) and isinstance(obj[0][0], str):
        return VideoFileSource(obj)
    if isinstance(obj, (list, tuple)) and len(obj) and (torch.is_tensor(obj[0]) or isinstance(obj[0], np.ndarray)):
        if torch.is_tensor(obj[0]) and obj[0].is_cuda:
            return DeviceArraySource(obj)
        return HostArraySource(obj)
    raise TypeError("expected a 2-D list of video paths, a FrameSource, or a list of uint8 [T,H,W] views")
