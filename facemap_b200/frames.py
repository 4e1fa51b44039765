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


class HostArraySource(FrameSource):
    """Views in host memory.  Each fetch copies through a pinned staging buffer (allocated once
    per view at the largest request seen) with a non-blocking H2D copy on the current stream."""

    def __init__(self, views, block_frames=None, pin=True):
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
        self._pinned = all(v.is_pinned() for v in self.views)
        self._stage = [None] * len(self.views)
        self._pin = pin

    def h2d_bytes(self, nframes):
        return int(nframes) * sum(h * w for h, w in zip(self.Ly, self.Lx))

    def fetch(self, t0, t1, device):
        a, b = self.clamp(t0, t1)
        out = []
        for i, v in enumerate(self.views):
            src = v[a:b]
            if not self._pinned and self._pin:
                st = self._stage[i]
                if st is None or st.shape[0] < b - a:
                    st = torch.empty((b - a,) + tuple(v.shape[1:]), dtype=torch.uint8).pin_memory()
                    self._stage[i] = st
                else:
                    # the previous copy out of this staging buffer must have drained
                    torch.cuda.current_stream().synchronize()
                st[:b - a].copy_(src)
                src = st[:b - a]
            out.append(src.to(device, non_blocking=True))
        return out


class VideoFileSource(FrameSource):
    """filenames: 2-D list [[view0, view1, ...] per sequential block] (facemap/process.py:652-654).
    Decoding stays on the host (OpenCV); decode failures repeat the previous frame like
    utils.get_frames (facemap/utils.py:545-547)."""

    def __init__(self, filenames):
        import cv2

        self._cv2 = cv2
        self.filenames = filenames
        self.caps = []
        cum = [0]
        self.Ly, self.Lx = [], []
        for b, block in enumerate(filenames):
            caps = []
            for f in block:
                cap = cv2.VideoCapture(f)
                if not cap.isOpened():
                    raise IOError(f"cannot open video {f}")
                caps.append(cap)
                if b == 0:
                    self.Ly.append(int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
                    self.Lx.append(int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)))
            self.caps.append(caps)
            cum.append(cum[-1] + int(caps[0].get(cv2.CAP_PROP_FRAME_COUNT)))
        self.cumframes = np.array(cum).astype(int)
        self.nframes = int(cum[-1])
        self.block_frames = [int(x) for x in np.diff(self.cumframes)]

    def h2d_bytes(self, nframes):
        return int(nframes) * sum(h * w for h, w in zip(self.Ly, self.Lx))

    def _read(self, view, a, b):
        cv2 = self._cv2
        out = np.zeros((b - a, self.Ly[view], self.Lx[view]), np.uint8)
        k = 0
        for blk in range(len(self.caps)):
            lo, hi = max(a, self.cumframes[blk]), min(b, self.cumframes[blk + 1])
            if lo >= hi:
                continue
            cap = self.caps[blk][view]
            first = int(lo - self.cumframes[blk])
            if int(cap.get(cv2.CAP_PROP_POS_FRAMES)) != first:
                cap.set(cv2.CAP_PROP_POS_FRAMES, first)
            for _ in range(hi - lo):
                ok, frame = cap.read()
                if ok:
                    out[k] = frame if frame.ndim == 2 else cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
                elif k > 0:
                    out[k] = out[k - 1]
                k += 1
        return out

    def fetch(self, t0, t1, device):
        a, b = self.clamp(t0, t1)
        return [torch.from_numpy(self._read(v, a, b)).pin_memory().to(device, non_blocking=True)
                for v in range(len(self.Ly))]

    def close(self):
        for caps in self.caps:
            for c in caps:
                c.release()


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
