"""Deterministic synthetic uint8 behaviour video.

The generator is integer-only so that the NumPy (host) and torch (device) variants
produce *identical bytes*: every intermediate is an exact integer (held in int64 or
in float64 below 2^53), so summation order cannot matter.

    f[t,y,x] = clip( B[y,x] + floor( sum_k c_k[t] * m_k[y,x] / 256 ) + noise(seed,t,y,x), 0, 255 )

* ``B``     smooth background in [60, 200]
* ``m_k``   R Gaussian-blob spatial modes, integers in [0, 255]
* ``c_k``   integer AR(1) traces (rho = 230/256), amplitude ~ (k+1)^-0.7
* ``noise`` 2-bit hash noise in {-2,-1,0,1}

Nothing here is on the timed hot path; it only makes inputs.
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


class SynthVideo:
    """One camera view of a synthetic video; frames are produced in [t0, t1) ranges."""

    def __init__(self, H: int, W: int, nframes: int, seed: int = 0, nmodes: int = 64,
                 noise_bits: int = 2):
        self.H, self.W, self.nframes = int(H), int(W), int(nframes)
        self.seed, self.R, self.noise_bits = int(seed), int(nmodes), int(noise_bits)
        rng = np.random.RandomState(1000 + self.seed)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
        # background: low-order smooth surface in [60, 200]
        bg = 130.0 + 35.0 * np.sin(2 * np.pi * (yy / H) * 0.7 + 0.3 * self.seed) \
            + 35.0 * np.cos(2 * np.pi * (xx / W) * 0.9 + 0.1)
        self.background = np.clip(np.rint(bg), 60, 200).astype(np.int64)  # (H, W)
        # spatial modes
        L = float(min(H, W))
        modes = np.empty((self.R, H * W), np.float64)
        for k in range(self.R):
            cy, cx = rng.uniform(0, H), rng.uniform(0, W)
            sig = rng.uniform(0.05, 0.25) * L
            g = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig))
            modes[k] = np.rint(255.0 * g).reshape(-1)
        self.modes = modes  # exact small integers stored in float64, (R, H*W)
        # integer AR(1) traces, shape (nframes, R)
        amp = np.maximum(2, np.rint(160.0 * (np.arange(self.R) + 1.0) ** -0.7)).astype(np.int64)
        drive = rng.randint(-64, 65, size=(self.nframes, self.R)).astype(np.int64)
        c = np.zeros((self.nframes, self.R), np.int64)
        prev = np.zeros(self.R, np.int64)
        for t in range(self.nframes):
            prev = ((230 * prev) >> 8) + drive[t]
            c[t] = prev
        # scale: c*amp/64, kept integer
        self.traces = ((c * amp) >> 6).astype(np.float64)  # (nframes, R), |.| < ~2^10

    # ---- hash noise -----------------------------------------------------------------------
    def _noise_np(self, t0: int, t1: int) -> np.ndarray:
        T, P = t1 - t0, self.H * self.W
        t = np.arange(t0, t1, dtype=np.int64)[:, None]
        p = np.arange(P, dtype=np.int64)[None, :]
        h = (t * 0x1E3779B1 + p * 0x05EBCA77 + (self.seed + 1) * 0x42B2AE3D) & _M32
        h ^= h >> 15
        h = (h * 0x2C1B3C6D) & _M32
        h ^= h >> 12
        h = (h * 0x297A2D39) & _M32
        h ^= h >> 15
        nb = self.noise_bits
        return ((h & ((1 << nb) - 1)) - (1 << (nb - 1))) if nb > 0 else np.zeros((T, P), np.int64)

    def frames(self, t0: int, t1: int) -> np.ndarray:
        """uint8 array (t1-t0, H, W), host / NumPy."""
        t0, t1 = int(t0), int(t1)
        s = self.traces[t0:t1] @ self.modes                # exact integers in float64
        s = np.floor(s / 256.0).astype(np.int64)
        s += self.background.reshape(1, -1)
        s += self._noise_np(t0, t1)
        return np.clip(s, 0, 255).astype(np.uint8).reshape(t1 - t0, self.H, self.W)

    def all_frames(self, chunk: int = 2000) -> np.ndarray:
        out = np.empty((self.nframes, self.H, self.W), np.uint8)
        for t0 in range(0, self.nframes, chunk):
            t1 = min(self.nframes, t0 + chunk)
            out[t0:t1] = self.frames(t0, t1)
        return out

    # ---- device variant (torch; plumbing only) -------------------------------------------------
    def frames_torch(self, t0: int, t1: int, device="cuda"):
        """Same bytes as :meth:`frames`, produced on ``device`` as a torch.uint8 tensor."""
        import torch

        t0, t1 = int(t0), int(t1)
        if not hasattr(self, "_dev") or self._dev[0] != str(device):
            self._dev = (str(device),
                         torch.from_numpy(self.modes).to(device),
                         torch.from_numpy(self.traces).to(device),
                         torch.from_numpy(self.background.reshape(1, -1)).to(device))
        _, modes, traces, bg = self._dev
        s = traces[t0:t1] @ modes                          # float64 matmul, exact integers
        s = torch.floor(s / 256.0).to(torch.int64)
        s += bg
        T, P = t1 - t0, self.H * self.W
        t = torch.arange(t0, t1, dtype=torch.int64, device=device)[:, None]
        p = torch.arange(P, dtype=torch.int64, device=device)[None, :]
        h = (t * 0x1E3779B1 + p * 0x05EBCA77 + (self.seed + 1) * 0x42B2AE3D) & _M32
        h = h ^ (h >> 15)
        h = (h * 0x2C1B3C6D) & _M32
        h = h ^ (h >> 12)
        h = (h * 0x297A2D39) & _M32
        h = h ^ (h >> 15)
        nb = self.noise_bits
        if nb > 0:
            s += (h & ((1 << nb) - 1)) - (1 << (nb - 1))
        return torch.clamp(s, 0, 255).to(torch.uint8).reshape(T, self.H, self.W)

    def all_frames_torch(self, device="cuda", chunk: int = 1000):
        import torch

        out = torch.empty((self.nframes, self.H, self.W), dtype=torch.uint8, device=device)
        for t0 in range(0, self.nframes, chunk):
            t1 = min(self.nframes, t0 + chunk)
            out[t0:t1] = self.frames_torch(t0, t1, device)
        return out
