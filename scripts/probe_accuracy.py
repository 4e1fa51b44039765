"""GPU probe: accuracy of the whole path against the live exact oracle (O2) on EVERY component — singular values,
mask subspace per prefix, projected traces per component — for a few configurations and Gram / projection modes.
Writes gpurun_out/accuracy_<tag>.json and prints a table (the source of profiles/r02_accuracy.md).

    python scripts/probe_accuracy.py [tag]          # modes from FM_GRAM_MODE / FM_PROJ_MODE (defaults of the library)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from facemap_b200 import process
from facemap_b200.frames import DeviceArraySource
from facemap_b200.synth import SynthVideo
from oracle import facemap_oracle as orc
from tests.helpers import component_errors, motion_roi

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
dev = torch.device("cuda", 0)
CASES = {
    # name: (views [(H, W, seed)], N, sbin, rois)
    "96x128x2600_s2": ([(96, 128, 5)], 2600, 2, None),
    "two_views_roi_s4": ([(96, 128, 0), (80, 112, 1)], 2200, 4, [motion_roi(0, 10, 70, 17, 90), motion_roi(1, 8, 60, 20, 100)]),
    "160x192x4100_s2": ([(160, 192, 7)], 4100, 2, None),
}
report = {}
for name, (views, N, sbin, rois) in CASES.items():
    import copy

    vids = [SynthVideo(H, W, N, seed=s).all_frames() for (H, W, s) in views]
    t = time.time()
    o = orc.run_oracle(vids, sbin=sbin, motSVD=True, movSVD=True, rois=copy.deepcopy(rois), svd="exact")
    t_or = time.time() - t
    p = process.process_source(DeviceArraySource([torch.from_numpy(v).to(dev) for v in vids]), sbin=sbin, motSVD=True,
                               movSVD=True, rois=copy.deepcopy(rois))
    print(f"== {name}: oracle {t_or:.1f} s; integer stages exact:",
          all(np.array_equal(a, b) for a, b in zip(p["avgmotion"], o["avgmotion"])),
          all(np.array_equal(a, b) for a, b in zip(p["motion"], o["motion"])))
    for mk, vk, sk in (("motMask", "motSVD", "motSv"), ("movMask", "movSVD", "movSv")):
        for i in range(len(p[mk])):
            U, Ur = p[mk][i], o[mk][i]
            if U.shape[0] == 0:
                continue
            last = i == len(p[mk]) - 1
            S, Sr = (p[sk].astype(np.float64), o[sk].astype(np.float64)) if last else (None, None)
            e = component_errors(U, Ur, p[vk][i], o[vk][i], S, Sr)
            key = f"{name}/{mk}[{i}]"
            report[key] = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in e.items()}
            marks = [k for k in (10, 50, 100, 150, 200, 250, 300, 400, 499) if k < U.shape[1]]
            print(key, "shape", U.shape)
            if S is not None:
                print("   S_k/S_0      ", " ".join(f"{Sr[k] / Sr[0]:9.1e}" for k in marks))
                print("   S rel err <=k", " ".join(f"{e['S_rel'][:k + 1].max():9.1e}" for k in marks))
            print("   prefix angle ", " ".join(f"{e['prefix_sin'][k]:9.1e}" for k in marks))
            print("   max angle <=k", " ".join(f"{max(e['prefix_sin'][:k + 1]):9.1e}" for k in marks))
            print("   trace rel <=k", " ".join(f"{e['trace_rel'][:k + 1].max():9.1e}" for k in marks))
            print("   1-|cos| <=k  ", " ".join(f"{e['one_minus_cos'][:k + 1].max():9.1e}" for k in marks))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open(f"gpurun_out/accuracy_{tag}.json", "w"))
