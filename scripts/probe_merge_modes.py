"""GPU probe: the whole path with the merge eigensolve on cuSOLVER (syevdx) and by block subspace iteration
(FM_MERGE_EIG), motion AND movie SVD of a 16 000-frame 640x480 video (15 chunks, 3 750 stacked components per
modality): every singular value and the leading masks of one against the other, and the time of the merge solves."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facemap_b200 import process                     # noqa: E402
from facemap_b200.frames import DeviceArraySource    # noqa: E402
from facemap_b200.synth import SynthVideo            # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
frames = SynthVideo(480, 640, 16000, seed=3).all_frames_torch(dev)
res = {}
for mode in ("cusolver", "subspace", "subspace"):
    os.environ["FM_MERGE_EIG"] = mode
    tm = {}
    out = process.process_source(DeviceArraySource([frames]), sbin=4, motSVD=True, movSVD=True, timing=tm)
    res[mode] = (out, {k: round(v, 2) for k, v in tm["device_ms"].items() if k.endswith(".syevd")})
ref, sub = res["cusolver"][0], res["subspace"][0]
rep = {"merge_syevd_ms": {m: res[m][1] for m in res}}
for key in ("motSv", "movSv"):
    a, b = ref[key].astype(np.float64), sub[key].astype(np.float64)
    rep[key + "_max_rel_diff_all_500"] = float(np.max(np.abs(a - b) / a))
    rep[key + "_S499_over_S0"] = float(a[-1] / a[0])
for key in ("motMask", "movMask"):
    U, V = ref[key][0].astype(np.float64), sub[key][0].astype(np.float64)
    c = np.abs(np.sum(U[:, :40] * V[:, :40], axis=0)) / (np.linalg.norm(U[:, :40], axis=0) * np.linalg.norm(V[:, :40], axis=0))
    rep[key + "_worst_angle_first_40"] = float(np.sqrt(np.max(1 - np.minimum(c, 1.0) ** 2)))
    rep[key + "_orthogonality_subspace"] = float(np.abs(V.T @ V - np.eye(V.shape[1])).max())
for key in ("motSVD", "movSVD"):
    a, b = ref[key][0][:, :40].astype(np.float64), sub[key][0][:, :40].astype(np.float64)
    rep[key + "_trace_rel_diff_first_40"] = float(np.max(np.abs(np.abs(a) - np.abs(b))) / np.max(np.abs(a)))
print(json.dumps(rep, indent=1))
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "probe_merge_modes.json"), "w"), indent=1)
