"""Multi-GPU check (torchrun, one rank per GPU): the sharded path against the single-GPU path on the
same video.  Integer-stage outputs must be bit-identical; SVD outputs equal to float32 round-off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from facemap_b200 import process
from facemap_b200.frames import DeviceArraySource
from facemap_b200.parallel import Shard
from facemap_b200.synth import SynthVideo
from tests.helpers import motion_roi, subspace_sin

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sh = Shard()
ok = True
for (H, W, N, sbin, mov, rois) in [(96, 128, 5300, 2, False, None), (64, 96, 3100, 4, True, [motion_roi(0, 8, 56, 9, 80)])]:
    v = torch.from_numpy(SynthVideo(H, W, N, seed=3).all_frames()).to(dev)
    import copy
    pd = process.process_source(DeviceArraySource([v]), sbin=sbin, motSVD=True, movSVD=mov, rois=copy.deepcopy(rois), dist=sh)
    if rank == 0:
        ps = process.process_source(DeviceArraySource([v]), sbin=sbin, motSVD=True, movSVD=mov, rois=copy.deepcopy(rois))
        for key in ("avgframe", "avgmotion", "motion"):
            for a, b in zip(pd[key], ps[key]):
                same = np.array_equal(a, b); ok &= same
                print(f"N={N} {key}: bit-identical {same}")
        for sk, mk, vk in (("motSv", "motMask", "motSVD"), ("movSv", "movMask", "movSVD")):
            if not len(ps[mk]):
                continue
            rel = float(np.abs(pd[sk] - ps[sk]).max() / ps[sk].max())
            ang = max(subspace_sin(a[:, :10], b[:, :10]) for a, b in zip(pd[mk], ps[mk]))
            tr = max(float(np.abs(np.abs(a[:, :5]) - np.abs(b[:, :5])).max() / np.abs(b[:, :5]).max()) for a, b in zip(pd[vk], ps[vk]))
            good = rel < 1e-5 and ang < 1e-4 and tr < 1e-4
            ok &= good
            print(f"N={N} {sk}: rel {rel:.2e} angle(10) {ang:.2e} trace {tr:.2e} -> {'ok' if good else 'FAIL'}")
    dist.barrier()
# non-power-of-two bins, two cameras: stage-1 means travel as float64 (fm_mean_ref), the motion accumulator uses
# the reference-arithmetic sums (fm_motion_ref) gathered over the shards -> bit-identical to one GPU
va = torch.from_numpy(SynthVideo(90, 120, 2600, seed=5).all_frames()).to(dev)
vb = torch.from_numpy(SynthVideo(72, 96, 2600, seed=6).all_frames()).to(dev)
pd = process.process_source(DeviceArraySource([va, vb]), sbin=3, motSVD=True, movSVD=False, dist=sh)
if rank == 0:
    ps = process.process_source(DeviceArraySource([va, vb]), sbin=3, motSVD=True, movSVD=False)
    for key in ("avgframe", "avgmotion", "motion"):
        same = all(np.array_equal(a, b) for a, b in zip(pd[key], ps[key])); ok &= same
        print(f"sbin 3, two views, {key}: bit-identical {same}")
dist.barrier()
# full-resolution ROI processors: sharded by frame range (running's predecessor comes from the one-frame halo),
# gathered; pupil / blink / running and the painted motion energy must be bit-identical to the single-GPU run
from tests.helpers import fullres_case
views, sbin, mot, mov, rois, fullSVD = fullres_case("eye_all")
vs = [torch.from_numpy(a).to(dev) for a in views]
pd = process.process_source(DeviceArraySource(vs), sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois), dist=sh)
if rank == 0:
    ps = process.process_source(DeviceArraySource(vs), sbin=sbin, motSVD=mot, movSVD=mov, rois=copy.deepcopy(rois))
    checks = [("blink", pd["blink"], ps["blink"]), ("running", pd["running"], ps["running"]), ("motion", pd["motion"], ps["motion"])]
    for key in ("com", "area", "axdir", "axlen", "area_smooth", "com_smooth"):
        checks.append(("pupil." + key, [p[key] for p in pd["pupil"]], [p[key] for p in ps["pupil"]]))
    for name, a, b in checks:
        same = len(a) == len(b) and all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
        ok &= same
        print(f"rois {name}: bit-identical {same}")
dist.barrier()
if rank == 0:
    print("DIST CHECK", "PASSED" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
