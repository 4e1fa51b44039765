"""GPU probe: pinned host -> device copy bandwidth, one rank per GPU all copying AT THE SAME TIME (torchrun) — the ceiling
of the e2e number at N GPUs: the ranks of one box share the host's memory and PCIe root complexes."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for mb in (1024, 4096):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    out = {}
    for name, a, b in (("H2D", h, d), ("D2H", d, h)):
        for _ in range(2):
            b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t) / 5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[name] = (mb << 20) / float(dt.item()) / 1e9
    if rank == 0:
        print(f"{world} rank(s) copying at once, {mb} MiB each: H2D {out['H2D']:.1f} GB/s per GPU ({world * out['H2D']:.0f} GB/s "
              f"aggregate), D2H {out['D2H']:.1f} GB/s per GPU ({world * out['D2H']:.0f} aggregate)", flush=True)
    del h, d
if world > 1:
    dist.destroy_process_group()
