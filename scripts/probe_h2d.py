"""GPU probe: pinned host -> device copy bandwidth (the ceiling of the e2e number)."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (64, 315, 1024, 4096):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"H2D {mb} MiB: {mb / 1024 / dt * 1.073741824:.1f} GB/s")
    for _ in range(2): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"D2H {mb} MiB: {mb / 1024 / dt * 1.073741824:.1f} GB/s")
