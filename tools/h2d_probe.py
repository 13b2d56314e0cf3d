"""Host<->device copy rate of this box for the sizes the per-cycle upload moves (pinned memory, one DMA each)."""
import sys, json, os
import torch
sys.path.insert(0, ".")
import bench
res = {"numa": bench.pin_to_gpu_numa(0)}
torch.cuda.init()
for mb in (0.25, 1.0, 3.5, 16.0, 64.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for direction in ("h2d", "d2h"):
        best = 1e9
        for _ in range(20):
            a.record()
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            b.record(); b.synchronize()
            best = min(best, a.elapsed_time(b))
        res[f"{direction}_{mb}MB"] = {"ms": round(best, 4), "GBps": round(n / best / 1e6, 2)}
print(json.dumps(res))
