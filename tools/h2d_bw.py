#!/usr/bin/env python3
"""Raw pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e leg: 608 B per record over PCIe)."""
import torch
n = 640 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 0.0
for _ in range(5):
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print("pinned H2D, 640 MiB: %.1f GB/s -> at 608 B per record at most %.3e verifies/s end to end" % (best, best * 1e9 / 608))
