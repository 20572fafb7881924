"""Pinned-memory copy bandwidth of this box's PCIe link at the bench's transfer sizes (D2H and H2D)."""
import torch

dev = torch.device("cuda:0")
for mb in (1, 8.4, 64, 256):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, (dst, src) in (("D2H", (h, d)), ("H2D", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{name} {mb:6.1f} MB: {us:8.1f} us  {n / us / 1e3:6.1f} GB/s")
