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

# ---- a step's results (observations 8.4 MB, rewards 0.5 MB, dones 32 KB) to pinned host memory: three copy-engine
# copies against ONE launch of the library's SM copy kernel storing straight into the (mapped) pinned buffers
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectorizedmultiagentsimulator_b200 import _native as N

lib = N.load()
shapes = [((4, 32768, 16), torch.float32), ((4, 32768), torch.float32), ((32768,), torch.bool)]
srcs = [torch.zeros(s, dtype=t, device=dev) for s, t in shapes]
dsts = [torch.empty(s, dtype=t).pin_memory() for s, t in shapes]
total = sum(x.numel() * x.element_size() for x in srcs)


def ce():
    for d, s in zip(dsts, srcs):
        d.copy_(s, non_blocking=True)


def sm():
    N.copy_buffers(lib, dev, list(zip(srcs, dsts)))


for name, fn in (("3 copy-engine copies", ce), ("1 SM copy kernel", sm)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"results of one step ({total / 1e6:.1f} MB) -> host, {name}: {us:7.1f} us  {total / us / 1e3:6.1f} GB/s")
