"""Per-step host time and CUDA-event time of graph-mode Environment.step, to locate outliers."""
import gc
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import vectorizedmultiagentsimulator_b200 as b200


def run(tag, flush_on=True, gc_off=False, steps=300, drop=True):
    env = b200.make_env("balance", num_envs=32768, device="cuda", seed=0, cuda_graph=True, n_agents=4)
    env.reset()
    acts = [env.get_random_actions() for _ in range(16)]
    for i in range(10):
        env.step(acts[i % 16])
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    if gc_off:
        gc.collect()
        gc.disable()
    evs, host, keep = [], [], []
    for i in range(steps):
        if flush_on:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        out = env.step(acts[i % 16])
        host.append((time.perf_counter() - t0) * 1e6)
        e1.record()
        evs.append((e0, e1))
        if not drop:
            keep.append(out)
            keep = keep[-4:]
    torch.cuda.synchronize()
    if gc_off:
        gc.enable()
    dev = [a.elapsed_time(b) * 1e3 for a, b in evs]
    order = sorted(range(steps), key=lambda i: -dev[i])[:8]
    med = sorted(dev)[steps // 2]
    hmed = sorted(host)[steps // 2]
    print(f"[{tag}] device median {med:.1f} us mean {sum(dev) / steps:.1f} us | host median {hmed:.1f} us mean {sum(host) / steps:.1f} us")
    print("   worst device brackets (step: device us / host us):", ", ".join(f"{i}: {dev[i]:.0f}/{host[i]:.0f}" for i in order))
    worst_host = sorted(range(steps), key=lambda i: -host[i])[:8]
    print("   worst host times (step: host us / device us):", ", ".join(f"{i}: {host[i]:.0f}/{dev[i]:.0f}" for i in worst_host))


run("flush, gc on")
run("flush, gc off", gc_off=True)
run("no flush, gc on", flush_on=False)
run("flush, outputs kept alive", drop=False)
