"""Per-kernel timeline of ONE Environment.step (CUDA-graph replay or eager) from torch.profiler.

    python tools/step_timeline.py balance 32768 n_agents=4 [--eager]

Prints kernel name, start offset within the step and duration, plus the step's span and the sum of
kernel times (the difference is launch gaps).  Not a bench value (CUPTI is attached).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import ProfilerActivity, profile

import vectorizedmultiagentsimulator_b200 as b200


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    eager = "--eager" in sys.argv
    name, n_envs = args[0], int(args[1])
    kwargs = {k: int(v) for k, v in (kv.split("=") for kv in args[2:])}
    env = b200.make_env(name, num_envs=n_envs, device="cuda", seed=0, cuda_graph=not eager, **kwargs)
    env.reset()
    acts = [env.get_random_actions() for _ in range(8)]
    for a in acts[:5]:
        env.step(a)
    torch.cuda.synchronize()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for a in acts[5:]:
            flush.zero_()
            torch.cuda.synchronize()
            env.step(a)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at the flush memsets (the 512 MiB fill is by far the longest event)
    steps, cur = [], []
    for e in evs:
        if e.time_range.elapsed_us() > 40 and ("fill" in e.name.lower() or "memset" in e.name.lower()):
            if cur:
                steps.append(cur)
            cur = []
        else:
            cur.append(e)
    if cur:
        steps.append(cur)
    last = steps[-1]
    t0 = last[0].time_range.start
    total = 0.0
    print(f"# {name} {n_envs} envs {kwargs} {'eager' if eager else 'graph'}: kernels of one step")
    for e in last:
        d = e.time_range.elapsed_us()
        total += d
        print(f"{e.time_range.start - t0:9.1f} us  +{d:7.1f} us  {e.name[:110]}")
    span = last[-1].time_range.end - t0
    print(f"# {len(last)} device activities, span {span:.1f} us, sum of durations {total:.1f} us, gaps {span - total:.1f} us")


main()
