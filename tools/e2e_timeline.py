"""Device timeline of the pipelined end-to-end loop of bench.py (torch.profiler): which stream runs
what, when — to see whether the download of step t-1 overlaps step t.  Not a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import ProfilerActivity, profile

import vectorizedmultiagentsimulator_b200 as b200

B = 32768
env = b200.make_env("balance", num_envs=B, device="cuda", seed=0, cuda_graph=True, n_agents=4)
gen = torch.Generator().manual_seed(0)
host_actions = [[(torch.rand(B, 2, generator=gen) * 2 - 1).pin_memory() for _ in env.agents] for _ in range(12)]
obs0, rew0, done0, _ = env.step([a.cuda() for a in host_actions[0]])
host_sets = [
    (torch.empty((4,) + tuple(obs0[0].shape)).pin_memory(), torch.empty((4,) + tuple(rew0[0].shape)).pin_memory(),
     torch.empty(done0.shape, dtype=done0.dtype).pin_memory())
    for _ in range(2)
]
copy_stream = torch.cuda.Stream()
pending = [None]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def step(i):
    main = torch.cuda.current_stream()
    if pending[0] is not None:
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(host_sets[i & 1], pending[0]):
                dst.copy_(src, non_blocking=True)
    obs, rews, dones, _ = env.step(host_actions[i])
    fresh = (torch.stack(obs), torch.stack(rews), dones)
    main.wait_stream(copy_stream)
    pending[0] = fresh


for i in range(6):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6, 10):
        flush.zero_()
        step(i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    stream = getattr(e, "device_resource_id", None)
    print(f"{e.time_range.start - t0:10.1f} us +{e.time_range.elapsed_us():8.1f} us  stream {stream}  {e.name[:90]}")
