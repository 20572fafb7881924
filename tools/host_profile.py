"""cProfile of the host side of graph-mode Environment.step (tiny batch: the GPU never limits)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import vectorizedmultiagentsimulator_b200 as b200

name = sys.argv[1] if len(sys.argv) > 1 else "balance"
kwargs = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[2:])}
env = b200.make_env(name, num_envs=32, device="cuda", seed=0, cuda_graph=True, **kwargs)
env.reset()
acts = [env.get_random_actions() for _ in range(16)]
for i in range(50):
    env.step(acts[i % 16])
torch.cuda.synchronize()
import time
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(2000):
        env.step(acts[i % 16])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"# {name}: host wall per step (32 envs, no profiler): {(t1 - t0) / 2000 * 1e6:.1f} us")
prof = cProfile.Profile()
prof.enable()
for i in range(2000):
    env.step(acts[i % 16])
prof.disable()
torch.cuda.synchronize()
st = pstats.Stats(prof)
st.sort_stats("tottime").print_stats(28)
