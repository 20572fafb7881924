"""Steps an environment eagerly a few times (target process for ncu captures).

    ncu --set full -k regex:cast_rays_batched -s 4 -c 1 python tools/run_steps.py flocking 32768 n_agents=5
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import vectorizedmultiagentsimulator_b200 as b200

name, n_envs = sys.argv[1], int(sys.argv[2])
kwargs = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
env = b200.make_env(name, num_envs=n_envs, device="cuda", seed=0, **kwargs)
env.reset()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(8):
    flush.zero_()
    env.step(env.get_random_actions())
torch.cuda.synchronize()
print("done")
