"""Steps a CUDA-graph environment a few times (target process for compute-sanitizer / ncu).

    compute-sanitizer --tool memcheck python tools/run_graph_steps.py balance 64 n_agents=4
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import vectorizedmultiagentsimulator_b200 as b200

name, n_envs = sys.argv[1], int(sys.argv[2])
kwargs = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
env = b200.make_env(name, num_envs=n_envs, device="cuda", seed=0, cuda_graph=True, **kwargs)
env.reset()
for t in range(8):
    env.step(env.get_random_actions())
    torch.cuda.synchronize()
    plan = env._one_call
    print("step", t, "state", env._one_call_state, None if plan is None else (plan.direct, plan.c.fused_kernel, plan.c.ingest_in_kernel, plan.c.n_segs, plan.c.n_mirrors))
print("done")
