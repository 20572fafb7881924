import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vectorizedmultiagentsimulator_b200 as b200
from envutil import flatten, sync_env
name, kwargs = "balance", dict(n_agents=4)
n_envs = 48
eager = b200.make_env(name, num_envs=n_envs, device="cuda", seed=0, **kwargs)
graph = b200.make_env(name, num_envs=n_envs, device="cuda", seed=0, cuda_graph=True, **kwargs)
sync_env(eager, graph)
gen = torch.Generator().manual_seed(3)
labels = None
for t in range(9):
    actions = [((torch.rand(n_envs, a.action_size, generator=gen) * 2 - 1)).cuda() for a in eager.agents]
    want = eager.step([a.clone() for a in actions]); got = graph.step([a.clone() for a in actions])
    fw, fg = flatten(want), flatten(got)
    for i, (g, w) in enumerate(zip(fg, fw)):
        if not torch.equal(g, w):
            d = (g.float() - w.float()).abs()
            rows = (d.reshape(n_envs, -1).max(1)[0] > 0).nonzero().flatten().tolist()
            print(f"step {t} output {i} shape {tuple(g.shape)} max diff {float(d.max()):.3e} envs {rows[:10]}")
    print("step", t, "slab equal", all(torch.equal(a, b) for a, b in zip(eager.world.slab.tensors(), graph.world.slab.tensors())),
          "shaping equal", torch.equal(eager.scenario.global_shaping, graph.scenario.global_shaping))
    if t == 4:
        eager.reset_at(5); graph.reset_at(5); sync_env(eager, graph)
        print("after reset_at+sync: slab equal", all(torch.equal(a, b) for a, b in zip(eager.world.slab.tensors(), graph.world.slab.tensors())),
              torch.equal(eager.scenario.global_shaping, graph.scenario.global_shaping))
