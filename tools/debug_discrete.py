import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vectorizedmultiagentsimulator_b200 as b200
for cont in (True, False):
    env = b200.make_env("navigation", num_envs=8, device="cuda", seed=0, n_agents=4, continuous_actions=cont)
    if cont:
        acts = [torch.rand(8, 2, device="cuda") for _ in env.agents]
    else:
        acts = [torch.randint(0, 9, (8, 1), device="cuda") for _ in env.agents]
    print("continuous", cont, "fused applies", env._fused_ingest_applies(acts))
    env.step(acts)
    torch.cuda.synchronize()
    be = env.world._get_backend()
    arr = be._ingest_arr
    for c, a, ag in zip(arr, acts, env.agents):
        print(ag.name, "kind", c.action_kind, "nvec", list(c.nvec)[:3], "size", c.action_size, "agent_index", c.agent_index, "entity", c.entity_index,
              "dyn", c.dynamics, "u ptr ok", c.u == ag.action.u.data_ptr(), "actions ptr ok", c.actions == a.data_ptr(), "|u|", float(ag.action.u.abs().sum()))
    print("force abs sum per agent", env.world.slab.force.abs().sum(dim=(0, 2)).tolist())
