import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vectorizedmultiagentsimulator_b200 as b200
for n_agents in (2, 4, 5, 7):
    env = b200.make_env("navigation", num_envs=8, device="cuda", seed=0, n_agents=n_agents, continuous_actions=False)
    acts = [torch.randint(0, 9, (8, 1), device="cuda") for _ in env.agents]
    env.step(acts)
    torch.cuda.synchronize()
    print(n_agents, "actions sums", [int(a.sum()) for a in acts], "|u|", [float(ag.action.u.abs().sum()) for ag in env.agents])
env = b200.make_env("balance", num_envs=8, device="cuda", seed=0, n_agents=4, continuous_actions=False)
acts = [torch.randint(0, 9, (8, 1), device="cuda") for _ in env.agents]
env.step(acts); torch.cuda.synchronize()
print("balance4", [float(ag.action.u.abs().sum()) for ag in env.agents])
