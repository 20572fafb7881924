"""The host side of the fused action ingest, without a GPU: ``CudaBackend.ingest_actions`` fills the
``VmasAgentActions`` array (pointers, sizes, dynamics codes and parameters, discrete-space tables) and
launches once per 16 agents.  A stand-in for ``_native.ingest_actions`` captures what would cross the C
ABI.  (Regression: a loop variable once shadowed the agent count in the discrete path, so the launch
covered ``nvec[-1]`` agents instead of all of them.)"""
import ctypes as C
import types

import pytest
import torch

import vectorizedmultiagentsimulator_b200 as b200
from oracle.backend import use_oracle
from vectorizedmultiagentsimulator_b200 import _native
from vectorizedmultiagentsimulator_b200.backend import CudaBackend


class _Capture:
    def __init__(self):
        self.calls = []
        for name in dir(_native):
            if name.isupper() or name == "AgentActionsC":
                setattr(self, name, getattr(_native, name))

    def ingest_actions(self, lib, dt, slab, chunk, n, clamp, bad_flag, steps=None, broad_phase=False):
        self.calls.append([{f[0]: (list(getattr(chunk[i], f[0])) if hasattr(getattr(chunk[i], f[0]), "__len__") else getattr(chunk[i], f[0])) for f in _native.AgentActionsC._fields_} for i in range(n)])
        return 1


def _fake_backend(env):
    world = env.world
    index = {id(e): i for i, e in enumerate(world.entities)}
    cap = _Capture()
    fake = types.SimpleNamespace(
        world=world, _native=cap, lib=None, _dev_tables=None, launches=0, refresh=lambda: None,
        index_of=lambda agent: index[id(agent)], tables=types.SimpleNamespace(n_masked=0),
    )
    return fake, cap


@pytest.mark.parametrize("n_agents,kind", [(2, _native.ACT_DISCRETE), (4, _native.ACT_MULTIDISCRETE), (5, _native.ACT_CONTINUOUS), (19, _native.ACT_DISCRETE)])
def test_every_agent_crosses_the_abi(n_agents, kind):
    with use_oracle():
        env = b200.make_env("navigation", num_envs=3, device="cpu", seed=0, n_agents=n_agents, continuous_actions=kind == _native.ACT_CONTINUOUS)
    fake, cap = _fake_backend(env)
    specs = [(a, _native.DYN_HOLONOMIC, torch.zeros(3, a.action_size)) for a in env.agents]
    if kind == _native.ACT_CONTINUOUS:
        actions = [torch.rand(3, 2) for _ in env.agents]
    else:
        actions = [torch.zeros(3, 2 if kind == _native.ACT_MULTIDISCRETE else 1, dtype=torch.int64) for _ in env.agents]
    CudaBackend.ingest_actions(fake, actions, specs, True, None, action_kind=kind)
    sent = [c for call in cap.calls for c in call]
    assert len(sent) == n_agents and len(cap.calls) == (n_agents + 15) // 16 and fake.launches == len(cap.calls)
    for j, (c, a, (agent, _, u)) in enumerate(zip(sent, actions, specs)):
        assert c["actions"] == a.data_ptr() and c["u"] == u.data_ptr(), f"agent {j}: pointers"
        assert c["action_size"] == 2 and c["agent_index"] == j and c["entity_index"] == env.world.entities.index(agent)
        assert c["action_kind"] == kind
        assert c["nvec"][:2] == ([3, 3] if kind != _native.ACT_CONTINUOUS else [0, 0])
        assert c["u_range"][:2] == [1.0, 1.0] and c["u_multiplier"][:2] == [1.0, 1.0]


def test_kinematic_models_carry_their_parameters():
    import crafted

    with use_oracle():
        env = b200.make_env(crafted.make_scenario("vectorizedmultiagentsimulator_b200", "dynamics_zoo"), num_envs=2, device="cpu", seed=0)
    fake, cap = _fake_backend(env)
    codes = dict(diff_rk4=_native.DYN_DIFF_DRIVE, diff_euler=_native.DYN_DIFF_DRIVE, bicycle=_native.DYN_BICYCLE,
                 bicycle_euler=_native.DYN_BICYCLE, drone=_native.DYN_DRONE, forward=_native.DYN_FORWARD,
                 rotation=_native.DYN_ROTATION, holo_rot=_native.DYN_HOLONOMIC_ROT, holo=_native.DYN_HOLONOMIC)
    agents = [a for a in env.agents if a.name in codes]
    specs = [(a, codes[a.name], torch.zeros(2, a.action_size)) for a in agents]
    CudaBackend.ingest_actions(fake, [torch.zeros(2, a.action_size) for a in agents], specs, True, None, action_kind=_native.ACT_CONTINUOUS)
    sent = {a.name: c for a, c in zip(agents, cap.calls[0])}
    assert len(cap.calls[0]) == len(agents)
    assert sent["diff_rk4"]["dyn_params"][3] == 1.0 and sent["diff_euler"]["dyn_params"][3] == 0.0
    assert sent["bicycle"]["dyn_params"][4:7] == pytest.approx([0.06, 0.05, 0.6])
    drone = next(a for a in agents if a.name == "drone")
    assert sent["drone"]["dyn_state"] == drone.dynamics.drone_state.data_ptr()
    assert sent["drone"]["dyn_params"][1] == pytest.approx(0.1) and sent["drone"]["dyn_params"][7] == pytest.approx(9.81)
    assert sent["forward"]["dyn_state"] in (None, 0)
