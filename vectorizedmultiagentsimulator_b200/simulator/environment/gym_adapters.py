"""gym / gymnasium style adapters over :class:`Environment` (ref vmas/simulator/environment/gym/*.py).

Thin host-side shims — lists of per-agent tensors in, numpy (or tensors) out — with the reference's
contracts: ``GymWrapper`` and ``GymnasiumWrapper`` serve one env (``num_envs == 1``) and strip the
batch dimension, ``GymnasiumVectorizedWrapper`` keeps it; the gymnasium flavours need
``terminated_truncated=True`` and return ``(obs, rews, terminated, truncated, info)``.  They derive
from ``gym.Env`` / ``gymnasium.Env`` when that package is importable (so ``isinstance`` checks of RL
libraries hold) and work without it otherwise: the spaces are this package's own.  Nothing here
touches the physics path.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .environment import Environment


def _env_base(module_name: str):
    try:
        return __import__(module_name).Env
    except Exception:  # noqa: BLE001  (not installed: a plain object base)
        return object


def _index0(x):
    """Drops the batch dimension of a nested output (dict / list / tensor) of a one-env batch."""
    if isinstance(x, dict):
        return {k: _index0(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_index0(v) for v in x)
    return x[0]


def _to_numpy(x):
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(v) for v in x)
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


class _Adapter:
    vectorized = False

    def _setup(self, env: Environment, return_numpy: bool):
        self._env = env
        self.return_numpy = return_numpy
        self.dict_spaces = env.dict_spaces
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    @property
    def env(self) -> Environment:
        return self._env

    @property
    def unwrapped(self) -> Environment:
        return self._env

    # -- conversions -----------------------------------------------------------------------
    def _out(self, data, item: bool = False):
        if data is None:
            return None
        if not self.vectorized:
            data = _index0(data)
            if item and isinstance(data, torch.Tensor):
                return data.item()
        return _to_numpy(data) if self.return_numpy else data

    def _per_agent(self, values, item: bool = False):
        if values is None:
            return None
        if isinstance(values, dict):
            return {k: self._out(v, item) for k, v in values.items()}
        return [self._out(v, item) for v in values]

    def _infos(self, infos):
        infos = self._per_agent(infos)
        if isinstance(infos, list):  # gym wants one dict: keyed by agent name
            return {agent.name: info for agent, info in zip(self._env.agents, infos)}
        return infos

    def _actions(self, actions) -> List[torch.Tensor]:
        env = self._env
        if isinstance(actions, dict):
            actions = [actions[agent.name] for agent in env.agents]
        assert len(actions) == env.n_agents, f"Expecting actions for {env.n_agents} agents, got {len(actions)} actions"
        dtype = torch.float32 if env.continuous_actions else torch.long
        return [
            torch.as_tensor(act, dtype=dtype, device=env.device).reshape(env.num_envs, env.get_agent_action_size(agent))
            for agent, act in zip(env.agents, actions)
        ]

    def render(self, agent_index_focus: Optional[int] = None, visualize_when_rgb: bool = False, **kwargs):
        return self._env.render(
            mode=getattr(self, "render_mode", "human"), env_index=0, agent_index_focus=agent_index_focus,
            visualize_when_rgb=visualize_when_rgb, **kwargs,
        )


class GymWrapper(_Adapter, _env_base("gym")):
    """``gym.Env`` view of ONE env: ``step`` -> (obs, rews, done, info) (ref gym/gym.py:14-74)."""

    metadata = Environment.metadata if hasattr(Environment, "metadata") else {}

    def __init__(self, env: Environment, return_numpy: bool = True):
        assert env.num_envs == 1, f"GymEnv wrapper is not vectorised, got env.num_envs: {env.num_envs}"
        assert not env.terminated_truncated, "GymWrapper is not compatible with termination and truncation flags. Please set `terminated_truncated=False` in the VMAS environment."
        self._setup(env, return_numpy)

    def step(self, action):
        obs, rews, done, info = self._env.step(self._actions(action))
        return self._per_agent(obs), self._per_agent(rews, item=True), self._out(done, item=True), self._infos(info)

    def reset(self, *, seed: Optional[int] = None, return_info: bool = False, options: Optional[dict] = None):
        if seed is not None:
            self._env.seed(seed)
        obs = self._env.reset_at(index=0)
        if return_info:
            info = [self._env.scenario.info(agent) for agent in self._env.agents]
            return self._per_agent(obs), self._infos(info)
        return self._per_agent(obs)


class GymnasiumWrapper(_Adapter, _env_base("gymnasium")):
    """``gymnasium.Env`` view of ONE env (ref gym/gymnasium.py:20-89)."""

    def __init__(self, env: Environment, return_numpy: bool = True, render_mode: str = "human"):
        assert env.num_envs == 1, f"GymnasiumEnv wrapper only supports singleton VMAS environment! For vectorized environments, use vectorized wrapper with `wrapper=gymnasium_vec`."
        assert env.terminated_truncated, "GymnasiumWrapper is only compatible with termination and truncation flags. Please set `terminated_truncated=True` in the VMAS environment."
        self._setup(env, return_numpy)
        self.render_mode = render_mode

    def step(self, action):
        obs, rews, terminated, truncated, info = self._env.step(self._actions(action))
        return (
            self._per_agent(obs), self._per_agent(rews, item=True), self._out(terminated, item=True),
            self._out(truncated, item=True), self._infos(info),
        )

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self._env.seed(seed)
        obs, info = self._env.reset_at(index=0, return_info=True)
        return self._per_agent(obs), self._infos(info)


class GymnasiumVectorizedWrapper(_Adapter, _env_base("gymnasium")):
    """All ``num_envs`` envs at once, gymnasium's 5-tuple (ref gym/gymnasium_vec.py:20-98)."""

    vectorized = True

    def __init__(self, env: Environment, return_numpy: bool = True, render_mode: str = "human"):
        assert env.terminated_truncated, "GymnasiumWrapper is only compatible with termination and truncation flags. Please set `terminated_truncated=True` in the VMAS environment."
        self._setup(env, return_numpy)
        self.render_mode = render_mode
        self.num_envs = env.num_envs

    def step(self, action):
        obs, rews, terminated, truncated, info = self._env.step(self._actions(action))
        return (
            self._per_agent(obs), self._per_agent(rews), self._out(terminated), self._out(truncated), self._infos(info),
        )

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self._env.seed(seed)
        obs, info = self._env.reset(return_info=True)
        return self._per_agent(obs), self._infos(info)
