"""``BaseScenario``: the contract task definitions implement (ref vmas/simulator/scenario.py:25-441).

Compulsory: ``make_world``, ``reset_world_at``, ``observation``, ``reward``.
Optional: ``done``, ``info``, ``process_action``, ``pre_step``, ``post_step``, ``extra_render``.
The three ``env_*`` methods are the glue ``Environment`` calls and must not be overridden.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch
from torch import Tensor

from .core import Agent, World
from .utils import (
    AGENT_INFO_TYPE,
    AGENT_OBS_TYPE,
    AGENT_REWARD_TYPE,
    INITIAL_VIEWER_SIZE,
    VIEWER_DEFAULT_ZOOM,
)


class BaseScenario(ABC):
    #: True when ``reset_world_at`` also accepts a ``[B]`` bool tensor flagging the envs to reset
    #: (``Environment.reset_at(mask)``: every finished env in one pass, no host sync).  The scenarios
    #: shipped with this package do; scenario files written for the reference take an int or None.
    supports_masked_reset = False
    #: True if ``observation()`` depends on the world state only — not on anything ``reward()``, ``info()`` or
    #: ``done()`` compute or cache.  The CUDA environment then evaluates the observations on a second
    #: stream, concurrently with the reward callbacks (off by default: third-party scenarios often share
    #: cached terms between the callbacks)
    observations_are_independent = False

    def __init__(self):
        self._world = None
        # rendering knobs are kept so scenario files that set them still load
        self.viewer_size = INITIAL_VIEWER_SIZE
        self.viewer_zoom = VIEWER_DEFAULT_ZOOM
        self.render_origin = (0.0, 0.0)
        self.plot_grid = False
        self.grid_spacing = 0.1
        self.visualize_semidims = True

    @property
    def world(self) -> World:
        assert self._world is not None, "You first need to set `self._world` in the `make_world` method"
        return self._world

    def to(self, device: torch.device):
        for attr, value in self.__dict__.items():
            if isinstance(value, Tensor):
                self.__dict__[attr] = value.to(device)
        self.world.to(device)

    @staticmethod
    def keep(owner, name: str, value: Tensor, env_index=None) -> Tensor:
        """Store a tensor that carries state from one step to the next (shaping terms, timers).

        The first call binds ``owner.<name>``; later calls write into the existing tensor so its
        address never changes — the rule that makes a scenario replayable as a CUDA graph.
        With ``env_index`` (an int, or a ``[B]`` bool mask) only those envs' rows are written.
        """
        cur = getattr(owner, name, None)
        if not isinstance(cur, Tensor) or cur.shape != value.shape or cur.dtype != value.dtype:
            setattr(owner, name, value.clone())
            return getattr(owner, name)
        if env_index is None:
            cur.copy_(value)
        elif isinstance(env_index, Tensor):  # [B] bool mask of the envs being reset: no host sync
            cur.copy_(torch.where(env_index.view(-1, *([1] * (cur.dim() - 1))), value, cur))
        else:
            cur[env_index] = value[env_index]
        return cur

    # ---- glue (do not override) -----------------------------------------------------------
    def env_make_world(self, batch_dim: int, device: torch.device, **kwargs) -> World:
        self._world = self.make_world(batch_dim, device, **kwargs)
        # pack every entity's state into the contiguous slab before the first reset
        ensure = getattr(self._world, "_ensure_slab", None)
        if ensure is not None:
            ensure()
        return self._world

    def env_reset_world_at(self, env_index: Optional[int]):
        self.world.reset(env_index)
        self.reset_world_at(env_index)

    def env_process_action(self, agent: Agent):
        if agent.action_script is not None:
            agent.action_callback(self.world)
        self.process_action(agent)
        agent.dynamics.check_and_process_action()

    # ---- compulsory -----------------------------------------------------------------------
    @abstractmethod
    def make_world(self, batch_dim: int, device: torch.device, **kwargs) -> World:
        raise NotImplementedError()

    @abstractmethod
    def reset_world_at(self, env_index: Optional[int] = None):
        raise NotImplementedError()

    @abstractmethod
    def observation(self, agent: Agent) -> AGENT_OBS_TYPE:
        raise NotImplementedError()

    @abstractmethod
    def reward(self, agent: Agent) -> AGENT_REWARD_TYPE:
        raise NotImplementedError()

    # ---- optional --------------------------------------------------------------------------
    def done(self) -> Tensor:
        never = getattr(self, "_never_done", None)
        if never is None or never.shape[0] != self.world.batch_dim or never.device != self.world.slab.pos.device:
            # allocated on the device once (a per-call torch.tensor([...]) would be an H2D copy per step)
            never = torch.zeros(self.world.batch_dim, dtype=torch.bool, device=self.world.device)
            self._never_done = never
        return never

    def info(self, agent: Agent) -> AGENT_INFO_TYPE:
        return {}

    def extra_render(self, env_index: int = 0):
        return []

    def process_action(self, agent: Agent):
        return

    def pre_step(self):
        return

    def post_step(self):
        return
