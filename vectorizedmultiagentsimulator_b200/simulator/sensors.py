"""Sensors.  ``Lidar.measure`` is one launch of the ray-cast kernel (ref vmas/simulator/sensors.py)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Tuple, Union

import torch

from .utils import Color


class Sensor(ABC):
    def __init__(self, world):
        super().__init__()
        self._world = world
        self._agent = None

    @property
    def agent(self):
        return self._agent

    @agent.setter
    def agent(self, agent):
        self._agent = agent

    @abstractmethod
    def measure(self):
        raise NotImplementedError

    def render(self, env_index: int = 0):
        raise NotImplementedError("Rendering is outside the scope of the B200 hot-path build")

    def to(self, device: torch.device):
        raise NotImplementedError


class Lidar(Sensor):
    """``n_rays`` equally spaced rays in the agent's frame (ref sensors.py:47-123).

    A full-circle sweep drops the duplicated end angle (ref sensors.py:61-68).  ``measure``
    returns ``[B, n_rays]`` ranges, ``max_range`` where nothing is hit.
    """

    def __init__(
        self,
        world,
        angle_start: float = 0.0,
        angle_end: float = 2 * torch.pi,
        n_rays: int = 8,
        max_range: float = 1.0,
        entity_filter: Callable = lambda _: True,
        render_color: Union[Color, Tuple[float, float, float]] = Color.GRAY,
        alpha: float = 1.0,
        render: bool = True,
    ):
        super().__init__(world)
        full_circle = (angle_start - angle_end) % (torch.pi * 2) < 1e-5
        sweep = torch.linspace(
            angle_start, angle_end, n_rays + 1 if full_circle else n_rays, device=world.device
        )[:n_rays]
        self._angles = sweep.repeat(world.batch_dim, 1)
        self._max_range = max_range
        self._last_measurement = None
        self._render = render
        self._entity_filter = entity_filter
        self._render_color = render_color
        self._alpha = alpha

    def to(self, device: torch.device):
        self._angles = self._angles.to(device)

    @property
    def entity_filter(self):
        return self._entity_filter

    @entity_filter.setter
    def entity_filter(self, entity_filter: Callable):
        self._entity_filter = entity_filter
        if self._world is not None:
            self._world._invalidate_plan()

    @property
    def render_color(self):
        if isinstance(self._render_color, Color):
            return self._render_color.value
        return self._render_color

    @property
    def alpha(self):
        return self._alpha

    def measure(self, vectorized: bool = True):
        # ``vectorized`` is accepted for API compatibility: both values run the same kernel
        # (the reference's per-ray python loop, sensors.py:102-113, has no counterpart here).
        measurement = self._world._get_backend().lidar_measure(self)
        self._last_measurement = measurement
        return measurement

    def set_render(self, render: bool):
        self._render = render
