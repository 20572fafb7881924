"""Constants and small helpers shared by the host-side object model.

Mirrors the public names scenarios import from the reference's
``vmas/simulator/utils.py`` (constants :22-35, ``Color`` :50-61, ``Observable`` /
``Observer`` :85-103, ``TorchUtils`` :166-236, ``ScenarioUtils`` :239-330).  Written
from the behaviour described in SURVEY.md; nothing here is on the CUDA hot path.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from enum import Enum
from typing import Dict, List, Sequence, Tuple, Union

import torch
from torch import Tensor

X, Y, Z = 0, 1, 2
ALPHABET = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
VIEWER_DEFAULT_ZOOM = 1.2
INITIAL_VIEWER_SIZE = (700, 700)

# physics constants (reference utils.py:28-35)
LINE_MIN_DIST = 4 / 6e2
COLLISION_FORCE = 100
JOINT_FORCE = 130
TORQUE_CONSTRAINT_FORCE = 1
DRAG = 0.25
LINEAR_FRICTION = 0.0
ANGULAR_FRICTION = 0.0

DEVICE_TYPING = Union[torch.device, str, int]
AGENT_OBS_TYPE = Union[Tensor, Dict[str, Tensor]]
AGENT_INFO_TYPE = Dict[str, Tensor]
AGENT_REWARD_TYPE = Tensor
OBS_TYPE = Union[List[AGENT_OBS_TYPE], Dict[str, AGENT_OBS_TYPE]]
INFO_TYPE = Union[List[AGENT_INFO_TYPE], Dict[str, AGENT_INFO_TYPE]]
REWARD_TYPE = Union[List[AGENT_REWARD_TYPE], Dict[str, AGENT_REWARD_TYPE]]
DONE_TYPE = Tensor


class Color(Enum):
    RED = (0.75, 0.25, 0.25)
    GREEN = (0.25, 0.75, 0.25)
    BLUE = (0.25, 0.25, 0.75)
    LIGHT_GREEN = (0.45, 0.95, 0.45)
    WHITE = (0.75, 0.75, 0.75)
    GRAY = (0.25, 0.25, 0.25)
    BLACK = (0.15, 0.15, 0.15)
    ORANGE = (1.00, 0.50, 0)
    PINK = (0.97, 0.51, 0.75)
    PURPLE = (0.60, 0.31, 0.64)
    YELLOW = (0.87, 0.87, 0)


def override(cls):
    """Marks a method as overriding ``cls``; fails at import time if it does not."""

    def check(method):
        if method.__name__ not in dir(cls):
            raise NameError(f"{method} does not override any method of {cls}")
        return method

    return check


class Observable:
    """Subject side of the observer protocol joints use to follow their end points."""

    def __init__(self):
        self._observers = []

    def subscribe(self, observer):
        self._observers.append(observer)

    def unsubscribe(self, observer):
        self._observers.remove(observer)

    def notify_observers(self, *args, **kwargs):
        for o in self._observers:
            o.notify(self, *args, **kwargs)


class Observer(ABC):
    @abstractmethod
    def notify(self, observable, *args, **kwargs):
        raise NotImplementedError


def extract_nested_with_index(data, index: int):
    if isinstance(data, Tensor):
        return data[index]
    if isinstance(data, dict):
        return {k: extract_nested_with_index(v, index) for k, v in data.items()}
    raise NotImplementedError(f"Invalid type of data {data}")


class TorchUtils:
    """Elementwise 2-D helpers kept for scenario code (reference utils.py:166-236)."""

    @staticmethod
    def clamp_with_norm(tensor: Tensor, max_norm: float) -> Tensor:
        norm = torch.linalg.vector_norm(tensor, dim=-1, keepdim=True)
        rescaled = (tensor / norm) * max_norm
        return torch.where(norm > max_norm, rescaled, tensor)

    @staticmethod
    def rotate_vector(vector: Tensor, angle: Tensor) -> Tensor:
        if angle.dim() == vector.dim():
            angle = angle.squeeze(-1)
        assert vector.shape[:-1] == angle.shape
        assert vector.shape[-1] == 2
        c, s = torch.cos(angle), torch.sin(angle)
        vx, vy = vector[..., X], vector[..., Y]
        return torch.stack([vx * c - vy * s, vx * s + vy * c], dim=-1)

    @staticmethod
    def cross(vector_a: Tensor, vector_b: Tensor) -> Tensor:
        return (
            vector_a[..., X] * vector_b[..., Y] - vector_a[..., Y] * vector_b[..., X]
        ).unsqueeze(-1)

    @staticmethod
    def compute_torque(f: Tensor, r: Tensor) -> Tensor:
        return TorchUtils.cross(r, f)

    @staticmethod
    def to_numpy(data):
        if isinstance(data, Tensor):
            return data.cpu().detach().numpy()
        if isinstance(data, dict):
            return {k: TorchUtils.to_numpy(v) for k, v in data.items()}
        if isinstance(data, Sequence):
            return [TorchUtils.to_numpy(v) for v in data]
        raise NotImplementedError(f"Invalid type of data {data}")

    @staticmethod
    def recursive_clone(value):
        if isinstance(value, Tensor):
            return value.clone()
        return {k: TorchUtils.recursive_clone(v) for k, v in value.items()}

    @staticmethod
    def recursive_require_grad_(value):
        if isinstance(value, Tensor):
            if torch.is_floating_point(value):
                value.requires_grad_(True)
        elif isinstance(value, dict):
            for v in value.values():
                TorchUtils.recursive_require_grad_(v)
        else:
            for v in value:
                TorchUtils.recursive_require_grad_(v)

    @staticmethod
    def where_from_index(env_index, new_value, old_value):
        mask = torch.zeros_like(old_value, dtype=torch.bool)
        mask[env_index] = True
        return torch.where(mask, new_value, old_value)


class ScenarioUtils:
    """Reset-time helpers (reference utils.py:239-330)."""

    @staticmethod
    def spawn_entities_randomly(
        entities,
        world,
        env_index,
        min_dist_between_entities: float,
        x_bounds: Tuple[float, float],
        y_bounds: Tuple[float, float],
        occupied_positions: Tensor = None,
        disable_warn: bool = False,
    ):
        if getattr(world, "uses_device_reset", False):
            # CUDA world: every entity of the call in one kernel launch, no host sync
            entities = list(entities)
            world.spawn_positions(
                entities, env_index, min_dist_between_entities, x_bounds, y_bounds, occupied_positions
            )
            for entity in entities:
                entity.notify_observers()
            return
        n = world.batch_dim if not isinstance(env_index, int) else 1
        if occupied_positions is None:
            occupied_positions = torch.zeros((n, 0, world.dim_p), device=world.device)
        for entity in entities:
            pos = ScenarioUtils.find_random_pos_for_entity(
                occupied_positions,
                env_index,
                world,
                min_dist_between_entities,
                x_bounds,
                y_bounds,
                disable_warn,
            )
            occupied_positions = torch.cat([occupied_positions, pos], dim=1)
            entity.set_pos(pos.squeeze(1), batch_index=env_index)

    @staticmethod
    def find_random_pos_for_entity(
        occupied_positions: Tensor,
        env_index,
        world,
        min_dist_between_entities: float,
        x_bounds: Tuple[float, float],
        y_bounds: Tuple[float, float],
        disable_warn: bool = False,
    ) -> Tensor:
        """Rejection-samples a ``[n, 1, 2]`` position at least ``min_dist`` from the occupied ones.

        Draw order (x then y, one ``uniform_`` each per attempt) follows the reference so a
        CPU world seeded the same way spawns the same layout.
        """
        if getattr(world, "uses_device_reset", False):
            out = world.spawn_positions(
                [None], env_index, min_dist_between_entities, x_bounds, y_bounds, occupied_positions,
                want_positions=True,
            )
            return out[env_index].unsqueeze(0) if isinstance(env_index, int) else out
        n = world.batch_dim if not isinstance(env_index, int) else 1
        pos = None
        tries = 0
        while True:
            px = torch.empty((n, 1, 1), device=world.device, dtype=torch.float32)
            px.uniform_(*x_bounds)
            py = torch.empty((n, 1, 1), device=world.device, dtype=torch.float32)
            py.uniform_(*y_bounds)
            proposal = torch.cat([px, py], dim=2)
            if pos is None:
                pos = proposal
            if occupied_positions.shape[1] == 0:
                break
            too_close = (torch.cdist(occupied_positions, pos) < min_dist_between_entities)
            overlaps = too_close.squeeze(2).any(dim=1)
            if not bool(overlaps.any()):
                break
            pos[overlaps] = proposal[overlaps]
            tries += 1
            if tries > 50_000 and not disable_warn:
                warnings.warn(
                    "Spawning an entity is taking many iterations: bounds or "
                    "min_dist_between_entities may be too tight. "
                    "Pass disable_warn=True to silence this."
                )
        return pos

    @staticmethod
    def check_kwargs_consumed(dictionary_of_kwargs: Dict, warn: bool = True):
        if len(dictionary_of_kwargs) == 0:
            return
        message = (
            f"Scenario kwargs: {dictionary_of_kwargs} passed but not used by the scenario."
        )
        if warn:
            warnings.warn(message + " This will turn into an error in future versions.")
        else:
            raise ValueError(message)
