"""Joints: distance / rotation constraints between two entities (ref vmas/simulator/joints.py).

A joint with ``dist > 0`` inserts a Line (or Box, when ``width > 0``) landmark between its two
end points and two zero-length constraints tying the landmark's ends to the entities; a joint
with ``dist == 0`` is a single constraint.  The constraint forces themselves are evaluated by
the CUDA substep kernel (``K_JOINT`` work items); this module only owns the static anchors and
the per-env ``fixed_rotation`` values the kernel reads.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import core as _core
from .utils import Color, Observer, TorchUtils, X, Y


class Joint(Observer):
    def __init__(
        self,
        entity_a,
        entity_b,
        anchor_a: Tuple[float, float] = (0.0, 0.0),
        anchor_b: Tuple[float, float] = (0.0, 0.0),
        rotate_a: bool = True,
        rotate_b: bool = True,
        dist: float = 0.0,
        collidable: bool = False,
        width: float = 0.0,
        mass: float = 1.0,
        fixed_rotation_a: Optional[float] = None,
        fixed_rotation_b: Optional[float] = None,
    ):
        assert entity_a != entity_b, "Cannot join same entity"
        for anchor in (anchor_a, anchor_b):
            assert (
                max(anchor) <= 1 and min(anchor) >= -1
            ), f"Joint anchor points should be between -1 and 1, got {anchor}"
        assert dist >= 0, f"Joint dist must be >= 0, got {dist}"
        if dist == 0:
            assert not collidable, "Cannot have collidable joint with dist 0"
            assert width == 0, "Cannot have width for joint with dist 0"
            assert (
                fixed_rotation_a == fixed_rotation_b
            ), "If dist is 0, fixed_rotation_a and fixed_rotation_b should be the same"
        if fixed_rotation_a is not None:
            assert not rotate_a, "If you provide a fixed rotation for a, rotate_a should be False"
        if fixed_rotation_b is not None:
            assert not rotate_b, "If you provide a fixed rotation for b, rotate_b should be False"
        if width > 0:
            assert collidable

        self.entity_a = entity_a
        self.entity_b = entity_b
        self.rotate_a = rotate_a
        self.rotate_b = rotate_b
        self.fixed_rotation_a = fixed_rotation_a
        self.fixed_rotation_b = fixed_rotation_b
        self.landmark = None
        self.joint_constraints = []

        if dist == 0:
            self.joint_constraints.append(
                JointConstraint(
                    entity_a,
                    entity_b,
                    anchor_a=anchor_a,
                    anchor_b=anchor_b,
                    dist=dist,
                    rotate=rotate_a and rotate_b,
                    fixed_rotation=fixed_rotation_a,
                )
            )
            return

        entity_a.subscribe(self)
        entity_b.subscribe(self)
        link_shape = (
            _core.Box(length=dist, width=width) if width != 0 else _core.Line(length=dist)
        )
        self.landmark = _core.Landmark(
            name=f"joint {entity_a.name} {entity_b.name}",
            collide=collidable,
            movable=True,
            rotatable=True,
            mass=mass,
            shape=link_shape,
            color=Color.BLACK,
            is_joint=True,
        )
        self.joint_constraints += [
            JointConstraint(
                self.landmark,
                entity_a,
                anchor_a=(-1, 0),
                anchor_b=anchor_a,
                dist=0.0,
                rotate=rotate_a,
                fixed_rotation=fixed_rotation_a,
            ),
            JointConstraint(
                self.landmark,
                entity_b,
                anchor_a=(1, 0),
                anchor_b=anchor_b,
                dist=0.0,
                rotate=rotate_b,
                fixed_rotation=fixed_rotation_b,
            ),
        ]

    def notify(self, observable, *args, **kwargs):
        """An end point was re-placed: put the link landmark between the two anchors again."""
        end_a = self.joint_constraints[0].pos_point(self.entity_a)
        end_b = self.joint_constraints[1].pos_point(self.entity_b)
        self.landmark.set_pos((end_a + end_b) / 2, batch_index=None)
        heading = torch.atan2(end_b[:, Y] - end_a[:, Y], end_b[:, X] - end_a[:, X]).unsqueeze(-1)
        self.landmark.set_rot(heading, batch_index=None)
        # a non-rotating end without an explicit angle keeps whatever angle it has now
        if not self.rotate_a and self.fixed_rotation_a is None:
            self.joint_constraints[0].fixed_rotation = heading - self.entity_a.state.rot
        if not self.rotate_b and self.fixed_rotation_b is None:
            self.joint_constraints[1].fixed_rotation = heading - self.entity_b.state.rot


class JointConstraint:
    """Uncollidable constraint tying two anchor points at a given distance (private)."""

    def __init__(
        self,
        entity_a,
        entity_b,
        anchor_a: Tuple[float, float] = (0.0, 0.0),
        anchor_b: Tuple[float, float] = (0.0, 0.0),
        dist: float = 0.0,
        rotate: bool = True,
        fixed_rotation: Optional[float] = None,
    ):
        assert entity_a != entity_b, "Cannot join same entity"
        for anchor in (anchor_a, anchor_b):
            assert (
                max(anchor) <= 1 and min(anchor) >= -1
            ), f"Joint anchor points should be between -1 and 1, got {anchor}"
        assert dist >= 0, f"Joint dist must be >= 0, got {dist}"
        if fixed_rotation is not None:
            assert not rotate, "If fixed rotation is provided, rotate should be False"
        if rotate:
            assert fixed_rotation is None, "If you provide a fixed rotation, rotate should be False"
            fixed_rotation = 0.0

        self.entity_a = entity_a
        self.entity_b = entity_b
        self.anchor_a = anchor_a
        self.anchor_b = anchor_b
        self.dist = dist
        self.rotate = rotate
        self._fixed_rotation = fixed_rotation
        self._fixed_rotation_version = 0
        self._world = None
        self._anchor_delta_cache = {}

    @property
    def fixed_rotation(self):
        return self._fixed_rotation

    @fixed_rotation.setter
    def fixed_rotation(self, value):
        self._fixed_rotation = value
        self._fixed_rotation_version += 1  # the backend re-uploads this constraint's column

    def _delta_anchor_tensor(self, entity):
        if entity is self.entity_a:
            anchor = self.anchor_a
        elif entity is self.entity_b:
            anchor = self.anchor_b
        else:
            raise AssertionError()
        pos = entity.state.pos
        key = (id(entity), pos.device)
        t = self._anchor_delta_cache.get(key)
        if t is None:
            t = torch.tensor(
                entity.shape.get_delta_from_anchor(anchor), device=pos.device, dtype=torch.float32
            )
            self._anchor_delta_cache[key] = t
        return t.unsqueeze(0).expand(pos.shape)

    def get_delta_anchor(self, entity):
        return TorchUtils.rotate_vector(self._delta_anchor_tensor(entity), entity.state.rot)

    def pos_point(self, entity):
        return entity.state.pos + self.get_delta_anchor(entity)

    def render(self, env_index: int = 0):
        raise NotImplementedError("Rendering is outside the scope of the B200 hot-path build")
