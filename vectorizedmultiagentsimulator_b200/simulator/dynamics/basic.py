"""The stateless action models: the action components are used directly as force / torque.

* :class:`Holonomic` — ``u[:, :2]`` is the force (the default model; ref dynamics/holonomic.py).
* :class:`HolonomicWithRotation` — additionally ``u[:, 2]`` is the torque (ref holonomic_with_rot.py).
* :class:`Forward` — one component pushing along the agent's heading (ref forward.py).
* :class:`Rotation` — one component used as torque (ref roatation.py).
* :class:`Static` — ignores the action (ref static.py).

The state setters copy into the world's force / torque slab; for the two holonomic models the
environment normally skips these Python hooks altogether and routes the action to the slab inside
the fused ``ingest_actions`` kernel.
"""
import torch

from ..utils import TorchUtils
from .common import Dynamics


class _Direct(Dynamics):
    """``u[:, :n_force]`` → force, ``u[:, n_force : n_force + n_torque]`` → torque."""

    n_force = 0
    n_torque = 0

    @property
    def needed_action_size(self) -> int:
        return self.n_force + self.n_torque

    def process_action(self):
        u = self.agent.action.u
        if self.n_force:
            self.agent.state.force = u[:, : self.n_force]
        if self.n_torque:
            self.agent.state.torque = u[:, self.n_force : self.n_force + self.n_torque]


class Holonomic(_Direct):
    n_force = 2


class HolonomicWithRotation(_Direct):
    n_force, n_torque = 2, 1


class Rotation(_Direct):
    n_torque = 1


class Static(_Direct):
    pass


class Forward(Dynamics):
    @property
    def needed_action_size(self) -> int:
        return 1

    def process_action(self):
        agent = self.agent
        body_force = torch.zeros(agent.batch_dim, 2, device=agent.device, dtype=torch.float32)
        body_force[:, 0] = agent.action.u[:, 0]
        agent.state.force = TorchUtils.rotate_vector(body_force, agent.state.rot)
