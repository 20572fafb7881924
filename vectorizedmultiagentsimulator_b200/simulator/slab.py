"""The world's state slab: the HBM layout the CUDA kernels stream.

The reference keeps one small tensor per entity per field (``EntityState._spawn``,
reference core.py:304-316) and re-allocates them every substep.  Here a world owns six
contiguous fp32 tensors and every ``entity.state.<field>`` is a strided *view* into them:

    pos      [B, E, 2]      vel      [B, E, 2]
    rot      [B, E]         ang_vel  [B, E]
    force    [B, A, 2]      torque   [B, A]        (A = agents, in ``world.agents`` order)

``E`` follows ``world.entities`` order (landmarks first, then agents — reference
core.py:1220-1222).  One env's data is a contiguous run of ``E*8`` / ``E*4`` bytes, so a
tile of consecutive envs is a single contiguous range for every field.
"""
from __future__ import annotations

from typing import List

import torch


class StateSlab:
    def __init__(self, batch_dim: int, device: torch.device, entities: List, agents: List):
        self.batch_dim = batch_dim
        self.device = device
        self.n_entities = len(entities)
        self.n_agents = len(agents)
        B, E, A = batch_dim, max(self.n_entities, 1), max(self.n_agents, 1)
        kw = dict(device=device, dtype=torch.float32)
        # CUDA: env-major, contiguous [B, E, ...] — what the kernels stream.  A CPU world can only
        # be driven by the test oracle; there the same logical [B, E, ...] tensors are backed by
        # entity-major storage so each entity's [B, 2] view is contiguous, like the per-entity
        # tensors of the reference (keeps the CPU baseline's memory behaviour honest).
        self.env_major = torch.device(device).type != "cpu"
        if self.env_major:
            self.pos = torch.zeros(B, E, 2, **kw)
            self.vel = torch.zeros(B, E, 2, **kw)
            self.rot = torch.zeros(B, E, **kw)
            self.ang_vel = torch.zeros(B, E, **kw)
            self.force = torch.zeros(B, A, 2, **kw)
            self.torque = torch.zeros(B, A, **kw)
        else:
            self.pos = torch.zeros(E, B, 2, **kw).permute(1, 0, 2)
            self.vel = torch.zeros(E, B, 2, **kw).permute(1, 0, 2)
            self.rot = torch.zeros(E, B, **kw).permute(1, 0)
            self.ang_vel = torch.zeros(E, B, **kw).permute(1, 0)
            self.force = torch.zeros(A, B, 2, **kw).permute(1, 0, 2)
            self.torque = torch.zeros(A, B, **kw).permute(1, 0)
        self.entity_names = [e.name for e in entities]

    def tensors(self):
        return (self.pos, self.vel, self.rot, self.ang_vel, self.force, self.torque)

    def bind(self, entities: List, agents: List):
        """Moves every entity's current state into the slab and re-points it at slab views."""
        for i, e in enumerate(entities):
            st = e.state
            views = {
                "pos": self.pos[:, i, :],
                "vel": self.vel[:, i, :],
                "rot": self.rot[:, i : i + 1],
                "ang_vel": self.ang_vel[:, i : i + 1],
            }
            for name, view in views.items():
                old = st._fields.get(name)
                if old is not None:
                    view.copy_(old)
                st._fields[name] = view
        for j, a in enumerate(agents):
            st = a.state
            views = {
                "force": self.force[:, j, :],
                "torque": self.torque[:, j : j + 1],
            }
            for name, view in views.items():
                old = st._fields.get(name)
                if old is not None:
                    view.copy_(old)
                st._fields[name] = view

    def state_dict(self):
        return {
            "pos": self.pos.clone(),
            "vel": self.vel.clone(),
            "rot": self.rot.clone(),
            "ang_vel": self.ang_vel.clone(),
            "force": self.force.clone(),
            "torque": self.torque.clone(),
        }

    def load_state_dict(self, sd):
        for k in ("pos", "vel", "rot", "ang_vel", "force", "torque"):
            if k in sd:
                getattr(self, k).copy_(sd[k])
