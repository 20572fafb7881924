"""ORACLE (test infrastructure) — plugs the CPU restatement behind the ``World`` API.

``use_oracle()`` is a context manager that swaps ``World``'s backend factory so worlds created
inside it (on ``device="cpu"``) step through :mod:`oracle.world_step` instead of the CUDA
kernels.  It exists so the host-side object model / ``Environment`` can be checked against the
reference on CPU, and so ``bench.py --impl reference`` can time a CPU baseline.
"""
from __future__ import annotations

import contextlib

import torch

from vectorizedmultiagentsimulator_b200.backend import PlanRuntime
from vectorizedmultiagentsimulator_b200.simulator.core import World

from . import queries, world_step


class OracleBackend(PlanRuntime):
    def __init__(self, world):
        super().__init__(world)
        # a CPU checker; bench.py's reference arm may also run this same eager torch op chain on a
        # CUDA device (`use_oracle(allow_cuda=True)`): the reference's PyTorch-CUDA path, timed
        assert _ALLOW_CUDA or torch.device(world.device).type == "cpu", "the oracle is a CPU checker"

    def _state(self):
        slab = self.world.slab
        return dict(
            pos=slab.pos, vel=slab.vel, rot=slab.rot, ang_vel=slab.ang_vel, force=slab.force, torque=slab.torque
        )

    def step(self):
        self.refresh()
        world_step.world_step(
            self.tables,
            self._state(),
            fixed_rot=self.per_env_fixed_rotations(),
            exact_broad_phase=self.world.exact_broad_phase,
            ent_gravity={
                i: e.gravity
                for i, e in enumerate(self.world.entities)
                if self.tables.desc.entities[i].get("gravity_per_env")
            },
        )

    def cast_rays(self, entity, angles, max_range, entity_filter):
        src = self.index_of(entity)
        targets = self.ray_targets(entity, entity_filter)
        slab = self.world.slab
        angles = angles.to(torch.float32)
        return queries.cast_rays(self.tables, slab.pos, slab.rot, src, targets, angles, max_range)

    def lidar_measure(self, sensor):
        return self.cast_rays(
            sensor.agent, sensor._angles + sensor.agent.state.rot, sensor._max_range, sensor.entity_filter
        )

    def lidar_measure_many(self, sensors):
        return torch.stack([self.lidar_measure(s) for s in sensors])

    def observe(self, plan):
        """CPU statement of ``World.observe``: the per-term torch expressions the reference's
        scenarios write (e.g. balance.py:236-262), concatenated per row."""
        from vectorizedmultiagentsimulator_b200.simulator import observe as O

        plan.compile(self.world)  # validates the plan
        rows = []
        for row in plan.rows:
            parts = []
            for t in row:
                if isinstance(t, O._State):
                    value = getattr(t.entity.state, t.field)
                    if t.minus is not None:
                        value = value - getattr(t.minus.state, t.field)
                    elif t.modulus is not None:
                        value = value % t.modulus
                    parts.append(value)
                elif isinstance(t, O._Lidar):
                    reading = t.sensor.measure()
                    parts.append(t.sensor._max_range - reading if t.range_minus_distance else reading)
                elif isinstance(t, O._Buffer):  # a per-env value another producer holds (ref transport.py:177-183)
                    src = t.source
                    held = src.tensor if hasattr(src, "_slot") else (src() if callable(src) else src)
                    parts.append(held.to(torch.float32).unsqueeze(-1))
                else:
                    parts.append(torch.zeros(self.world.batch_dim, t.width, device=self.world.device))
            rows.append(torch.cat(parts, dim=-1))
        return torch.stack(rows)

    def run_program(self, prog, observe=None):
        """CPU statement of a ``program.StepProgram``: the torch ops the instruction list stands for
        (ref scenarios/balance.py:197-263: per-pair queries, the shaping pattern, elementwise glue)."""
        from vectorizedmultiagentsimulator_b200.simulator import program as SP

        regs = {}
        for op, dst, a, b, arg, imm, entities in prog.instr:
            if op == SP.OP_OVERLAP:
                regs[dst] = self.pair_overlap(*entities)
            elif op == SP.OP_DISTANCE:
                regs[dst] = self.pair_distance(*entities)
            elif op == SP.OP_CENTER_DISTANCE:
                regs[dst] = torch.linalg.vector_norm(entities[0].state.pos - entities[1].state.pos, dim=-1)
            elif op == SP.OP_SHAPING:
                prev = prog.resolve(prog.buffers[a])
                dist = torch.linalg.vector_norm(entities[0].state.pos - entities[1].state.pos, dim=-1)
                shaping = dist * imm
                regs[dst], regs[dst + 1] = prev - shaping, dist
                prev.copy_(shaping)
            elif op == SP.OP_LOAD_F32:
                regs[dst] = prog.resolve(prog.buffers[a]).clone()
            elif op == SP.OP_LOAD_BOOL:
                regs[dst] = prog.resolve(prog.buffers[a]).to(torch.bool)
            elif op == SP.OP_CONST:
                regs[dst] = torch.tensor(imm, dtype=torch.float32)
            elif op == SP.OP_ADD:
                regs[dst] = regs[a] + regs[b]
            elif op == SP.OP_SUB:
                regs[dst] = regs[a] - regs[b]
            elif op == SP.OP_MUL:
                regs[dst] = regs[a] * regs[b]
            elif op == SP.OP_MIN:
                regs[dst] = torch.minimum(regs[a], regs[b])
            elif op == SP.OP_MAX:
                regs[dst] = torch.maximum(regs[a], regs[b])
            elif op == SP.OP_NEG:
                regs[dst] = -regs[a]
            elif op == SP.OP_OR:
                regs[dst] = regs[a].to(torch.bool) | regs[b].to(torch.bool)
            elif op == SP.OP_AND:
                regs[dst] = regs[a].to(torch.bool) & regs[b].to(torch.bool)
            elif op == SP.OP_NOT:
                regs[dst] = ~regs[a].to(torch.bool)
            elif op == SP.OP_LT:
                regs[dst] = regs[a] < regs[b]
            elif op == SP.OP_LE:
                regs[dst] = regs[a] <= regs[b]
            elif op == SP.OP_WHERE:
                regs[dst] = torch.where(regs[a].to(torch.bool), regs[b], regs[arg])
            elif op in (SP.OP_STORE_F32, SP.OP_STORE_BOOL):
                out = prog.resolve(prog.buffers[b])
                out.copy_(regs[a].expand_as(out).to(out.dtype))
            else:
                raise ValueError(f"unknown program op {op}")
        return self.observe(observe) if observe is not None else None

    def distance_shaping(self, pairs, factor, prev):
        """CPU statement of ``World.distance_shaping`` (ref scenarios/balance.py:197-214)."""
        dist = torch.stack([torch.linalg.vector_norm(a.state.pos - b.state.pos, dim=-1) for a, b in pairs])
        shaping = dist * factor
        rew = prev - shaping
        prev.copy_(shaping)
        return dist, rew

    def pair_query_many(self, pairs, mode):
        if mode == 0:
            return torch.stack([self.pair_distance(a, b) for a, b in pairs])
        if mode == 1:
            return torch.stack([self.pair_overlap(a, b) for a, b in pairs])
        return torch.stack([torch.linalg.vector_norm(a.state.pos - b.state.pos, dim=-1) for a, b in pairs])

    def pair_distance(self, a, b):
        self.refresh()
        slab = self.world.slab
        return queries.pair_distance(self.tables, slab.pos, slab.rot, self.index_of(a), self.index_of(b))

    def pair_overlap(self, a, b):
        self.refresh()
        slab = self.world.slab
        return queries.pair_overlap(self.tables, slab.pos, slab.rot, self.index_of(a), self.index_of(b))

    def distance_from_point(self, entity, point):
        self.refresh()
        slab = self.world.slab
        return queries.distance_from_point(self.tables, slab.pos, slab.rot, self.index_of(entity), point)


_ALLOW_CUDA = False


@contextlib.contextmanager
def use_oracle(allow_cuda: bool = False):
    global _ALLOW_CUDA
    previous, previous_allow = World._backend_factory, _ALLOW_CUDA
    World._backend_factory = OracleBackend
    _ALLOW_CUDA = allow_cuda
    try:
        yield
    finally:
        World._backend_factory = previous
        _ALLOW_CUDA = previous_allow
