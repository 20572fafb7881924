"""Host-side checks of the observation-plan compiler (no GPU): the column table
``vmas_b200_gather_observations`` consumes, and the CPU statement of ``World.observe``."""
import numpy as np
import pytest
import torch

import vectorizedmultiagentsimulator_b200 as b200
from oracle.backend import use_oracle
from vectorizedmultiagentsimulator_b200.simulator import observe as O


def _env(name="balance", **kwargs):
    with use_oracle():
        env = b200.make_env(name, num_envs=5, device="cpu", seed=0, **kwargs)
    return env


def test_column_table_encodes_terms():
    env = _env(n_agents=3)
    world = env.world
    a0, a1 = world.agents[0], world.agents[1]
    line = env.scenario.line
    plan = O.ObservationPlan(
        [[O.pos(a), O.rel_vel(a, line), O.blank(2), O.rot_remainder(line, torch.pi), O.ang_vel(a)] for a in (a0, a1)]
    )
    cols, lidars = plan.compile(world)
    assert cols.shape == (2, 8, 4) and cols.dtype == np.int32 and not lidars
    ents = world.entities
    ia, il = ents.index(a1), ents.index(line)
    row = cols[1]
    # pos(a1): two COPY columns reading pos[2 * ia + k]
    assert [tuple(c) for c in row[0:2]] == [(O.OP_COPY, (O.FIELD_POS << 24) | (2 * ia + k), 0, 0) for k in range(2)]
    # rel_vel(a1, line): DIFF of vel columns
    assert [tuple(c) for c in row[2:4]] == [
        (O.OP_DIFF, (O.FIELD_VEL << 24) | (2 * ia + k), (O.FIELD_VEL << 24) | (2 * il + k), 0) for k in range(2)
    ]
    # blank columns are left to the scenario
    assert (row[4:6, 0] == O.OP_SKIP).all()
    # remainder carries the fp32 bit pattern of the modulus
    op, src, _, bits = (int(v) for v in row[6])
    assert op == O.OP_REMAINDER and src == (O.FIELD_ROT << 24) | il
    assert np.array([bits], dtype=np.int32).view(np.float32)[0] == np.float32(torch.pi)
    assert tuple(row[7]) == (O.OP_COPY, (O.FIELD_ANG_VEL << 24) | ia, 0, 0)
    assert plan.column_of(0, plan.rows[0][2]) == 4


def test_rows_must_have_equal_width():
    env = _env(n_agents=3)
    a0, a1 = env.world.agents[:2]
    with pytest.raises(ValueError):
        O.ObservationPlan([[O.pos(a0)], [O.pos(a1), O.rot(a1)]])


def test_cpu_statement_matches_per_term_expressions():
    env = _env("navigation", n_agents=3)
    world = env.world
    env.step(env.get_random_actions())
    agents = world.agents
    plan = O.ObservationPlan(
        [[O.pos(a), O.vel(a), O.rel_pos(a, a.goal), O.lidar(a.sensors[0], range_minus_distance=True)] for a in agents]
    )
    block = world.observe(plan)
    assert block.shape == (3, 5, 6 + 12)
    for i, a in enumerate(agents):
        want = torch.cat(
            [a.state.pos, a.state.vel, a.state.pos - a.goal.state.pos, a.sensors[0]._max_range - a.sensors[0].measure()],
            dim=-1,
        )
        assert torch.equal(block[i], want)
    cols, lidars = plan.compile(world)
    assert [(r, c) for r, c, _, _ in lidars] == [(0, 6), (1, 6), (2, 6)] and all(f for *_, f in lidars)
    assert (cols[:, 6:, 0] == O.OP_SKIP).all()


def test_plan_recompiles_when_the_world_changes():
    env = _env(n_agents=3)
    world = env.world
    plan = O.ObservationPlan([[O.pos(a)] for a in world.agents])
    first, _ = plan.compile(world)
    assert plan.compile(world)[0] is first  # cached per plan version
    world._plan_version += 1
    assert plan.compile(world)[0] is not first
