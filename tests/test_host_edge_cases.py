"""Edge cases of the host layer and the plan compiler on the CPU oracle backend: the smallest and the
emptiest worlds the reference accepts, and the documented limits."""
import pytest
import torch

from oracle.backend import use_oracle
from vectorizedmultiagentsimulator_b200 import _native
from vectorizedmultiagentsimulator_b200.simulator import plan as P
from vectorizedmultiagentsimulator_b200.simulator.core import Agent, Landmark, Sphere, World


def _world(batch_dim, n_agents, n_landmarks, collide=True):
    world = World(batch_dim, "cpu")
    for i in range(n_agents):
        world.add_agent(Agent(name=f"a{i}", shape=Sphere(0.05), collide=collide))
    for i in range(n_landmarks):
        world.add_landmark(Landmark(name=f"l{i}", shape=Sphere(0.05), collide=collide))
    world._ensure_slab()
    return world


def test_single_env_single_agent_world_steps():
    with use_oracle():
        world = _world(1, 1, 0)
        desc = P.describe_world(world)
        assert desc.n_entities == 1 and len(desc.items) == 0
        tables = P.build_tables(desc)  # zero items: the tables keep one placeholder row, n_items says 0
        assert _native.make_config(tables).n_items == 0
        # one movable, rotatable agent: 12 E + 24 M + 12 R + 12 A bytes per env per substep
        assert P.algorithmic_bytes_per_env_substep(desc) == 12 + 24 + 12 + 12
        world.agents[0].state.force = torch.tensor([[1.0, 0.0]])
        world.step()
        assert float(world.agents[0].state.vel[0, 0]) > 0 and float(world.agents[0].state.pos[0, 0]) > 0


def test_world_without_collidable_pairs_has_no_work_items():
    with use_oracle():
        world = _world(3, 2, 2, collide=False)
        desc = P.describe_world(world)
        assert len(desc.items) == 0
        world.step()  # nothing to resolve, must still integrate
        assert torch.isfinite(world.slab.pos).all()
    # the device tables of an item-free world can be built (zero-row tables) for every generic mapping
    tables = P.build_tables(desc)
    for mapping in ("thread_per_env", "lanes_per_env"):
        dt = _native.DeviceTables(tables, None, torch.device("cpu"), mapping=mapping)
        assert dt.cfg.n_items == 0 and dt.cfg.batch_dim == 3


def test_static_landmarks_only_generate_no_pairs_among_themselves():
    with use_oracle():
        world = _world(2, 1, 5)
        desc = P.describe_world(world)
        # 5 agent-landmark pairs; static landmarks never collide with each other (ref core.py:2791-2794)
        assert len(desc.items) == 5
        assert all(it["kind"] == P.K_SS for it in desc.items)


def test_entity_limit_is_reported():
    assert _native.lane_layout(128) == (32, 4)
    with pytest.raises(NotImplementedError, match="128"):
        _native.lane_layout(129)


def test_reset_before_any_step_and_with_every_selector():
    with use_oracle():
        world = _world(4, 2, 1)
        for selector in (None, 2, torch.tensor([True, False, False, True])):
            world.slab.pos.fill_(1.0)
            world.reset(selector)
            rows = [0, 1, 2, 3] if selector is None else ([2] if isinstance(selector, int) else [0, 3])
            for i in range(4):
                assert bool((world.slab.pos[i] == 0).all()) == (i in rows)
        assert world.reset_count.tolist() == [2, 1, 2, 2]
