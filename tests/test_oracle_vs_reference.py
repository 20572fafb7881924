"""The CPU oracle against the UNMODIFIED reference, live, in this container (no fixtures in between).

The reference is rolled out with seeded random actions — other seeds, batch sizes and scenario
arguments than the committed golden fixtures use — and at every step the oracle receives exactly what
the reference's ``World.step`` received (teacher forcing) and must return what it returned: bit for
bit for the physics, the LIDAR readings and the distance / overlap queries.  This is the pin the
``oracle/`` docstrings refer to; it needs ``/root/reference`` (``-m reference`` tests are skipped
elsewhere, where ``tests/test_oracle_golden.py`` checks the same functions against the fixtures).
"""
import itertools

import pytest
import torch

from oracle import queries as Q
from oracle import world_step as WS
from refutil import import_reference, per_env_fixed_rotations, post_step, pre_step, world_state
from vectorizedmultiagentsimulator_b200.simulator import plan as P

pytestmark = pytest.mark.reference

# name, kwargs, num_envs, steps, seed
CASES = [
    ("balance", dict(n_agents=3), 20, 30, 3),
    ("balance", dict(n_agents=5, package_mass=7), 9, 20, 4),
    ("transport", dict(n_agents=3, n_packages=2), 12, 20, 5),
    ("navigation", dict(n_agents=5), 12, 20, 6),
    ("flocking", dict(n_agents=4), 12, 15, 7),
    ("pollock", dict(lidar=True), 4, 6, 8),
    ("waterfall", dict(), 8, 12, 9),
    ("reverse_transport", dict(), 8, 12, 10),
    ("joint_passage", dict(), 6, 10, 11),
    ("wheel", dict(), 8, 12, 12),
    ("wind_flocking", dict(), 8, 10, 13),
    # crafted worlds (tests/crafted.py): action clamps, angular friction, joints with anchors apart,
    # a one-env batch without work items, 70 entities — branches no reference scenario takes
    ("crafted_clamps", dict(), 21, 15, 14),
    ("crafted_joints_apart", dict(), 10, 5, 15),
    ("crafted_lonely", dict(), 1, 5, 16),
    ("crafted_crowd", dict(), 3, 5, 17),
]
STATE = ("pos", "vel", "rot", "ang_vel")


@pytest.mark.parametrize("name,kwargs,num_envs,steps,seed", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_oracle_equals_live_reference_bit_for_bit(name, kwargs, num_envs, steps, seed):
    vmas = import_reference()
    scenario = name
    if name.startswith("crafted_"):
        import crafted

        scenario = crafted.make_scenario("vmas", name[len("crafted_"):], seed=1000 + seed)
    env = vmas.make_env(scenario, num_envs=num_envs, device="cpu", seed=seed, **kwargs)
    world = env.world
    desc = P.describe_world(world)  # the plan compiler reads the reference's own objects
    tables = P.build_tables(desc)
    ents = world.entities
    gen = torch.Generator().manual_seed(100 + seed)
    for t in range(steps):
        actions = [
            (torch.rand(num_envs, a.action_size, generator=gen) * 2 - 1) * a.action.u_range_tensor for a in env.agents
        ]
        pre_step(env, actions)
        state = world_state(world)  # what the reference's World.step is about to consume
        fixed_rot = per_env_fixed_rotations(world, desc)
        gravity = {i: e.gravity.clone() for i, e in enumerate(ents) if desc.entities[i].get("gravity_per_env")}
        world.step()
        want = world_state(world)
        WS.world_step(tables, state, fixed_rot=fixed_rot, **({"ent_gravity": gravity} if gravity else {}))
        for k in STATE:
            assert torch.equal(state[k], want[k]), f"{name} step {t}: {k} max |diff| {float((state[k] - want[k]).abs().max())}"
        post_step(env)
        if t % 4 == 0:  # LIDAR of every sensor on the post-step state
            for i, a in enumerate(ents):
                for s in getattr(a, "sensors", None) or []:
                    targets = [j for j, e in enumerate(ents) if e is not a and s.entity_filter(e)]
                    got = Q.cast_rays(
                        tables, want["pos"], want["rot"], i, targets, s._angles + want["rot"][:, i].unsqueeze(-1),
                        float(s._max_range),
                    )
                    assert torch.equal(got, s.measure()), f"{name} step {t}: lidar of entity {i}"
    # distance / overlap queries on the final state
    final = world_state(world)
    for a, b in list(itertools.permutations(range(len(ents)), 2))[:40]:
        assert torch.equal(Q.pair_distance(tables, final["pos"], final["rot"], a, b), world.get_distance(ents[a], ents[b]))
        assert torch.equal(Q.pair_overlap(tables, final["pos"], final["rot"], a, b), world.is_overlapping(ents[a], ents[b]))
        point = torch.randn(num_envs, 2, generator=gen)
        assert torch.equal(
            Q.distance_from_point(tables, final["pos"], final["rot"], a, point),
            world.get_distance_from_point(ents[a], point),
        )
