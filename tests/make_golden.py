"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/make_golden.py

For every scenario below the reference is rolled out on CPU with seeded random actions and,
per step, the exact inputs of ``World.step`` (state slab incl. the processed action forces,
per-env joint rotations) and its outputs are recorded, together with the world description
(``plan.describe_world`` of the *reference* world), LIDAR measurements and a sample of
distance / overlap queries.  The fixtures travel to the GPU box; the reference does not.
"""
import itertools
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from refutil import import_reference, per_env_fixed_rotations, post_step, pre_step, world_state  # noqa: E402

from vectorizedmultiagentsimulator_b200.simulator import plan as P  # noqa: E402

# name, kwargs, num_envs, steps
CASES = [
    ("balance", dict(n_agents=4), 64, 100),  # BASELINE.json configs[0] (PR1 reference case)
    ("transport", dict(n_agents=4), 32, 25),
    ("navigation", dict(n_agents=8), 32, 25),
    ("flocking", dict(n_agents=5), 32, 25),
    ("pollock", dict(lidar=True), 8, 12),
    ("waterfall", dict(), 16, 20),
    ("reverse_transport", dict(), 16, 20),
    ("joint_passage", dict(), 16, 20),
    ("multi_give_way", dict(), 16, 20),
    ("give_way", dict(), 16, 20),
    ("wheel", dict(), 16, 20),
    ("dropout", dict(), 16, 15),
    ("wind_flocking", dict(), 16, 15),  # per-env gravity tensors (Entity.gravity as [B, 2])
    ("football", dict(), 8, 10),
    ("passage", dict(), 16, 15),
    # crafted worlds (tests/crafted.py) for the branches no reference scenario takes
    ("crafted_clamps", dict(), 33, 12),  # max_f / f_range / max_t / t_range, angular + linear friction
    ("crafted_joints_apart", dict(), 16, 4),  # joints with anchors clearly apart (strict-tolerance joint test)
    ("crafted_lonely", dict(), 1, 6),  # batch_dim = 1, no work item
    ("crafted_crowd", dict(), 4, 6),  # 70 entities
]


def record(vmas, name, kwargs, num_envs, steps):
    scenario = name
    if name.startswith("crafted_"):
        import crafted

        scenario = crafted.make_scenario("vmas", name[len("crafted_"):])
    env = vmas.make_env(scenario, num_envs=num_envs, device="cpu", seed=0, **kwargs)
    world = env.world
    desc = P.describe_world(world)
    fix = dict(name=name, kwargs=kwargs, desc=desc.to_json(), steps=[], lidar=[], queries=[])
    gen = torch.Generator().manual_seed(1)
    idx = {id(e): i for i, e in enumerate(world.entities)}
    prev_out = None
    for t in range(steps):
        actions = [
            (torch.rand(num_envs, a.action_size, generator=gen) * 2 - 1) * a.action.u_range_tensor
            for a in env.agents
        ]
        pre_step(env, actions)
        state_in = world_state(world)
        fixed = per_env_fixed_rotations(world, desc)
        gravity = {
            i: e.gravity.clone() for i, e in enumerate(world.entities) if desc.entities[i].get("gravity_per_env")
        }
        world.step()
        state_out = world_state(world)
        entry = dict(force=state_in["force"], torque=state_in["torque"], out=state_out)
        same = prev_out is not None and all(
            torch.equal(state_in[k], prev_out[k]) for k in ("pos", "vel", "rot", "ang_vel")
        )
        if not same:
            entry["state_in"] = {k: state_in[k] for k in ("pos", "vel", "rot", "ang_vel")}
        if fixed:
            entry["fixed_rot"] = fixed
        if gravity:
            entry["ent_gravity"] = gravity
        fix["steps"].append(entry)
        prev_out = state_out
        obs, rews, dones, infos = post_step(env)
        if t < 3 or t == steps - 1:
            entry["obs"] = [o.clone() for o in obs]
            entry["rews"] = [r.clone() for r in rews]
            entry["dones"] = dones.clone()
        # LIDAR: every sensor of every agent on the post-step state
        if t % 3 == 0:
            for a in world.agents:
                for s in a.sensors:
                    targets = [i for i, e in enumerate(world.entities) if e is not a and s.entity_filter(e)]
                    fix["lidar"].append(
                        dict(
                            step=t,
                            src=idx[id(a)],
                            targets=targets,
                            angles=s._angles.clone(),
                            max_range=float(s._max_range),
                            out=s.measure().clone(),
                        )
                    )
    # distance / overlap queries on the final state
    ents = world.entities
    rnd = random.Random(0)
    pairs = list(itertools.permutations(range(len(ents)), 2))
    rnd.shuffle(pairs)
    final = world_state(world)
    fix["final_state"] = final
    for a, b in pairs[:60]:
        pt = torch.randn(num_envs, 2, generator=gen)
        fix["queries"].append(
            dict(
                a=a,
                b=b,
                distance=world.get_distance(ents[a], ents[b]).clone(),
                overlap=world.is_overlapping(ents[a], ents[b]).clone(),
                point=pt,
                point_distance=world.get_distance_from_point(ents[a], pt).clone(),
            )
        )
    return fix


def main():
    vmas = import_reference()
    out_dir = os.path.join(HERE, "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, kwargs, num_envs, steps in CASES:
        path = os.path.join(out_dir, f"{name}.pt")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue  # fixtures are append-only; pass --all to regenerate everything
        fix = record(vmas, name, kwargs, num_envs, steps)
        torch.save(fix, path)
        print(f"{name:20s} B={num_envs:3d} T={steps:3d} lidar={len(fix['lidar']):3d} -> {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
