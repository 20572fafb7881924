"""Crafted worlds for the branches no reference scenario exercises (TEST INFRASTRUCTURE).

``max_f`` / ``max_t`` / ``t_range`` clamps of the action force and torque (ref core.py:2018-2041),
angular friction (ref core.py:2054-2102), joints whose anchors are clearly apart (ref
core.py:2201-2292: the joint arithmetic away from the |delta| -> 0 amplification of a fresh
reset), a one-env batch, a world without any work item, and a world beyond 64 entities.

The scenario classes are built from whichever namespace they are given — the UNMODIFIED
reference's ``vmas`` (``tests/make_golden.py`` and ``tests/test_oracle_vs_reference.py`` use them
to produce fixtures and to pin the oracle) or this package's ``simulator`` modules.
"""
import importlib
import math

import torch


def _ns(root):
    """Module namespace of the reference (``root='vmas'``) or of this package."""
    mod = lambda name: importlib.import_module(f"{root}.simulator.{name}")  # noqa: E731
    core = mod("core")
    return dict(
        Agent=core.Agent, Landmark=core.Landmark, World=core.World, Sphere=core.Sphere, Box=core.Box,
        Line=core.Line, BaseScenario=mod("scenario").BaseScenario, Joint=mod("joints").Joint,
        HolonomicWithRotation=mod("dynamics.holonomic_with_rot").HolonomicWithRotation,
        DiffDrive=mod("dynamics.diff_drive").DiffDrive, KinematicBicycle=mod("dynamics.kinematic_bicycle").KinematicBicycle,
        Drone=mod("dynamics.drone").Drone, Forward=mod("dynamics.forward").Forward,
        Rotation=mod("dynamics.roatation").Rotation, Static=mod("dynamics.static").Static,
    )


def _scatter(world, entities, gen, spread, env_index, rot_spread=math.pi):
    """Deterministic random placement (own CPU generator: the same layout in either namespace)."""
    n = world.batch_dim
    for e in entities:
        pos = (torch.rand(n, 2, generator=gen) * 2 - 1) * spread
        rot = (torch.rand(n, 1, generator=gen) * 2 - 1) * rot_spread
        e.set_pos(pos.to(world.device) if env_index is None else pos[env_index].to(world.device), batch_index=env_index)
        e.set_rot(rot.to(world.device) if env_index is None else rot[env_index].to(world.device), batch_index=env_index)


def make_scenario(root, kind, seed=1234):
    ns = _ns(root)
    Agent, Landmark, World = ns["Agent"], ns["Landmark"], ns["World"]
    Sphere, Box, Line, Joint = ns["Sphere"], ns["Box"], ns["Line"], ns["Joint"]
    Rot = ns["HolonomicWithRotation"]

    class Crafted(ns["BaseScenario"]):
        def make_world(self, batch_dim, device, **kwargs):
            self.kind = kind
            self.gen = torch.Generator().manual_seed(seed)
            if kind == "clamps":
                # every action clamp, both frictions (world-wide and per entity), speed limits, contacts
                world = World(
                    batch_dim, device, substeps=2, drag=0.2, linear_friction=0.15, angular_friction=0.1,
                    x_semidim=0.6, y_semidim=0.6,
                )
                specs = [
                    dict(shape=Sphere(0.06), max_f=0.6, max_t=0.004, u_multiplier=[1.0, 1.0, 0.01]),
                    dict(shape=Sphere(0.05), f_range=0.4, t_range=0.003, u_multiplier=[1.0, 1.0, 0.01]),
                    dict(shape=Box(0.16, 0.08), max_f=0.9, f_range=0.7, max_t=0.02, t_range=0.015,
                         u_multiplier=[1.2, 1.2, 0.03], max_speed=0.35, angular_friction=0.3),
                    dict(shape=Line(0.2), v_range=0.25, t_range=0.01, u_multiplier=[1.0, 1.0, 0.02],
                         linear_friction=0.05),
                ]
                for i, s in enumerate(specs):
                    world.add_agent(Agent(name=f"agent_{i}", rotatable=True, dynamics=Rot(), **s))
                world.add_landmark(Landmark("bar", shape=Line(0.3), movable=True, rotatable=True, collide=True, mass=2.0))
                world.add_landmark(Landmark("crate", shape=Box(0.12, 0.12), movable=True, rotatable=True, collide=True,
                                            mass=3.0, angular_friction=0.2))
                self.spread = 0.3
            elif kind == "joints_apart":
                # joints stretched / compressed well beyond the contact margin, all three rotation modes
                world = World(batch_dim, device, substeps=3, drag=0.25, joint_force=8, torque_constraint_force=0.02)
                for i in range(4):
                    world.add_agent(Agent(name=f"agent_{i}", shape=Sphere(0.04), rotatable=True, dynamics=Rot(),
                                          u_multiplier=[1.0, 1.0, 0.01]))
                a = world.agents
                world.add_landmark(Landmark("beam", shape=Line(0.4), movable=True, rotatable=True, collide=True, mass=1.5))
                beam = world.landmarks[0]
                self.joints = [
                    Joint(a[0], a[1], anchor_a=(0, 0), anchor_b=(0, 0), dist=0.25, rotate_a=True, rotate_b=True,
                          collidable=False, width=0, mass=1),
                    Joint(a[2], beam, anchor_a=(0, 0), anchor_b=(-1, 0), dist=0.1, rotate_a=False, rotate_b=True,
                          fixed_rotation_a=0.3, collidable=True, width=0.01, mass=1),
                    Joint(a[3], beam, anchor_a=(0, 0), anchor_b=(1, 0), dist=0.0, rotate_a=False, rotate_b=False,
                          fixed_rotation_a=0.5, fixed_rotation_b=0.5),
                ]
                for j in self.joints:
                    world.add_joint(j)
                self.spread = 0.35
            elif kind == "dynamics_zoo":
                # one agent per action -> force / torque model, no contacts in the way
                world = World(batch_dim, device, substeps=1, drag=0.25)
                zoo = [
                    ("diff_rk4", ns["DiffDrive"](world, integration="rk4"), Sphere(0.05), dict(u_range=[1.0, 2.0], u_multiplier=[0.6, 1.0])),
                    ("diff_euler", ns["DiffDrive"](world, integration="euler"), Box(0.12, 0.08), dict(u_range=[1.0, 2.0])),
                    ("bicycle", ns["KinematicBicycle"](world, width=0.08, l_f=0.06, l_r=0.05, max_steering_angle=0.6),
                     Box(0.12, 0.08), dict(u_range=[1.0, 1.0], u_multiplier=[0.8, 1.0])),
                    ("bicycle_euler", ns["KinematicBicycle"](world, width=0.08, l_f=0.05, l_r=0.07, max_steering_angle=0.4,
                                                             integration="euler"), Sphere(0.05), dict(u_range=[1.0, 1.0])),
                    ("drone", ns["Drone"](world), Sphere(0.05), dict(u_range=[0.0001, 0.0001, 0.0001, 0.0001], mass=0.1)),
                    ("forward", ns["Forward"](), Sphere(0.05), dict(u_range=[1.0])),
                    ("rotation", ns["Rotation"](), Box(0.1, 0.06), dict(u_range=[0.02])),
                    ("holo_rot", Rot(), Sphere(0.05), dict(u_range=[1.0, 1.0, 0.01])),
                    ("static", ns["Static"](), Sphere(0.05), dict()),
                ]
                for name, dyn, shape, kw in zoo:
                    world.add_agent(Agent(name=name, shape=shape, rotatable=True, collide=False, dynamics=dyn,
                                          action_size=dyn.needed_action_size, **kw))
                world.add_agent(Agent(name="holo", shape=Sphere(0.05), collide=False))
                self.spread = 0.8
            elif kind == "lonely":
                # no work item at all: one non-colliding agent
                world = World(batch_dim, device, drag=0.25, angular_friction=0.05)
                world.add_agent(Agent(name="agent_0", shape=Sphere(0.05), collide=False, rotatable=True, dynamics=Rot(),
                                      max_t=0.01, u_multiplier=[1.0, 1.0, 0.02]))
                self.spread = 0.5
            elif kind == "crowd":
                # 70 entities (beyond the 64 a warp's lanes cover with one entity each), few of them colliding
                world = World(batch_dim, device, substeps=1, drag=0.25, x_semidim=1.0, y_semidim=1.0)
                for i in range(6):
                    world.add_agent(Agent(name=f"agent_{i}", shape=Sphere(0.05), max_f=0.8))
                for i in range(64):
                    shape = Sphere(0.03) if i % 3 else Box(0.08, 0.05)
                    world.add_landmark(Landmark(f"lm_{i}", shape=shape, collide=i < 6, movable=i < 3, rotatable=i < 3))
                self.spread = 0.25
            else:
                raise ValueError(kind)
            return world

        def reset_world_at(self, env_index=None):
            world = self.world
            movable = list(world.entities)  # a joint's own landmark included: its anchors end up apart too
            # joints: modest angles (the rotation constraint grows like exp(|angle difference|) and throws
            # bodies to infinity when started half a turn apart), anchors far apart
            _scatter(world, movable, self.gen, self.spread, env_index, 0.4 if kind == "joints_apart" else math.pi)
            n = world.batch_dim
            for e in movable:
                if e.movable:
                    v = (torch.rand(n, 2, generator=self.gen) * 2 - 1) * 0.4
                    e.set_vel(v.to(world.device) if env_index is None else v[env_index].to(world.device), batch_index=env_index)
                if e.rotatable:
                    w = (torch.rand(n, 1, generator=self.gen) * 2 - 1) * (0.3 if kind == "joints_apart" else 1.5)
                    e.set_ang_vel(w.to(world.device) if env_index is None else w[env_index].to(world.device), batch_index=env_index)

        def reward(self, agent):
            return torch.zeros(self.world.batch_dim, device=self.world.device)

        def observation(self, agent):
            return torch.cat([agent.state.pos, agent.state.vel], dim=-1)

    return Crafted()


KINDS = ("clamps", "joints_apart", "lonely", "crowd", "dynamics_zoo")
