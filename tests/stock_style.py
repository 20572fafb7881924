"""A scenario written the way the reference's stock scenario files are written (TEST INFRASTRUCTURE):
per-agent callbacks that call ``world.is_overlapping`` / ``world.get_distance`` / ``Lidar.measure`` one
pair and one sensor at a time and assemble observations with ``torch.cat`` — no batched primitive of
this package, no observation plan.  It is the shape of third-party scenario code the drop-in boundary
has to serve (cf. ref scenarios/balance.py:216-263, navigation.py:203-265); tests step it on the CUDA
backend and on the CPU oracle and compare.
"""
import importlib

import torch


def make_scenario(root="vectorizedmultiagentsimulator_b200"):
    """The scenario built from the reference's modules (``root="vmas"``) or from this package's."""
    core = importlib.import_module(f"{root}.simulator.core")
    Agent, Box, Landmark, Line, Sphere, World = core.Agent, core.Box, core.Landmark, core.Line, core.Sphere, core.World
    BaseScenario = importlib.import_module(f"{root}.simulator.scenario").BaseScenario
    Lidar = importlib.import_module(f"{root}.simulator.sensors").Lidar
    return _build(Agent, Box, Landmark, Line, Sphere, World, BaseScenario, Lidar)()


def _build(Agent, Box, Landmark, Line, Sphere, World, BaseScenario, Lidar):
    class Scenario(BaseScenario):
        def make_world(self, batch_dim, device, **kwargs):
            self.n_agents = kwargs.pop("n_agents", 3)
            self.collision_penalty = kwargs.pop("collision_penalty", -0.5)
            world = World(batch_dim, device, substeps=2, x_semidim=1.2, y_semidim=1.2, drag=0.2)
            sees_obstacles = lambda e: e.name.startswith(("wall", "rod", "crate"))  # noqa: E731
            for i in range(self.n_agents):
                agent = Agent(
                    name=f"agent_{i}",
                    shape=Sphere(0.06),
                    u_multiplier=0.8,
                    sensors=[Lidar(world, n_rays=10, max_range=0.5, entity_filter=sees_obstacles)],
                )
                world.add_agent(agent)
            self.crate = Landmark("crate", shape=Box(0.2, 0.14), movable=True, rotatable=True, collide=True, mass=2.0)
            self.goal = Landmark("goal", shape=Sphere(0.1), collide=False)
            self.walls = [Landmark(f"wall_{k}", shape=Box(0.5, 0.08), collide=True) for k in range(2)]
            self.rod = Landmark("rod", shape=Line(0.4), movable=True, rotatable=True, collide=True)
            for lm in [self.crate, self.goal, self.rod] + self.walls:
                world.add_landmark(lm)
            return world

        def reset_world_at(self, env_index=None):
            world = self.world
            gen = torch.Generator().manual_seed(42 if env_index is None else 43 + int(env_index))
            n = world.batch_dim

            def place(entity, spread, rotate=False):
                pos = ((torch.rand(n, 2, generator=gen) * 2 - 1) * spread).to(world.device)
                entity.set_pos(pos if env_index is None else pos[env_index], batch_index=env_index)
                if rotate:
                    rot = ((torch.rand(n, 1, generator=gen) * 2 - 1) * 3.0).to(world.device)
                    entity.set_rot(rot if env_index is None else rot[env_index], batch_index=env_index)

            for agent in world.agents:
                place(agent, 0.5)
            place(self.crate, 0.4, rotate=True)
            place(self.goal, 0.9)
            place(self.rod, 0.5, rotate=True)
            for wall in self.walls:
                place(wall, 0.8, rotate=True)

        def reward(self, agent):
            world = self.world
            if agent is world.agents[0]:
                self.crate_dist = world.get_distance(self.crate, self.goal)
                self.shared = -self.crate_dist
            rew = self.shared.clone()
            for other in world.agents:
                if other is not agent:
                    rew[world.is_overlapping(agent, other)] += self.collision_penalty
            for wall in self.walls:
                rew[world.is_overlapping(agent, wall)] += self.collision_penalty
            rew += -0.1 * world.get_distance(agent, self.rod)
            return rew

        def observation(self, agent):
            lidar = agent.sensors[0].measure()
            return torch.cat(
                [
                    agent.state.pos,
                    agent.state.vel,
                    self.crate.state.pos - agent.state.pos,
                    self.crate.state.rot % torch.pi,
                    self.goal.state.pos - self.crate.state.pos,
                    lidar,
                ],
                dim=-1,
            )

        def done(self):
            return self.world.is_overlapping(self.crate, self.goal)

        def info(self, agent):
            return {"crate_dist": self.crate_dist, "rod_dist": self.world.get_distance(agent, self.rod)}

    return Scenario
