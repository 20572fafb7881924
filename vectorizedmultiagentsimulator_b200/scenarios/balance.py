"""``balance``: agents under a bar carry a package to a goal against gravity.

Task definition of the reference's ``vmas/scenarios/balance.py`` (world :17-84, reset :86-213,
reward :220-239, observation :241-257, done :259-263) re-written on the public API for the
B200 build: same entities, constants, random-draw order, observation layout and reward, but
no host synchronisation in ``reward`` (masked assignment → ``torch.where``) and the three
overlap tests are single kernel launches.
"""
import torch

from ..simulator.core import Agent, Box, Landmark, Line, Sphere, World
from ..simulator.scenario import BaseScenario
from ..simulator.utils import Color, ScenarioUtils


class Scenario(BaseScenario):
    def make_world(self, batch_dim: int, device: torch.device, **kwargs):
        self.n_agents = kwargs.pop("n_agents", 3)
        self.package_mass = kwargs.pop("package_mass", 5)
        self.random_package_pos_on_line = kwargs.pop("random_package_pos_on_line", True)
        ScenarioUtils.check_kwargs_consumed(kwargs)
        assert self.n_agents > 1

        self.line_length = 0.8
        self.agent_radius = 0.03
        self.shaping_factor = 100
        self.fall_reward = -10
        self.visualize_semidims = False

        world = World(batch_dim, device, gravity=(0.0, -0.05), y_semidim=1)
        for i in range(self.n_agents):
            world.add_agent(
                Agent(name=f"agent_{i}", shape=Sphere(self.agent_radius), u_multiplier=0.7)
            )

        goal = Landmark(name="goal", collide=False, shape=Sphere(), color=Color.LIGHT_GREEN)
        world.add_landmark(goal)
        self.package = Landmark(
            name="package",
            collide=True,
            movable=True,
            shape=Sphere(),
            mass=self.package_mass,
            color=Color.RED,
        )
        self.package.goal = goal
        world.add_landmark(self.package)
        self.line = Landmark(
            name="line",
            shape=Line(length=self.line_length),
            collide=True,
            movable=True,
            rotatable=True,
            mass=5,
            color=Color.BLACK,
        )
        world.add_landmark(self.line)
        self.floor = Landmark(
            name="floor", collide=True, shape=Box(length=10, width=1), color=Color.WHITE
        )
        world.add_landmark(self.floor)

        self.pos_rew = torch.zeros(batch_dim, device=device, dtype=torch.float32)
        self.ground_rew = self.pos_rew.clone()
        return world

    # -- reset ---------------------------------------------------------------------------
    def _uniform(self, n, low, high):
        return torch.zeros((n, 1), device=self.world.device, dtype=torch.float32).uniform_(low, high)

    def reset_world_at(self, env_index: int = None):
        world = self.world
        n = 1 if env_index is not None else world.batch_dim
        half = self.line_length / 2
        r_pkg = self.package.shape.radius
        dev = dict(device=world.device, dtype=torch.float32)

        # draw order matters for seed-for-seed equality with the reference
        goal_pos = torch.cat([self._uniform(n, -1.0, 1.0), self._uniform(n, 0.0, world.y_semidim)], dim=1)
        line_pos = torch.cat(
            [
                self._uniform(n, -1.0 + half, 1.0 - half),
                torch.full((n, 1), -world.y_semidim + self.agent_radius * 2, **dev),
            ],
            dim=1,
        )
        spread = (-half + r_pkg, half - r_pkg) if self.random_package_pos_on_line else (0.0, 0.0)
        package_rel = torch.cat([self._uniform(n, *spread), torch.full((n, 1), r_pkg, **dev)], dim=1)

        span = self.line_length - self.agent_radius
        for i, agent in enumerate(world.agents):
            offset = torch.tensor(
                [-span / 2 + i * span / (self.n_agents - 1), -self.agent_radius * 2], **dev
            )
            agent.set_pos(line_pos + offset, batch_index=env_index)

        self.line.set_pos(line_pos, batch_index=env_index)
        self.package.goal.set_pos(goal_pos, batch_index=env_index)
        self.line.set_rot(torch.zeros(1, **dev), batch_index=env_index)
        self.package.set_pos(line_pos + package_rel, batch_index=env_index)
        self.floor.set_pos(
            torch.tensor(
                [0, -world.y_semidim - self.floor.shape.width / 2 - self.agent_radius],
                device=world.device,
            ),
            batch_index=env_index,
        )
        self.compute_on_the_ground()
        dist = torch.linalg.vector_norm(self.package.state.pos - self.package.goal.state.pos, dim=1)
        self.keep(self, "global_shaping", dist * self.shaping_factor, env_index)

    def compute_on_the_ground(self):
        self.on_the_ground = self.world.is_overlapping(self.line, self.floor) + self.world.is_overlapping(
            self.package, self.floor
        )

    # -- per-step callbacks --------------------------------------------------------------------
    def reward(self, agent: Agent):
        if agent is self.world.agents[0]:
            self.compute_on_the_ground()
            self.package_dist = torch.linalg.vector_norm(
                self.package.state.pos - self.package.goal.state.pos, dim=1
            )
            self.ground_rew = torch.where(
                self.on_the_ground, float(self.fall_reward), 0.0
            ).to(torch.float32)
            shaping = self.package_dist * self.shaping_factor
            self.pos_rew = self.global_shaping - shaping
            self.keep(self, "global_shaping", shaping)  # carried to the next step: in place
        return self.ground_rew + self.pos_rew

    def _observe_all(self):
        """Observations of every agent in one pass over the state slab -> ``[A, B, 16]``.

        Same elementwise arithmetic as a per-agent ``torch.cat`` of the nine terms, but the
        agent-independent terms are computed once and the agent-relative ones for all agents at
        once (the agents are consecutive rows of the slab).  Agent-major output: each agent's
        ``[B, 16]`` observation is a contiguous slice.
        """
        world = self.world
        slab = world.slab
        ents = world.entities
        a0, n = ents.index(world.agents[0]), len(world.agents)
        ip, il, ig = ents.index(self.package), ents.index(self.line), ents.index(self.package.goal)
        apos = slab.pos[:, a0 : a0 + n].transpose(0, 1)  # [A, B, 2] views of the slab
        avel = slab.vel[:, a0 : a0 + n].transpose(0, 1)
        pkg_pos, line_pos = slab.pos[:, ip].unsqueeze(0), slab.pos[:, il].unsqueeze(0)  # [1, B, 2]
        shared = torch.cat(
            [
                pkg_pos - slab.pos[:, ig].unsqueeze(0),
                slab.vel[:, ip].unsqueeze(0),
                slab.vel[:, il].unsqueeze(0),
                slab.ang_vel[:, il : il + 1].unsqueeze(0),
                (slab.rot[:, il : il + 1] % torch.pi).unsqueeze(0),
            ],
            dim=-1,
        )
        return torch.cat(
            [apos, avel, apos - pkg_pos, apos - line_pos, shared.expand(n, -1, -1)], dim=-1
        )

    def observation(self, agent: Agent):
        agents = self.world.agents
        if agent is agents[0] or getattr(self, "_obs_all", None) is None:
            self._obs_all = self._observe_all()
        return self._obs_all[agents.index(agent)]

    def done(self):
        return self.on_the_ground + self.world.is_overlapping(self.package, self.package.goal)

    def info(self, agent: Agent):
        return {"pos_rew": self.pos_rew, "ground_rew": self.ground_rew}
