"""``transport``: agents push heavy box packages onto a goal.

Task definition of the reference's ``vmas/scenarios/transport.py`` (world :16-67, reset :69-127,
reward :129-166, observation :168-185, done :187-194) re-written on the public API.  Two extra
kwargs build the BASELINE.json variant that the reference does not ship
(``n_lines`` movable/rotatable Line landmarks, ``substeps``); with their defaults the world is
the stock one.
"""
import torch

from ..simulator import observe as O
from ..simulator.core import Agent, Box, Landmark, Line, Sphere, World
from ..simulator.scenario import BaseScenario
from ..simulator.utils import Color, ScenarioUtils


class Scenario(BaseScenario):
    supports_masked_reset = True  # reset_world_at(env_index): None, an int, or a [B] bool mask

    def make_world(self, batch_dim: int, device: torch.device, **kwargs):
        self._obs_plan = self._obs_all = self._shaping_block = self._zero = None
        n_agents = kwargs.pop("n_agents", 4)
        self.n_packages = kwargs.pop("n_packages", 1)
        self.package_width = kwargs.pop("package_width", 0.15)
        self.package_length = kwargs.pop("package_length", 0.15)
        self.package_mass = kwargs.pop("package_mass", 50)
        self.n_lines = kwargs.pop("n_lines", 0)  # B200-bench variant only
        self.line_length = kwargs.pop("line_length", 0.3)
        substeps = kwargs.pop("substeps", 1)
        ScenarioUtils.check_kwargs_consumed(kwargs)

        self.shaping_factor = 100
        self.world_semidim = 1
        self.agent_radius = 0.03

        semidim = self.world_semidim + 2 * self.agent_radius + max(self.package_length, self.package_width)
        world = World(batch_dim, device, x_semidim=semidim, y_semidim=semidim, substeps=substeps)
        for i in range(n_agents):
            world.add_agent(Agent(name=f"agent_{i}", shape=Sphere(self.agent_radius), u_multiplier=0.6))
        goal = Landmark(name="goal", collide=False, shape=Sphere(radius=0.15), color=Color.LIGHT_GREEN)
        world.add_landmark(goal)
        self.packages = []
        for i in range(self.n_packages):
            package = Landmark(
                name=f"package {i}",
                collide=True,
                movable=True,
                mass=self.package_mass,
                shape=Box(length=self.package_length, width=self.package_width),
                color=Color.RED,
            )
            package.goal = goal
            self.packages.append(package)
            world.add_landmark(package)
        self.lines = []
        for i in range(self.n_lines):
            line = Landmark(
                name=f"line {i}",
                collide=True,
                movable=True,
                rotatable=True,
                shape=Line(length=self.line_length),
                color=Color.BLACK,
            )
            self.lines.append(line)
            world.add_landmark(line)
        return world

    def reset_world_at(self, env_index: int = None):
        world = self.world
        bounds = (-self.world_semidim, self.world_semidim)
        ScenarioUtils.spawn_entities_randomly(
            world.agents,
            world,
            env_index,
            min_dist_between_entities=self.agent_radius * 2,
            x_bounds=bounds,
            y_bounds=bounds,
        )
        occupied = torch.stack([a.state.pos for a in world.agents], dim=1)
        if isinstance(env_index, int):  # None / bool mask: a row per env
            occupied = occupied[env_index].unsqueeze(0)
        goal = world.landmarks[0]
        ScenarioUtils.spawn_entities_randomly(
            [goal] + self.packages + self.lines,
            world,
            env_index,
            min_dist_between_entities=max(
                p.shape.circumscribed_radius() + goal.shape.radius + 0.01 for p in self.packages
            ),
            x_bounds=bounds,
            y_bounds=bounds,
            occupied_positions=occupied,
        )
        for package in self.packages:
            package.on_goal = world.is_overlapping(package, package.goal)
            dist = torch.linalg.vector_norm(package.state.pos - package.goal.state.pos, dim=1)
            self.keep(package, "global_shaping", dist * self.shaping_factor, env_index)

    def reward(self, agent: Agent):
        if agent is self.world.agents[0]:
            rew = torch.zeros(self.world.batch_dim, device=self.world.device, dtype=torch.float32)
            if getattr(self, "_colors", None) is None:  # constants: uploaded once, not every step
                self._colors = (
                    torch.tensor(Color.RED.value, device=self.world.device, dtype=torch.float32),
                    torch.tensor(Color.GREEN.value, device=self.world.device, dtype=torch.float32),
                )
            red, green = self._colors
            block = getattr(self, "_shaping_block", None)
            if block is None:  # every package's carried shaping term as one [K, B] block
                block = self._shaping_block = torch.stack([p.global_shaping for p in self.packages])
                for k, package in enumerate(self.packages):
                    package.global_shaping = block[k]
            dist, shaped = self.world.distance_shaping(
                [(p, p.goal) for p in self.packages], self.shaping_factor, block
            )
            zero = self._zero_reward()
            for k, package in enumerate(self.packages):
                package.dist_to_goal = dist[k]
                package.on_goal = self.world.is_overlapping(package, package.goal)
                package.color = torch.where(package.on_goal.unsqueeze(-1), green, red)
                rew = rew + torch.where(package.on_goal, zero, shaped[k])
            self.rew = rew
        return self.rew

    def _zero_reward(self):
        zero = getattr(self, "_zero", None)
        if zero is None or zero.device != self.world.slab.pos.device:
            zero = self._zero = torch.tensor(0.0, dtype=torch.float32, device=self.world.slab.pos.device)
        return zero

    def observation(self, agent: Agent):
        agents = self.world.agents
        if agent is agents[0] or getattr(self, "_obs_all", None) is None:
            plan = getattr(self, "_obs_plan", None)
            if plan is None:
                self._on_goal_terms = [O.blank(1) for _ in self.packages]
                plan = self._obs_plan = O.ObservationPlan(
                    [
                        [O.pos(a), O.vel(a)]
                        + [
                            t
                            for package, flag in zip(self.packages, self._on_goal_terms)
                            for t in (O.rel_pos(package, package.goal), O.rel_pos(package, a), O.vel(package), flag)
                        ]
                        for a in agents
                    ]
                )
            block = self.world.observe(plan)  # [A, B, 4 + 7 * n_packages], one launch
            for package, flag in zip(self.packages, self._on_goal_terms):
                block[:, :, plan.column_of(0, flag)] = package.on_goal  # bool -> 0. / 1., every agent
            self._obs_all = block
        row = self._obs_all[agents.index(agent)]
        if agent is agents[-1]:
            self._obs_all = None  # one sweep over the agents per block: a later call measures anew
        return row

    def done(self):
        return torch.stack([p.on_goal for p in self.packages], dim=1).all(dim=-1)
