"""``flocking``: agents keep formation around a scripted target among static obstacles.

Task definition of the reference's ``vmas/scenarios/flocking.py`` (world :18-83, scripted
target :85-90, reset :92-150, reward :152-191, observation :193-202, info :204-210) re-written
on the public API, without host synchronisation in the reward.
"""
from typing import Dict

import torch
from torch import Tensor

from ..simulator.core import Agent, Landmark, Sphere, World
from ..simulator.scenario import BaseScenario
from ..simulator.sensors import Lidar
from ..simulator.utils import Color, ScenarioUtils, Y


class Scenario(BaseScenario):
    def make_world(self, batch_dim: int, device: torch.device, **kwargs):
        n_agents = kwargs.pop("n_agents", 4)
        n_obstacles = kwargs.pop("n_obstacles", 5)
        self._min_dist_between_entities = kwargs.pop("min_dist_between_entities", 0.15)
        self.n_lidar_rays = kwargs.pop("n_lidar_rays", 12)
        self.collision_reward = kwargs.pop("collision_reward", -0.1)
        self.dist_shaping_factor = kwargs.pop("dist_shaping_factor", 1)
        ScenarioUtils.check_kwargs_consumed(kwargs)

        self.plot_grid = True
        self.desired_distance = 0.1
        self.min_collision_distance = 0.005
        self.x_dim = 1
        self.y_dim = 1

        # the scripted target's action is in range by construction: skip the sync-ing assert
        world = World(batch_dim, device, collision_force=400, substeps=5, check_scripted_actions=False)
        self._target = Agent(
            name="target",
            collide=True,
            color=Color.GREEN,
            render_action=True,
            action_script=self._circle_script(),
        )
        world.add_agent(self._target)
        sees_landmarks = lambda e: not isinstance(e, Agent)  # noqa: E731
        for i in range(n_agents):
            agent = Agent(
                name=f"agent_{i}",
                collide=True,
                sensors=[Lidar(world, n_rays=self.n_lidar_rays, max_range=0.2, entity_filter=sees_landmarks)],
                render_action=True,
            )
            agent.collision_rew = torch.zeros(batch_dim, device=device)
            agent.dist_rew = agent.collision_rew.clone()
            world.add_agent(agent)

        self.obstacles = []
        for i in range(n_obstacles):
            obstacle = Landmark(
                name=f"obstacle_{i}", collide=True, movable=False, shape=Sphere(radius=0.1), color=Color.RED
            )
            world.add_landmark(obstacle)
            self.obstacles.append(obstacle)
        return world

    def _circle_script(self):
        def script(agent, world):
            phase = self.t / 30
            agent.action.u = torch.stack([torch.cos(phase), torch.sin(phase)], dim=1)

        return script

    def _separation_cost(self, agent: Agent, rows=None):
        """mean over the other agents of (distance - desired)^2, times the shaping factor."""
        dists = [
            torch.linalg.vector_norm(agent.state.pos - other.state.pos, dim=-1)
            for other in self.world.agents
            if other is not agent
        ]
        stacked = torch.stack(dists, dim=1)
        return (stacked - self.desired_distance).pow(2).mean(-1) * self.dist_shaping_factor

    def reset_world_at(self, env_index: int = None):
        world = self.world
        n = 1 if env_index is not None else world.batch_dim
        target_pos = torch.zeros((n, world.dim_p), device=world.device, dtype=torch.float32)
        target_pos[:, Y] = -self.y_dim
        self._target.set_pos(target_pos, batch_index=env_index)
        ScenarioUtils.spawn_entities_randomly(
            self.obstacles + world.policy_agents,
            world,
            env_index,
            self._min_dist_between_entities,
            x_bounds=(-self.x_dim, self.x_dim),
            y_bounds=(-self.y_dim, self.y_dim),
            occupied_positions=target_pos.unsqueeze(1),
        )
        for agent in world.policy_agents:
            self.keep(agent, "distance_shaping", self._separation_cost(agent), env_index)
        self.keep(self, "t", torch.zeros(world.batch_dim, device=world.device), env_index)

    def reward(self, agent: Agent):
        world = self.world
        if world.policy_agents.index(agent) == 0:
            self.t += 1
            if self.collision_reward != 0:
                for a in world.policy_agents:
                    a.collision_rew = torch.zeros_like(a.collision_rew)
                agents = world.agents
                for i, a in enumerate(agents):
                    for b in agents[i + 1 :]:
                        touching = world.get_distance(a, b) <= self.min_collision_distance
                        penalty = torch.where(touching, float(self.collision_reward), 0.0)
                        if a.action_script is None:
                            a.collision_rew = a.collision_rew + penalty
                        if b.action_script is None:
                            b.collision_rew = b.collision_rew + penalty
        cost = self._separation_cost(agent)
        agent.dist_rew = agent.distance_shaping - cost
        self.keep(agent, "distance_shaping", cost)
        return agent.collision_rew + agent.dist_rew

    def observation(self, agent: Agent):
        return torch.cat(
            [
                agent.state.pos,
                agent.state.vel,
                agent.state.pos - self._target.state.pos,
                agent.sensors[0].measure(),
            ],
            dim=-1,
        )

    def info(self, agent: Agent) -> Dict[str, Tensor]:
        return {"agent_collision_rew": agent.collision_rew, "agent_distance_rew": agent.dist_rew}
