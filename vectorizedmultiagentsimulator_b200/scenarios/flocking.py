"""``flocking``: agents keep formation around a scripted target among static obstacles.

Task definition of the reference's ``vmas/scenarios/flocking.py`` (world :18-83, scripted
target :85-90, reset :92-150, reward :152-191, observation :193-202, info :204-210) re-written
on the public API, without host synchronisation in the reward.
"""
from typing import Dict

import torch
from torch import Tensor

from ..simulator import observe as O
from ..simulator.core import Agent, Landmark, Sphere, World
from ..simulator.scenario import BaseScenario
from ..simulator.sensors import Lidar
from ..simulator.utils import Color, ScenarioUtils, Y


class Scenario(BaseScenario):
    supports_masked_reset = True  # reset_world_at(env_index): None, an int, or a [B] bool mask
    #: observation() reads the world state only (nothing reward() / done() computed): the environment
    #: may run it on a side stream next to the reward callbacks
    observations_are_independent = True

    def make_world(self, batch_dim: int, device: torch.device, **kwargs):
        self._batch = self._obs_all = None  # caches of the batched callbacks belong to one world
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
        n = 1 if isinstance(env_index, int) else world.batch_dim  # None / bool mask: a row per env
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

    # ---- batched fast path: all policy agents at once ([A, B, ...] blocks of the state slab) -------
    def _batch_setup(self):
        world = self.world
        cache = getattr(self, "_batch", None)
        if cache is not None and cache["world"] is world and cache["version"] == world._plan_version:
            return cache
        agents, policy, ents = world.agents, world.policy_agents, world.entities
        dev = world.device
        a0 = ents.index(agents[0])
        assert [ents.index(a) for a in agents] == list(range(a0, a0 + len(agents)))
        assert agents[0] is self._target and policy == agents[1:]
        n = len(agents)
        pairs = [(agents[i], agents[j]) for i in range(n) for j in range(i + 1, n)]
        incidence = torch.zeros(len(policy), len(pairs), device=dev)
        for k, (a, b) in enumerate(pairs):
            for e in (a, b):
                if e.action_script is None:
                    incidence[policy.index(e), k] = 1.0
        pair_row = {}
        for k, (a, b) in enumerate(pairs):
            pair_row[(id(a), id(b))] = pair_row[(id(b), id(a))] = k
        # row of `pairs` holding the distance from policy agent i to every other agent, in
        # world.agents order with itself skipped  -> [P, n-1]
        others = torch.tensor(
            [[pair_row[(id(agents[i]), id(agents[j]))] for j in range(n) if j != i] for i in range(1, n)], device=dev
        )
        cache = dict(
            world=world,
            version=world._plan_version,
            a0=a0,
            n=n,
            pairs=pairs,
            incidence=incidence,
            others=others,
            sensors=[a.sensors[0] for a in policy],
            shaping=torch.stack([a.distance_shaping for a in policy]),
        )
        for i, a in enumerate(policy):
            a.distance_shaping = cache["shaping"][i]
        self._batch = cache
        return cache

    def reward(self, agent: Agent):
        world = self.world
        policy = world.policy_agents
        c = self._batch_setup()
        if agent is policy[0]:
            self.t += 1
            n_env = world.batch_dim
            if self.collision_reward != 0:
                touching = world.get_distances(c["pairs"]) <= self.min_collision_distance  # [K, B]
                collision_rew = (c["incidence"] @ touching.to(torch.float32)) * float(self.collision_reward)
            else:
                collision_rew = torch.zeros(len(policy), n_env, device=world.device)
            to_others = world.get_center_distances(c["pairs"])[c["others"]]  # [P, n-1, B]
            cost = (to_others - self.desired_distance).pow(2).mean(1) * self.dist_shaping_factor
            dist_rew = c["shaping"] - cost
            c["shaping"].copy_(cost)
            for i, a in enumerate(policy):
                a.collision_rew, a.dist_rew = collision_rew[i], dist_rew[i]
        return agent.collision_rew + agent.dist_rew

    def observation(self, agent: Agent):
        world = self.world
        policy = world.policy_agents
        c = self._batch_setup()
        if agent is policy[0] or getattr(self, "_obs_all", None) is None:
            plan = c.get("obs_plan")
            if plan is None:
                plan = c["obs_plan"] = O.ObservationPlan(
                    [[O.pos(a), O.vel(a), O.rel_pos(a, self._target), O.lidar(a.sensors[0])] for a in policy]
                )
            self._obs_all = world.observe(plan)
        row = self._obs_all[policy.index(agent)]
        if agent is policy[-1]:
            self._obs_all = None  # one sweep over the agents per block: a later call measures anew
        return row

    def info(self, agent: Agent) -> Dict[str, Tensor]:
        return {"agent_collision_rew": agent.collision_rew, "agent_distance_rew": agent.dist_rew}
