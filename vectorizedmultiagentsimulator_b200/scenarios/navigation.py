"""``navigation``: every agent drives to its own goal, sensing the others with a LIDAR.

Task definition of the reference's ``vmas/scenarios/navigation.py`` (world :23-141, reset
:143-195, reward :197-245, observation :247-266, done :268-280, info :282-287) re-written on the
public API.  Pairwise agent-collision penalties use the distance kernel for all envs at once
instead of a host-synchronising ``world.collides`` test per pair.
"""
from typing import Dict

import torch
from torch import Tensor

from ..simulator import observe as O
from ..simulator.core import Agent, Landmark, Sphere, World
from ..simulator.scenario import BaseScenario
from ..simulator.sensors import Lidar
from ..simulator.utils import ScenarioUtils

_PALETTE = [
    (0.22, 0.49, 0.72),
    (1.00, 0.50, 0),
    (0.30, 0.69, 0.29),
    (0.97, 0.51, 0.75),
    (0.60, 0.31, 0.64),
    (0.89, 0.10, 0.11),
    (0.87, 0.87, 0),
]


class Scenario(BaseScenario):
    supports_masked_reset = True  # reset_world_at(env_index): None, an int, or a [B] bool mask
    #: observation() reads the world state only (nothing reward() / done() computed): the environment
    #: may run it on a side stream next to the reward callbacks
    observations_are_independent = True

    def make_world(self, batch_dim: int, device: torch.device, **kwargs):
        self._batch = self._obs_all = None  # caches of the batched callbacks belong to one world
        self.plot_grid = False
        self.n_agents = kwargs.pop("n_agents", 4)
        self.collisions = kwargs.pop("collisions", True)
        self.world_spawning_x = kwargs.pop("world_spawning_x", 1)
        self.world_spawning_y = kwargs.pop("world_spawning_y", 1)
        self.enforce_bounds = kwargs.pop("enforce_bounds", False)
        self.agents_with_same_goal = kwargs.pop("agents_with_same_goal", 1)
        self.split_goals = kwargs.pop("split_goals", False)
        self.observe_all_goals = kwargs.pop("observe_all_goals", False)
        self.lidar_range = kwargs.pop("lidar_range", 0.35)
        self.agent_radius = kwargs.pop("agent_radius", 0.1)
        self.comms_range = kwargs.pop("comms_range", 0)
        self.n_lidar_rays = kwargs.pop("n_lidar_rays", 12)
        self.shared_rew = kwargs.pop("shared_rew", True)
        self.pos_shaping_factor = kwargs.pop("pos_shaping_factor", 1)
        self.final_reward = kwargs.pop("final_reward", 0.01)
        self.agent_collision_penalty = kwargs.pop("agent_collision_penalty", -1)
        ScenarioUtils.check_kwargs_consumed(kwargs)

        self.min_distance_between_entities = self.agent_radius * 2 + 0.05
        self.min_collision_distance = 0.005
        self.x_semidim = self.world_spawning_x if self.enforce_bounds else None
        self.y_semidim = self.world_spawning_y if self.enforce_bounds else None

        assert 1 <= self.agents_with_same_goal <= self.n_agents
        if self.agents_with_same_goal > 1:
            assert not self.collisions, "If agents share goals they cannot be collidables"
        if self.split_goals:
            assert (
                self.n_agents % 2 == 0 and self.agents_with_same_goal == self.n_agents // 2
            ), "Splitting the goals is allowed when the agents are even and half the team has the same goal"

        world = World(batch_dim, device, substeps=2, x_semidim=self.x_semidim, y_semidim=self.y_semidim)

        # the reference draws the extra colours from the env RNG here; keep the draw
        extra_colors = torch.randn((max(self.n_agents - len(_PALETTE), 0), 3), device=device)
        sees_agents = lambda e: isinstance(e, Agent)  # noqa: E731

        for i in range(self.n_agents):
            color = _PALETTE[i] if i < len(_PALETTE) else extra_colors[i - len(_PALETTE)]
            sensors = None
            if self.collisions:
                sensors = [
                    Lidar(world, n_rays=self.n_lidar_rays, max_range=self.lidar_range, entity_filter=sees_agents)
                ]
            agent = Agent(
                name=f"agent_{i}",
                collide=self.collisions,
                color=color,
                shape=Sphere(radius=self.agent_radius),
                render_action=True,
                sensors=sensors,
            )
            agent.pos_rew = torch.zeros(batch_dim, device=device)
            agent.agent_collision_rew = agent.pos_rew.clone()
            world.add_agent(agent)
            goal = Landmark(name=f"goal {i}", collide=False, color=color)
            world.add_landmark(goal)
            agent.goal = goal

        self.pos_rew = torch.zeros(batch_dim, device=device)
        self.final_rew = self.pos_rew.clone()
        return world

    def reset_world_at(self, env_index: int = None):
        world = self.world
        xb = (-self.world_spawning_x, self.world_spawning_x)
        yb = (-self.world_spawning_y, self.world_spawning_y)
        ScenarioUtils.spawn_entities_randomly(
            world.agents, world, env_index, self.min_distance_between_entities, xb, yb
        )
        occupied = torch.stack([a.state.pos for a in world.agents], dim=1)
        if isinstance(env_index, int):  # None / bool mask: a row per env
            occupied = occupied[env_index].unsqueeze(0)
        goal_positions = []
        for _ in world.agents:
            p = ScenarioUtils.find_random_pos_for_entity(
                occupied_positions=occupied,
                env_index=env_index,
                world=world,
                min_dist_between_entities=self.min_distance_between_entities,
                x_bounds=xb,
                y_bounds=yb,
            )
            goal_positions.append(p.squeeze(1))
            occupied = torch.cat([occupied, p], dim=1)

        for i, agent in enumerate(world.agents):
            if self.split_goals:
                which = int(i // self.agents_with_same_goal)
            else:
                which = 0 if i < self.agents_with_same_goal else i
            agent.goal.set_pos(goal_positions[which], batch_index=env_index)
            dist = torch.linalg.vector_norm(agent.state.pos - agent.goal.state.pos, dim=1)
            self.keep(agent, "pos_shaping", dist * self.pos_shaping_factor, env_index)

    # ---- batched fast path ---------------------------------------------------------------------
    # All agents are consecutive rows of the state slab, so the per-agent terms of reward and
    # observation are evaluated for every agent at once ([A, B, ...] tensors), the pairwise
    # collision tests and the LIDARs in one kernel launch each.  Arithmetic per element is the same
    # as in the per-agent formulation below (which remains the path for non-default options).
    def _batched_ok(self) -> bool:
        return self.collisions and not self.observe_all_goals

    def _batch_setup(self):
        world = self.world
        cache = getattr(self, "_batch", None)
        if cache is not None and cache["world"] is world and cache["version"] == world._plan_version:
            return cache
        agents, ents = world.agents, world.entities
        dev = world.device
        a0 = ents.index(agents[0])
        assert [ents.index(a) for a in agents] == list(range(a0, a0 + len(agents)))
        pairs = [(agents[i], agents[j]) for i in range(len(agents)) for j in range(i) if world.static_collides(agents[i], agents[j])]
        incidence = torch.zeros(len(agents), max(len(pairs), 1), device=dev)
        for k, (a, b) in enumerate(pairs):
            incidence[agents.index(a), k] = 1.0
            incidence[agents.index(b), k] = 1.0
        cache = dict(
            world=world,
            version=world._plan_version,
            a0=a0,
            n=len(agents),
            goal_idx=torch.tensor([ents.index(a.goal) for a in agents], device=dev),
            goal_radius=torch.tensor([a.goal.shape.radius for a in agents], device=dev).unsqueeze(-1),
            agent_radius=torch.tensor([a.shape.radius for a in agents], device=dev).unsqueeze(-1),
            pairs=pairs,
            goal_pairs=[(a, a.goal) for a in agents],
            incidence=incidence,
            sensors=[a.sensors[0] for a in agents],
            max_range=torch.tensor([a.sensors[0]._max_range for a in agents], device=dev).view(-1, 1, 1),
            # every agent's shaping term lives in one [A, B] block; agent.pos_shaping is row i of it
            pos_shaping=torch.stack([a.pos_shaping for a in agents]),
            final_reward=torch.tensor(float(self.final_reward), dtype=torch.float32, device=dev),
            zero=torch.tensor(0.0, dtype=torch.float32, device=dev),
        )
        for i, a in enumerate(agents):
            a.pos_shaping = cache["pos_shaping"][i]
        self._batch = cache
        return cache

    def _agent_goal_offsets(self, c):
        slab = self.world.slab
        apos = slab.pos[:, c["a0"] : c["a0"] + c["n"]].transpose(0, 1)  # [A, B, 2] view
        gpos = slab.pos.index_select(1, c["goal_idx"]).transpose(0, 1)
        return apos, apos - gpos

    def _reward_batched(self, agent: Agent):
        agents = self.world.agents
        c = self._batch_setup()
        if agent is agents[0]:
            # distances to the goals, shaping differences and the carried shaping block: one launch
            dist, pos_rew_all = self.world.distance_shaping(c["goal_pairs"], self.pos_shaping_factor, c["pos_shaping"])
            on_goal = dist < c["goal_radius"]
            shared = torch.zeros_like(self.pos_rew)
            for i, a in enumerate(agents):
                a.distance_to_goal, a.on_goal, a.pos_rew = dist[i], on_goal[i], pos_rew_all[i]
                shared = shared + pos_rew_all[i]  # sequential, like the reference's running sum
            self.pos_rew = shared
            self.all_goal_reached = on_goal.all(dim=0)
            self.final_rew = torch.where(self.all_goal_reached, c["final_reward"], c["zero"])
            c["dist_for_done"] = dist  # consumed by the next done()
            if c["pairs"]:
                touching = (self.world.get_distances(c["pairs"]) <= self.min_collision_distance) & self.world.collide_gates(
                    c["pairs"]
                ).unsqueeze(-1)
                collision_rew = (c["incidence"] @ touching.to(torch.float32)) * float(self.agent_collision_penalty)
            else:
                collision_rew = torch.zeros(len(agents), self.world.batch_dim, device=self.world.device)
            # (pos + final) + collision for every agent at once, same order as the per-agent sum
            if self.shared_rew:
                total = (self.pos_rew + self.final_rew).unsqueeze(0) + collision_rew
            else:
                total = (pos_rew_all + self.final_rew.unsqueeze(0)) + collision_rew
            for i, a in enumerate(agents):
                a.agent_collision_rew, a._total_rew = collision_rew[i], total[i]
        return agent._total_rew

    def _observation_batched(self, agent: Agent):
        agents = self.world.agents
        c = self._batch_setup()
        if agent is agents[0] or getattr(self, "_obs_all", None) is None:
            plan = c.get("obs_plan")
            if plan is None:
                plan = c["obs_plan"] = O.ObservationPlan(
                    [
                        [O.pos(a), O.vel(a), O.rel_pos(a, a.goal), O.lidar(a.sensors[0], range_minus_distance=True)]
                        for a in agents
                    ]
                )
            self._obs_all = self.world.observe(plan)  # [A, B, 18]: one gather + one LIDAR launch
        row = self._obs_all[agents.index(agent)]
        if agent is agents[-1]:
            self._obs_all = None  # one sweep over the agents per block: a later call measures anew
        return row

    def _done_batched(self):
        c = self._batch_setup()
        dist = c.pop("dist_for_done", None)
        if dist is None:  # no reward() since the last done(): measure the current state
            _, offset = self._agent_goal_offsets(c)
            dist = torch.linalg.vector_norm(offset, dim=-1)
        return (dist < c["agent_radius"]).all(dim=0)

    def _agent_progress(self, agent: Agent):
        agent.distance_to_goal = torch.linalg.vector_norm(agent.state.pos - agent.goal.state.pos, dim=-1)
        agent.on_goal = agent.distance_to_goal < agent.goal.shape.radius
        shaping = agent.distance_to_goal * self.pos_shaping_factor
        agent.pos_rew = agent.pos_shaping - shaping
        self.keep(agent, "pos_shaping", shaping)
        return agent.pos_rew

    def reward(self, agent: Agent):
        if self._batched_ok():
            return self._reward_batched(agent)
        agents = self.world.agents
        if agent is agents[0]:
            pos_rew = torch.zeros_like(self.pos_rew)
            for a in agents:
                pos_rew = pos_rew + self._agent_progress(a)
                a.agent_collision_rew = torch.zeros_like(a.agent_collision_rew)
            self.pos_rew = pos_rew
            self.all_goal_reached = torch.stack([a.on_goal for a in agents], dim=-1).all(dim=-1)
            self.final_rew = torch.where(self.all_goal_reached, float(self.final_reward), 0.0).to(
                torch.float32
            )
            for i, a in enumerate(agents):
                for j in range(i):
                    b = agents[j]
                    if not self.world.static_collides(a, b):
                        continue
                    # the reference gates this on ``world.collides`` (true iff the pair overlaps
                    # in SOME env of the batch); same gate as a device-side flag, without a sync
                    touching = (
                        self.world.get_distance(a, b) <= self.min_collision_distance
                    ) & self.world.collides_tensor(a, b)
                    penalty = torch.where(touching, float(self.agent_collision_penalty), 0.0)
                    a.agent_collision_rew = a.agent_collision_rew + penalty
                    b.agent_collision_rew = b.agent_collision_rew + penalty
        pos_reward = self.pos_rew if self.shared_rew else agent.pos_rew
        return pos_reward + self.final_rew + agent.agent_collision_rew

    def observation(self, agent: Agent):
        if self._batched_ok():
            return self._observation_batched(agent)
        if self.observe_all_goals:
            goal_rel = [agent.state.pos - a.goal.state.pos for a in self.world.agents]
        else:
            goal_rel = [agent.state.pos - agent.goal.state.pos]
        parts = [agent.state.pos, agent.state.vel] + goal_rel
        if self.collisions:
            lidar = agent.sensors[0]
            parts.append(lidar._max_range - lidar.measure())
        return torch.cat(parts, dim=-1)

    def done(self):
        if self._batched_ok():
            return self._done_batched()
        reached = [
            torch.linalg.vector_norm(a.state.pos - a.goal.state.pos, dim=-1) < a.shape.radius
            for a in self.world.agents
        ]
        return torch.stack(reached, dim=-1).all(-1)

    def info(self, agent: Agent) -> Dict[str, Tensor]:
        return {
            "pos_rew": self.pos_rew if self.shared_rew else agent.pos_rew,
            "final_rew": self.final_rew,
            "agent_collisions": agent.agent_collision_rew,
        }
