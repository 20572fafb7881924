"""Device-side episode reset on the GPU: the kernels behind
``vmas_b200_reset_state`` / ``vmas_b200_spawn_entities`` against the numpy oracle (bit for bit — same
Philox counters, same fp32 arithmetic), and ``Environment.reset_at`` with an env index or a bool mask.
"""
import numpy as np
import pytest
import torch

from oracle import reset as R

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _world(n_envs=300, name="flocking", **kwargs):
    import vectorizedmultiagentsimulator_b200 as b200

    kwargs = kwargs or dict(n_agents=5)
    env = b200.make_env(name, num_envs=n_envs, device=DEV, seed=0, **kwargs)
    assert env.world.uses_device_reset
    return env, env.world, env.world._get_backend()


def _np_pos(world):
    return world.slab.pos.detach().cpu().numpy().copy()


SELECTIONS = ["all", "index", "mask"]


@pytest.mark.parametrize("selection", SELECTIONS)
@pytest.mark.parametrize("occupied_kind", ["none", "per_env", "shared"])
def test_spawn_kernel_matches_numpy_oracle(selection, occupied_kind):
    env, world, backend = _world()
    B, ents = world.batch_dim, world.entities
    gen = torch.Generator().manual_seed(3)
    world.slab.pos.copy_((torch.rand(B, len(ents), 2, generator=gen) * 2 - 1).to(DEV))
    reset_count = torch.randint(0, 5, (B,), generator=gen, dtype=torch.int32).to(DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    occupied = None
    if occupied_kind == "per_env":
        occupied = (torch.rand(B, 3, 2, generator=gen) * 2 - 1).to(DEV)
    elif occupied_kind == "shared":
        occupied = torch.tensor([[[0.0, 0.0], [0.5, 0.5]]], device=DEV)
    env_index, mask_t, mask_np = None, None, None
    if selection == "index":
        env_index = 123
    elif selection == "mask":
        mask_t = (torch.rand(B, generator=gen) < 0.3).to(DEV)
        mask_np = mask_t.cpu().numpy()
    spawn = [ents[0], ents[3], None, ents[7], ents[1]]
    occ_ents = [ents[9], ents[10]]
    want_pos = _np_pos(world)
    want_out, want_exhausted = R.spawn_entities(
        want_pos,
        [0, 3, -1, 7, 1],
        min_dist=0.25,
        x_bounds=(-1.0, 1.0),
        y_bounds=(-0.8, 0.6),
        seed=0x1234_5678_9ABC,
        stream_id=5,
        reset_count=reset_count.cpu().numpy(),
        occupied_entities=[9, 10],
        occupied=None if occupied is None else occupied.cpu().numpy(),
        env_index=env_index,
        env_mask=mask_np,
    )
    out = backend.spawn(
        spawn,
        env_index if env_index is not None else mask_t,
        0.25,
        (-1.0, 1.0),
        (-0.8, 0.6),
        seed=0x1234_5678_9ABC,
        stream_id=5,
        reset_count=reset_count,
        status=status,
        occupied=occupied,
        occupied_entities=occ_ents,
        want_positions=True,
    )
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want_out)
    assert np.array_equal(_np_pos(world), want_pos)
    assert int(status.item()) == want_exhausted == 0


def test_spawn_reports_exhaustion_like_the_oracle():
    env, world, backend = _world(n_envs=64)
    ents = world.entities
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    want_pos = _np_pos(world)
    _, want_exhausted = R.spawn_entities(
        want_pos, [0, 1, 2, 3], min_dist=1.5, x_bounds=(0, 1), y_bounds=(0, 1), seed=7, max_tries=33
    )
    backend.spawn(ents[:4], None, 1.5, (0, 1), (0, 1), seed=7, stream_id=0, reset_count=None, status=status, max_tries=33)
    assert int(status.item()) == want_exhausted == 64
    assert np.array_equal(_np_pos(world), want_pos)  # the last proposal is kept, as in the oracle


@pytest.mark.parametrize("selection", SELECTIONS)
def test_reset_state_kernel(selection):
    env, world, backend = _world(n_envs=200, name="balance", n_agents=4)
    slab = world.slab
    gen = torch.Generator().manual_seed(1)
    for t in slab.tensors():
        t.copy_(torch.randn(t.shape, generator=gen).to(DEV))
    before = {k: v.cpu().numpy().copy() for k, v in slab.state_dict().items()}
    count = np.arange(200, dtype=np.int32)
    count_t = torch.from_numpy(count.copy()).to(DEV)
    env_index, mask = None, None
    if selection == "index":
        env_index = 199
    elif selection == "mask":
        mask = torch.rand(200, generator=gen) < 0.5
    R.reset_state(before, count, env_index=env_index, env_mask=None if mask is None else mask.numpy())
    backend.reset_state(env_index if env_index is not None else (None if mask is None else mask.to(DEV)), count_t)
    torch.cuda.synchronize()
    for k, v in slab.state_dict().items():
        assert np.array_equal(v.cpu().numpy(), before[k]), k
    assert np.array_equal(count_t.cpu().numpy(), count)


CASES = [
    ("balance", dict(n_agents=4)),
    ("transport", dict(n_agents=4)),
    ("navigation", dict(n_agents=4)),
    ("flocking", dict(n_agents=5)),
]


def _make(name, kwargs, n_envs, **extra):
    import vectorizedmultiagentsimulator_b200 as b200

    return b200.make_env(name, num_envs=n_envs, device=DEV, seed=4, **kwargs, **extra)


def _actions(env, gen):
    return [(torch.rand(env.num_envs, 2, generator=gen) * 2 - 1).to(DEV) for _ in env.agents]


def _slab(env):
    return {k: v.clone() for k, v in env.world.slab.state_dict().items()}


@pytest.mark.parametrize("name,kwargs", CASES)
def test_env_masked_reset(name, kwargs):
    n_envs = 512
    env = _make(name, kwargs, n_envs)
    gen = torch.Generator().manual_seed(0)
    for _ in range(3):
        env.step(_actions(env, gen))
    before, steps_before = _slab(env), env.steps.clone()
    mask = (torch.rand(n_envs, generator=gen) < 0.25).to(DEV)
    launches = env.world._get_backend().launches
    obs = env.reset_at(mask)
    assert env.world._get_backend().launches > launches  # the reset ran this library's kernels
    after = _slab(env)
    assert all(torch.isfinite(o).all() for o in obs)
    for k in before:
        assert torch.equal(after[k][~mask], before[k][~mask]), f"{name}: {k} of an unflagged env changed"
    assert float(after["vel"][mask].abs().max()) == 0.0 and float(after["ang_vel"][mask].abs().max()) == 0.0
    assert not torch.equal(after["pos"][mask], before["pos"][mask])
    assert float(env.steps[mask].abs().max()) == 0.0 and torch.equal(env.steps[~mask], steps_before[~mask])
    assert torch.equal(env.world.reset_count, 1 + mask.to(torch.int32))
    assert env.world.spawn_failures() == 0
    env.step(_actions(env, gen))  # and the env keeps stepping


@pytest.mark.parametrize("name,kwargs", CASES)
def test_masked_reset_equals_one_reset_at_per_env(name, kwargs):
    n_envs = 96
    a, b = _make(name, kwargs, n_envs), _make(name, kwargs, n_envs)
    assert all(torch.equal(x, y) for x, y in zip(_slab(a).values(), _slab(b).values()))  # same seed, same layout
    gen = torch.Generator().manual_seed(2)
    for _ in range(2):
        act = _actions(a, gen)
        a.step([t.clone() for t in act])
        b.step([t.clone() for t in act])
    flagged = [0, 17, 18, 95]
    mask = torch.zeros(n_envs, dtype=torch.bool, device=DEV)
    mask[flagged] = True
    obs_a = a.reset_at(mask)
    for i in flagged:
        obs_b = b.reset_at(i)
    for (k, x), y in zip(_slab(a).items(), _slab(b).values()):
        assert torch.equal(x, y), f"{name}: {k}"
    assert all(torch.equal(x, y) for x, y in zip(obs_a, obs_b))


def test_spawned_layout_respects_the_scenario_constraints():
    env = _make("navigation", dict(n_agents=6), 2048)
    sc, world = env.scenario, env.world
    env.reset_at(torch.ones(2048, dtype=torch.bool, device=DEV))  # second episode of every env
    pts = torch.stack([a.state.pos for a in world.agents] + [a.goal.state.pos for a in world.agents], dim=1)
    assert float(pts[..., 0].abs().max()) <= sc.world_spawning_x and float(pts[..., 1].abs().max()) <= sc.world_spawning_y
    d = torch.cdist(pts, pts) + torch.eye(pts.shape[1], device=DEV) * 10
    assert float(d.min()) >= sc.min_distance_between_entities - 1e-6
    assert world.spawn_failures() == 0
    # different envs, different layouts; both episodes of an env differ as well
    assert not torch.equal(pts[0], pts[1])


def test_seed_reproduces_the_layout():
    a = _make("flocking", dict(n_agents=5), 256)
    b = _make("flocking", dict(n_agents=5), 256)
    assert torch.equal(a.world.slab.pos, b.world.slab.pos)
    a.reset(seed=11)
    b.reset(seed=12)
    assert not torch.equal(a.world.slab.pos, b.world.slab.pos)
    b.world.reset_count.copy_(a.world.reset_count - 1)
    b.reset(seed=11)
    assert torch.equal(a.world.slab.pos, b.world.slab.pos)


def test_masked_reset_between_graph_replays():
    n_envs = 256
    eager = _make("navigation", dict(n_agents=4), n_envs)
    graph = _make("navigation", dict(n_agents=4), n_envs, cuda_graph=True)
    gen = torch.Generator().manual_seed(5)
    mask = (torch.rand(n_envs, generator=gen) < 0.5).to(DEV)
    for t in range(8):
        act = _actions(eager, gen)
        want = eager.step([x.clone() for x in act])
        got = graph.step([x.clone() for x in act])
        for g, w in zip(got[0] + got[1] + [got[2]], want[0] + want[1] + [want[2]]):
            assert torch.equal(g, w), f"step {t}"
        if t == 4:
            assert graph._graph is not None
            for o_g, o_e in zip(graph.reset_at(mask), eager.reset_at(mask)):
                assert torch.equal(o_g, o_e)
    assert graph.graph_replays > 0


def test_shards_reset_like_the_unsharded_job():
    """Two shard envs (``shard.make_shard_env``, here both on cuda:0) against the unsharded job."""
    from vectorizedmultiagentsimulator_b200 import shard

    total, kwargs = 1000, dict(n_agents=5)
    full = shard.make_shard_env("flocking", total, 0, 1, DEV, seed=6, **kwargs)
    mask = (torch.rand(total, generator=torch.Generator().manual_seed(0)) < 0.4).to(DEV)
    full.reset_at(mask)
    for rank in range(2):
        lo, hi = shard.shard_bounds(total, rank, 2)
        part = shard.make_shard_env("flocking", total, rank, 2, DEV, seed=6, **kwargs)
        part.reset_at(mask[lo:hi])
        for (k, got), want in zip(_slab(part).items(), _slab(full).values()):
            assert torch.equal(got, want[lo:hi]), f"rank {rank}: {k}"


def _staggered(name, kwargs, n_envs, **extra):
    env = _make(name, kwargs, n_envs, max_steps=4, **extra)
    env.steps.copy_((torch.arange(n_envs, dtype=torch.float32) % 4).to(DEV))  # episodes end at different steps
    return env


@pytest.mark.parametrize("name,kwargs", CASES)
def test_auto_reset_equals_step_then_reset_at_dones(name, kwargs):
    n_envs = 200
    auto = _staggered(name, kwargs, n_envs, auto_reset=True)
    manual = _staggered(name, kwargs, n_envs)
    gen = torch.Generator().manual_seed(1)
    for t in range(9):
        act = _actions(auto, gen)
        got = auto.step([a.clone() for a in act])
        want = manual.step([a.clone() for a in act])
        want_obs = manual.reset_at(want[2])
        for g, w in zip(got[0] + got[1] + [got[2]], want_obs + want[1] + [want[2]]):
            assert torch.equal(g, w), f"{name} step {t}"
        assert torch.equal(auto.steps, manual.steps)
    assert int(auto.world.reset_count.min()) >= 3 and auto.world.spawn_failures() == 0


@pytest.mark.parametrize("name,kwargs", CASES)
def test_auto_reset_inside_the_step_graph(name, kwargs):
    """Graph mode captures the reset kernels with the step; replays must equal the eager env."""
    n_envs = 128
    eager = _staggered(name, kwargs, n_envs, auto_reset=True)
    graph = _staggered(name, kwargs, n_envs, auto_reset=True, cuda_graph=True)
    gen = torch.Generator().manual_seed(3)
    for t in range(10):
        act = _actions(eager, gen)
        want = eager.step([a.clone() for a in act])
        got = graph.step([a.clone() for a in act])
        assert torch.equal(got[2], want[2]), f"{name} step {t}: dones"
        assert torch.equal(graph.steps, eager.steps)
        for g, w in zip(got[0] + got[1], want[0] + want[1]):
            assert torch.equal(g, w), f"{name} step {t}"
    assert graph._graph is not None and graph.graph_replays >= 6
    assert torch.equal(graph.world.reset_count, eager.world.reset_count)
    assert int(graph.world.reset_count.min()) >= 3
