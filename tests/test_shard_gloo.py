"""The N>1 path on CPU: two gloo ranks each step their shard of the env batch (CPU oracle backend).

Checks the shard arithmetic, that a sharded run equals the matching slice of an unsharded run
(envs are independent; flocking is sphere-only so the batch-wide broad phase is result-neutral),
and the max-/sum-over-ranks reductions bench.py relies on.  A second job runs the ranks through
``shard.make_shard_env`` with the device-reset host path (the stand-in of test_reset_host_path.py):
there the shards also *reset* like the unsharded job, so nothing has to be copied between them.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from vectorizedmultiagentsimulator_b200 import shard  # noqa: E402

TOTAL, STEPS = 12, 4


def _rollout(lo, hi):
    """Steps envs [lo, hi) of the reference job: state of the full-batch env, sliced."""
    import vectorizedmultiagentsimulator_b200 as b200
    from oracle.backend import use_oracle

    torch.set_num_threads(1)
    with use_oracle():
        full = b200.make_env("flocking", num_envs=TOTAL, device="cpu", seed=0, n_agents=3)
        env = b200.make_env("flocking", num_envs=hi - lo, device="cpu", seed=1, n_agents=3)
        # the shard starts from its slice of the full job's initial state
        for k, v in full.world.slab.state_dict().items():
            getattr(env.world.slab, k).copy_(v[lo:hi])
        for src, dst in zip(full.world.agents, env.world.agents):
            if hasattr(src, "distance_shaping"):
                dst.distance_shaping.copy_(src.distance_shaping[lo:hi])
        gen = torch.Generator().manual_seed(5)
        out = None
        for _ in range(STEPS):
            acts = [(torch.rand(TOTAL, 2, generator=gen) * 2 - 1)[lo:hi] for _ in env.agents]
            out = env.step(acts)
        return torch.stack(out[0], 1), torch.stack(out[1], 1)


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.dist_env()[:2] == (rank, world)
        lo, hi = shard.shard_bounds(TOTAL, rank, world)
        obs, rew = _rollout(lo, hi)
        torch.save((lo, hi, obs, rew), os.path.join(tmpdir, f"shard{rank}.pt"))
        # control-plane reductions
        assert shard.max_over_ranks(float(rank + 1)) == float(world)
        assert shard.sum_over_ranks(float(hi - lo)) == float(TOTAL)
        agg = shard.aggregate_throughput(float(hi - lo) * STEPS, 1.0 + rank)
        assert abs(agg - TOTAL * STEPS / float(world)) < 1e-9
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for total in (1, 7, 32768, 262144):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_gloo_ranks_reproduce_the_unsharded_run(tmp_path):
    world, port = 2, 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    full_obs, full_rew = _rollout(0, TOTAL)
    for r in range(world):
        lo, hi, obs, rew = torch.load(os.path.join(str(tmp_path), f"shard{r}.pt"))
        assert torch.equal(obs, full_obs[lo:hi]), f"rank {r} observations differ from the unsharded run"
        assert torch.equal(rew, full_rew[lo:hi])


def _self_contained_rollout(rank, world):
    """make_shard_env -> steps -> masked reset -> steps, with the device-reset host path."""
    from test_reset_host_path import HostPathBackend
    from vectorizedmultiagentsimulator_b200.simulator.core import World

    torch.set_num_threads(1)
    World._backend_factory = staticmethod(lambda w: HostPathBackend(w))
    World.uses_device_reset = property(lambda self: True)
    lo, hi = shard.shard_bounds(TOTAL, rank, world)
    env = shard.make_shard_env("flocking", TOTAL, rank, world, "cpu", seed=3, n_agents=3)
    gen = torch.Generator().manual_seed(5)
    done = torch.rand(TOTAL, generator=gen) < 0.5
    out = None
    for t in range(STEPS):
        acts = [(torch.rand(TOTAL, 2, generator=gen) * 2 - 1)[lo:hi] for _ in env.agents]
        out = env.step(acts)
        if t == 1:
            env.reset_at(done[lo:hi])
    return lo, hi, torch.stack(out[0], 1), torch.stack(out[1], 1), env.world.slab.pos.clone()


def _self_contained_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.save(_self_contained_rollout(rank, world), os.path.join(tmpdir, f"self{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks_reset_and_step_like_the_unsharded_job(tmp_path):
    world, port = 2, 30500 + os.getpid() % 1000
    mp.spawn(_self_contained_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vectorizedmultiagentsimulator_b200.simulator.core import World

    saved = World._backend_factory, World.uses_device_reset
    try:
        _, _, full_obs, full_rew, full_pos = _self_contained_rollout(0, 1)
    finally:
        World._backend_factory, World.uses_device_reset = saved
    for r in range(world):
        lo, hi, obs, rew, pos = torch.load(os.path.join(str(tmp_path), f"self{r}.pt"))
        assert torch.equal(pos, full_pos[lo:hi]), f"rank {r}: positions differ from the unsharded job"
        assert torch.equal(obs, full_obs[lo:hi]) and torch.equal(rew, full_rew[lo:hi])
