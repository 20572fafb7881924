"""``StepProgram`` (the scenario's reward / done glue as one launch, fused with the observation gather):
the CUDA interpreter against the torch interpretation of the CPU oracle backend, every opcode."""
import pytest
import torch

import vectorizedmultiagentsimulator_b200 as b200
from envutil import sync_env
from oracle.backend import use_oracle
from vectorizedmultiagentsimulator_b200.simulator import observe as O
from vectorizedmultiagentsimulator_b200.simulator.program import StepProgram

pytestmark = pytest.mark.gpu


def _build(env, carried, flag, level):
    w, sc = env.world, env.scenario
    a0, a1 = w.agents[0], w.agents[1]
    p = StepProgram(w)
    o1, o2 = p.overlap(sc.line, sc.floor), p.overlap(sc.package, sc.package.goal)
    d1, d2, c1 = p.distance(a0, sc.line), p.distance(sc.package, sc.floor), p.center_distance(a0, a1)
    rew, dist = p.shaping(sc.package, sc.package.goal, 3.5, prev=carried)
    f, lvl = p.load(flag, is_bool=True), p.load(level)
    k = p.const(0.25)
    outs = dict(  # (16 buffers per program: 3 inputs + 13 outputs here)
        o1=p.store(o1), o2=p.store(o2), d1=p.store(d1), d2=p.store(d2), c1=p.store(c1), rew=p.store(rew), dist=p.store(dist),
        add=p.store(p.add(d1, k)), sub=p.store(p.sub(c1, lvl)), mul=p.store(p.mul(d2, lvl)), mn=p.store(p.minimum(d1, d2)),
        lor=p.store(p.logical_or(o1, f)), land=p.store(p.logical_and(p.logical_not(o2), f)),
    )
    # a second program for the remaining opcodes
    q = StepProgram(w)
    e1, e2 = q.distance(a0, sc.line), q.center_distance(a0, a1)
    lvl2 = q.load(level)
    outs2 = dict(
        mx=q.store(q.maximum(e1, e2)), neg=q.store(q.neg(e2)), lt=q.store(q.lt(e1, lvl2)), le=q.store(q.le(e2, e2)),
        where=q.store(q.where(q.lt(e1, e2), e1, q.const(-1.0))), where_b=q.store(q.where(q.le(lvl2, e2), q.lt(e1, e2), q.le(e1, e2))),
    )
    return (p.finalize(), outs), (q.finalize(), outs2)


def test_every_opcode_matches_the_torch_interpretation():
    n = 300
    with use_oracle():
        cpu = b200.make_env("balance", num_envs=n, device="cpu", seed=0, n_agents=3)
    gpu = b200.make_env("balance", num_envs=n, device="cuda", seed=0, n_agents=3)
    gen = torch.Generator().manual_seed(0)
    for _ in range(25):  # roll until things touch
        acts = [torch.rand(n, 2, generator=gen) * 2 - 1 for _ in cpu.agents]
        cpu.step(acts)
    sync_env(cpu, gpu)
    carried = torch.rand(n, generator=gen)
    flag = torch.rand(n, generator=gen) > 0.5
    level = torch.rand(n, generator=gen) * 0.4
    with use_oracle():
        progs_cpu = _build(cpu, carried.clone(), flag.clone(), level.clone())
    progs_gpu = _build(gpu, carried.cuda(), flag.cuda(), level.cuda())
    plan_cpu = O.ObservationPlan([[O.pos(a), O.vel(a)] for a in cpu.world.agents])
    plan_gpu = O.ObservationPlan([[O.pos(a), O.vel(a)] for a in gpu.world.agents])
    for k, ((pc, oc), (pg, og)) in enumerate(zip(progs_cpu, progs_gpu)):
        with use_oracle():
            obs_c = pc.run(observe=plan_cpu if k == 0 else None)
        obs_g = pg.run(observe=plan_gpu if k == 0 else None)
        if k == 0:
            assert torch.equal(obs_g.cpu(), obs_c)
        for name in oc:
            want, got = oc[name].tensor, og[name].tensor.cpu()
            assert want.dtype == got.dtype, name
            if want.dtype == torch.bool:
                assert torch.equal(got, want), f"{name}: {int((got != want).sum())} flags differ"
            else:
                err = (got - want).abs()
                assert bool((err <= 1e-6 + 1e-5 * want.abs()).all()), f"{name}: max |err| {float(err.max())}"
    assert bool(progs_gpu[0][1]["o1"].tensor.any()) or bool(progs_gpu[0][1]["lor"].tensor.any())
