"""How often does sharding the batch change a result?  (CPU oracle; the measurement DESIGN.md quotes.)

The reference activates a line / box collision pair for ALL envs as soon as ANY env of the batch has
the two circumscribed circles overlapping (ref core.py:2797-2801); an env whose own circles do not
overlap can still receive a force from that pair, because the contact threshold adds the minimum
distance on top of the shapes (a centre distance in ``(R_a + R_b, R_a + R_b + d_min]``).  A shard
evaluates the activation over its own envs only, so a sharded job equals the unsharded one except in
exactly those envs — and only when no env of the SHARD activates the pair while some env of another
shard does.  This test rolls ``balance`` (the benched, sharded workload: L-S, B-S and B-L pairs) out,
steps every state once unsharded and once as two shards, counts the differing envs and checks that
every one of them is explained by that mechanism.
"""
import torch

from golden_util import load, teacher_forced_steps
from oracle import world_step as WS

KEYS = ("pos", "vel", "rot", "ang_vel")


def test_two_shards_differ_from_unsharded_only_through_the_batch_wide_activation():
    fix, desc, tables = load("balance")  # 64 envs, 100 reference steps
    B = desc.batch_dim
    half = B // 2
    assert tables.n_masked > 0, "balance has line / box pairs"
    masked = [int(k) for k in tables.masked_items[: tables.n_masked]]
    diff_envs = explained = steps = 0
    for t, state_in, _, _ in teacher_forced_steps(fix):
        full = {k: v.clone() for k, v in state_in.items()}
        WS.world_step(tables, full)
        parts = []
        for lo, hi in ((0, half), (half, B)):
            part = {k: v[lo:hi].clone() for k, v in state_in.items()}
            WS.world_step(tables, part)
            parts.append(part)
        differs = torch.zeros(B, dtype=torch.bool)
        for k in KEYS:
            sharded = torch.cat([p[k] for p in parts])
            differs |= (sharded != full[k]).flatten(1).any(1)
        # the mechanism: a pair active in the whole batch but in no env of the shard
        act_full = WS.broad_phase_active_many(tables, masked, state_in["pos"])
        can_differ = torch.zeros(B, dtype=torch.bool)
        for lo, hi in ((0, half), (half, B)):
            act = WS.broad_phase_active_many(tables, masked, state_in["pos"][lo:hi])
            if any(f and not s for f, s in zip(act_full, act)):
                can_differ[lo:hi] = True
        assert not bool((differs & ~can_differ).any()), f"step {t}: a shard differs without a lost activation"
        diff_envs += int(differs.sum())
        explained += int((differs & can_differ).sum())
        steps += 1
    print(f"balance, {B} envs as 2 shards, {steps} steps: {diff_envs} of {steps * B} env-steps differ "
          f"({100.0 * diff_envs / (steps * B):.3f} %), all inside shards that lost a batch-wide activation")
    assert diff_envs == explained
    assert diff_envs <= 0.02 * steps * B  # rare: the shell between R_a + R_b and R_a + R_b + d_min is thin
