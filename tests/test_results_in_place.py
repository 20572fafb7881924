"""Host logic of the captured step's "results in place" plan (``Environment._results_in_place``): which hand-out
copies turn into an observation-block redirect, which into one more STORE of the step program (a mirror), and
which stay copies.  Pure bookkeeping on tensors' addresses — checked here on CPU tensors (the GPU tests check the
launches it drives: tests/test_env_gpu.py)."""
import torch

import vectorizedmultiagentsimulator_b200 as b200
from oracle.backend import use_oracle
from vectorizedmultiagentsimulator_b200 import _native as N
from vectorizedmultiagentsimulator_b200.simulator import program as SP


def _setup(n_envs=8):
    with use_oracle():
        env = b200.make_env("balance", num_envs=n_envs, device="cpu", seed=0, n_agents=4)
        env.reset()
        env.step(env.get_random_actions())
    sc = env.scenario
    prog, plan = sc._step_program().finalize(), sc._observation_plan()
    cols, _ = plan.compile(env.world)
    index = {id(e): i for i, e in enumerate(env.world.entities)}
    instrs = prog.instructions(lambda e: index[id(e)])
    c = N.StepProgramC()
    c.n_instr = len(instrs)
    for k, (op, dst, a, b, arg, imm) in enumerate(instrs):
        ins = c.instr[k]
        ins.op, ins.dst, ins.a, ins.b, ins.arg, ins.imm = op, dst, a, b, arg, imm
    out = torch.zeros(plan.n_rows, n_envs, plan.width)
    return env, prog, plan, cols, instrs, c, out


def test_observation_block_and_program_outputs_are_written_in_place():
    env, prog, plan, cols, instrs, c, out = _setup()
    B = env.num_envs
    other = torch.zeros(B)  # a result the program does not produce: stays a copy
    items = [(out.reshape(-1), 0, 0)]  # one pack covering the whole observation block
    offset = 0
    for t in (prog.out_rew.tensor, prog.out_rew.tensor, prog.out_pos_rew.tensor, other):
        items.append((t, 1, offset))
        offset += 4 * B
    items.append((prog.out_done.tensor, 2, 0))
    c2, instrs2, rest, obs_to, mirrors = env._results_in_place(c, prog, instrs, items, cols, out)
    assert obs_to == (0, 0)
    assert [(r[1], r[2]) for r in rest] == [(1, 12 * B)] and rest[0][0] is other
    n_slots = len(prog.buffers)
    assert [m[0] for m in mirrors] == [n_slots, n_slots + 1, n_slots + 2, n_slots + 3]
    assert [(m[1], m[2]) for m in mirrors] == [(1, 0), (1, 4 * B), (1, 8 * B), (2, 0)]
    # one STORE per mirrored leaf, of the register the original store of that output writes
    extra = instrs2[len(instrs):]
    assert c2.n_instr == len(instrs) + 4 and len(extra) == 4
    store_reg = {b: (op, a) for op, _, a, b, _, _ in instrs if op in (SP.OP_STORE_F32, SP.OP_STORE_BOOL)}
    want = [store_reg[o._slot] for o in (prog.out_rew, prog.out_rew, prog.out_pos_rew, prog.out_done)]
    assert [(op, a) for op, _, a, _, _, _ in extra] == want
    assert [b for _, _, _, b, _, _ in extra] == [m[0] for m in mirrors]
    # the program struct handed to the kernel carries the extra instructions behind the original ones
    for k, (op, dst, a, b, arg, imm) in enumerate(instrs2):
        ins = c2.instr[k]
        assert (ins.op, ins.a, ins.b) == (op, a, b)
    assert c.n_instr == len(instrs)  # (the eager program is left alone)


def test_observation_rows_handed_out_piece_by_piece_count_as_one_block():
    env, prog, plan, cols, instrs, c, out = _setup()
    row_bytes = out[0].numel() * 4
    pieces = [(out[r].reshape(-1), 1, 64 + r * row_bytes) for r in range(plan.n_rows)]
    _, _, rest, obs_to, _ = env._results_in_place(c, prog, instrs, pieces, cols, out)
    assert obs_to == (1, 64) and rest == []
    # ... but not when a piece goes elsewhere, or the destination is not 16-byte aligned
    moved = pieces[:-1] + [(pieces[-1][0], 2, 0)]
    _, _, rest, obs_to, _ = env._results_in_place(c, prog, instrs, moved, cols, out)
    assert obs_to is None and len(rest) == plan.n_rows
    shifted = [(src, block, offset + 4) for src, block, offset in pieces]
    _, _, rest, obs_to, _ = env._results_in_place(c, prog, instrs, shifted, cols, out)
    assert obs_to is None and len(rest) == plan.n_rows


def test_mirrors_stop_at_the_programs_buffer_and_instruction_limits():
    env, prog, plan, cols, instrs, c, out = _setup()
    B = env.num_envs
    many = [(prog.out_rew.tensor, 1, 4 * B * k) for k in range(40)]
    _, instrs2, rest, _, mirrors = env._results_in_place(c, prog, instrs, many, cols, out)
    assert len(prog.buffers) + len(mirrors) == N.PROG_MAX_BUFFERS
    assert len(rest) == 40 - len(mirrors) and len(instrs2) <= N.PROG_MAX_INSTR
