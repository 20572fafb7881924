"""``make_env(..., wrapper=...)``: the gym / gymnasium adapters over the Environment (CPU oracle backend).

Contracts of the reference's adapters (vmas/simulator/environment/gym/*.py): one-env wrappers strip the
batch dimension and return python scalars for rewards / done, the vectorised one keeps ``[num_envs]``;
gymnasium flavours return the 5-tuple and need ``terminated_truncated=True``; values equal what the
bare Environment returns for the same actions.
"""
import numpy as np
import pytest
import torch

import vectorizedmultiagentsimulator_b200 as b200
from oracle.backend import use_oracle


def _pair(wrapper, num_envs, **kw):
    with use_oracle():
        bare = b200.make_env("balance", num_envs=num_envs, device="cpu", seed=0, n_agents=3, **kw)
        wrapped = b200.make_env("balance", num_envs=num_envs, device="cpu", seed=0, n_agents=3, wrapper=wrapper, **kw)
    return bare, wrapped


def test_gym_wrapper_single_env():
    bare, env = _pair("gym", 1)
    with use_oracle():
        obs = env.reset(seed=4)
        assert isinstance(obs, list) and obs[0].shape == (bare.observation_space[0].shape[0],) and isinstance(obs[0], np.ndarray)
        bare.seed(4)
        want0 = bare.reset_at(index=0)
        assert np.array_equal(obs[0], want0[0][0].numpy())
        gen = torch.Generator().manual_seed(0)
        for _ in range(3):
            acts = [(torch.rand(2, generator=gen) * 2 - 1).numpy() for _ in bare.agents]
            o, r, d, info = env.step(acts)
            wo, wr, wd, _ = bare.step([torch.as_tensor(a).reshape(1, 2) for a in acts])
            assert np.array_equal(o[1], wo[1][0].numpy()) and r[0] == float(wr[0][0]) and d == bool(wd[0])
            assert isinstance(r[0], float) and isinstance(d, bool) and set(info) == {a.name for a in bare.agents}


def test_gymnasium_wrappers():
    with pytest.raises(AssertionError):
        _pair("gymnasium", 1)  # needs terminated_truncated=True
    with pytest.raises(AssertionError):
        _pair("gym", 2)  # not vectorised
    bare, env = _pair("gymnasium", 1, terminated_truncated=True)
    with use_oracle():
        obs, info = env.reset(seed=1)
        assert len(obs) == 3 and isinstance(info, dict)
        o, r, term, trunc, info = env.step([np.zeros(2, np.float32)] * 3)
        assert isinstance(term, bool) and isinstance(trunc, bool) and isinstance(r[0], float)
    bare, vec = _pair("gymnasium_vec", 5, terminated_truncated=True, wrapper_kwargs=dict(return_numpy=False))
    with use_oracle():
        obs, _ = vec.reset(seed=2)
        assert obs[0].shape[0] == 5 and isinstance(obs[0], torch.Tensor)
        acts = [torch.zeros(5, 2) for _ in range(3)]
        o, r, term, trunc, _ = vec.step(acts)
        bare.reset(seed=2)
        wo, wr, wterm, wtrunc, _ = bare.step(acts)
        assert torch.equal(o[2], wo[2]) and torch.equal(r[1], wr[1]) and torch.equal(term, wterm) and term.shape == (5,)


def test_rllib_wrapper_fails_up_front_with_instructions():
    with use_oracle(), pytest.raises(ImportError, match="VectorEnvWrapper"):
        b200.make_env("balance", num_envs=2, device="cpu", seed=0, wrapper="rllib")


def test_stack_views_is_a_view_when_the_slices_are_adjacent_and_a_copy_otherwise():
    import vectorizedmultiagentsimulator_b200 as b200

    block = torch.arange(4 * 6 * 3, dtype=torch.float32).reshape(4, 6, 3)
    rows = list(block.unbind(0))
    view = b200.stack_views(rows)
    assert view.shape == (4, 6, 3) and view.data_ptr() == block.data_ptr() and torch.equal(view, block)
    tail = b200.stack_views(rows[1:3])  # a run that starts inside the block
    assert tail.data_ptr() == rows[1].data_ptr() and torch.equal(tail, block[1:3])
    # not adjacent / not the same storage / strided: a plain stack
    for parts in ([rows[0], rows[2]], [rows[0], rows[1].clone()], [block[:, 0], block[:, 1]]):
        out = b200.stack_views(parts)
        assert torch.equal(out, torch.stack(parts)) and out.data_ptr() not in {p.data_ptr() for p in parts}
