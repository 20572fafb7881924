"""Counts aten ops per phase of Environment.step on the CPU (oracle backend) — each op is one
kernel launch on the GPU, so this is the launch budget of the scenario callbacks.  A developer aid
that lives under tests/ because it drives the CPU oracle (test infrastructure).

    python tests/count_ops.py balance n_agents=4
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
from torch.utils._python_dispatch import TorchDispatchMode

import vectorizedmultiagentsimulator_b200 as vmas
from oracle.backend import use_oracle


class Counter(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.phase = "other"
        self.counts = collections.Counter()
        self.ops = collections.defaultdict(collections.Counter)

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.__name__
        if not any(s in name for s in ("view", "unsqueeze", "squeeze", "expand", "select", "slice", "transpose",
                                       "alias", "detach", "permute", "as_strided", "unbind", "_unsafe_view", "t.default",
                                       "reshape", "sym_", "is_", "split", "unfold")):
            self.counts[self.phase] += 1
            self.ops[self.phase][name] += 1
        return func(*args, **(kwargs or {}))


def main():
    name = sys.argv[1]
    kwargs = {}
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        kwargs[k] = int(v)
    with use_oracle():
        env = vmas.make_env(name, num_envs=32, device="cpu", seed=0, **kwargs)
        env.reset()
        acts = env.get_random_actions()
        env.step(acts)
        c = Counter()
        sc = env.scenario
        for meth in ("reward", "observation", "done", "info", "process_action", "pre_step", "post_step"):
            orig = getattr(sc, meth)

            def wrapped(*a, _o=orig, _m=meth, **k):
                prev, c.phase = c.phase, _m
                try:
                    return _o(*a, **k)
                finally:
                    c.phase = prev

            setattr(sc, meth, wrapped)
        ws = env.world.step

        def wstep():
            prev, c.phase = c.phase, "world.step(oracle)"
            try:
                return ws()
            finally:
                c.phase = prev

        env.world.step = wstep
        be = env.world._backend
        c.native = collections.Counter()
        for attr in dir(be):
            fn = getattr(be, attr)
            if attr.startswith("_") or not callable(fn) or attr in ("step", "refresh"):
                continue

            def nat(*a, _f=fn, _n=attr, **k):
                prev, c.phase = c.phase, "native"
                c.native[(prev, _n)] += 1
                try:
                    return _f(*a, **k)
                finally:
                    c.phase = prev

            setattr(be, attr, nat)
        with c:
            env.step(acts)
    print("native calls (one launch each on the GPU):", dict(c.native))
    for ph, n in c.counts.most_common():
        print(f"{ph:22s} {n}")
        if ph not in ("world.step(oracle)", "native"):
            print("    ", dict(c.ops[ph].most_common(12)))


main()
