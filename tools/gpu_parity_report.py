"""Prints per-scenario / per-field / per-entity errors of the CUDA step vs the golden fixtures."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import STATE_KEYS, golden_names, load, teacher_forced_steps
from vectorizedmultiagentsimulator_b200 import _native
from test_cabi_gpu import _Slab, _device_tables

lib = _native.load()
dev = torch.device("cuda:0")
names = sys.argv[1:] or golden_names()
for name in names:
    fix, desc, tables = load(name)
    worst = {k: (0.0, None) for k in STATE_KEYS}
    for t, state_in, fixed_rot, want in teacher_forced_steps(fix):
        dt = _device_tables(tables, fixed_rot, dev) if (t == 0 or fixed_rot) else dt
        slab = _Slab(state_in, dev)
        _native.world_step(lib, dt, slab)
        for k in STATE_KEYS:
            err = (slab.t[k].cpu() - want[k]).abs()
            m = float(err.max())
            if m > worst[k][0]:
                idx = (err == err.max()).nonzero()[0].tolist()
                worst[k] = (m, (t, idx))
    print(name, "E=%d items=%d" % (desc.n_entities, len(desc.items)))
    for k in STATE_KEYS:
        m, where = worst[k]
        ent = ""
        if where and k in ("pos", "vel", "rot", "ang_vel"):
            ent = desc.entities[where[1][1]]["name"]
        print(f"   {k:8s} {m:.3e} at {where} {ent}")
