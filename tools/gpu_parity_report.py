"""Per-scenario errors of the CUDA step vs the reference's golden fixtures (teacher-forced).

    [VMAS_B200_ARITH=fast] python tools/gpu_parity_report.py [scenario ...]

Per world: max |err| per field and the largest fraction of the parity tolerance
(1e-5 + 1e-4 |want|, the north star's 1e-4 relative) any element used.
"""
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
print(f"# arithmetic={_native.ARITH}  mapping={os.environ.get('VMAS_B200_SPEC_MAPPING', _native.DEFAULT_SPEC_MAPPING)}")
for name in names:
    fix, desc, tables = load(name)
    worst = {k: 0.0 for k in STATE_KEYS}
    used = 0.0
    for t, state_in, fixed_rot, want in teacher_forced_steps(fix):
        dt = _device_tables(tables, fixed_rot, dev, ent_gravity=state_in.get("ent_gravity")) if (t == 0 or fixed_rot or "ent_gravity" in state_in) else dt
        slab = _Slab(state_in, dev)
        _native.world_step(lib, dt, slab)
        for k in STATE_KEYS:
            err = (slab.t[k].cpu() - want[k]).abs()
            worst[k] = max(worst[k], float(err.max()))
            used = max(used, float((err / (1e-5 + 1e-4 * want[k].abs())).max()))
    fields = "  ".join(f"{k} {worst[k]:.2e}" for k in ("pos", "vel", "rot", "ang_vel"))
    print(f"{name:22s} E={desc.n_entities:3d} items={len(desc.items):4d} mapping={dt.mapping:12s} tolerance used {used:7.3f}   {fields}")
