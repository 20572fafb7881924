"""Parity of the sm_100a kernels, called through the C ABI, against reference roll-outs.

Teacher-forced: every golden step's exact ``World.step`` input (state slab + processed action
forces) is loaded on the GPU, ``vmas_b200_world_step`` runs once, and the result is compared
with what the unmodified reference produced (tests/golden/, made by tests/make_golden.py) and
with the CPU oracle run live on the same input.

Tolerance (north star: 1e-4 relative, fp32): |got - want| <= 1e-5 + 1e-4 * |want|.
Worlds with zero-length joints add the reference's own sensitivity: the joint force is
c * delta/|delta| * pen with |delta| down to 1e-6 (anchors coincide right after a reset), which
amplifies a 1-ulp difference in an anchor position (CUDA vs SLEEF sin/cos) by up to 1e5.  For
those worlds the bound is widened by 4x what the CPU oracle itself moves when its inputs are
perturbed by 1 ulp (measured live, per step and field; see DESIGN.md "parity envelope").
"""
import pytest
import torch

from golden_util import same_result, STATE_KEYS, golden_names, load, teacher_forced_steps
from oracle import queries as Q
from oracle import world_step as WS
from vectorizedmultiagentsimulator_b200 import _native
from vectorizedmultiagentsimulator_b200.simulator.slab import StateSlab

pytestmark = pytest.mark.gpu

RTOL = 1e-4
JOINT_WORLDS = {"waterfall", "joint_passage", "wheel"}


class _Slab:
    """Bare state slab (no World object): exactly what the C ABI consumes."""

    def __init__(self, state, device):
        self.t = {k: state[k].to(device).contiguous() for k in STATE_KEYS}

    def tensors(self):
        return tuple(self.t[k] for k in STATE_KEYS)


ATOL = 1e-5


def _ulp_sensitivity(tables, state_in, fixed_rot, trials=3):
    """max |oracle(x) - oracle(x perturbed by +-1 ulp)| per field."""
    base = {k: v.clone() for k, v in state_in.items()}
    WS.world_step(tables, base, fixed_rot=fixed_rot)
    gen = torch.Generator().manual_seed(0)
    worst = {k: 0.0 for k in STATE_KEYS}
    for _ in range(trials):
        pert = {k: v.clone() for k, v in state_in.items()}
        for k in ("pos", "rot"):
            sign = torch.randint(0, 3, pert[k].shape, generator=gen).float() - 1.0
            pert[k] = pert[k] * (1.0 + sign * 2.0**-23)
        WS.world_step(tables, pert, fixed_rot=fixed_rot)
        for k in STATE_KEYS:
            worst[k] = max(worst[k], float((pert[k] - base[k]).abs().max()))
    return worst


def _device_tables(tables, fixed_rot, device, mapping=None, ent_gravity=None):
    dt = _native.DeviceTables(tables, None, device, mapping=mapping)
    for k, v in fixed_rot.items():
        dt.joint_rot[:, k] = v.reshape(-1).to(device)
    for e, g in (ent_gravity or {}).items():
        dt.ent_gravity[:, e] = g.to(device)
    return dt


def _close(got, want, atol=ATOL):
    err = (got.cpu() - want).abs()
    bound = atol + RTOL * want.abs()
    return bool((err <= bound).all()), float(err.max())


@pytest.mark.parametrize("name", golden_names())
def test_world_step_vs_reference_golden(name):
    if _native.ARITH == "fast" and name == "crafted_clamps":
        # the reason the fast-arithmetic build is opt-in: torques over a small moment of inertia land at 1.2x
        # the tolerance (profiles/r2c_parity_fast.txt)
        pytest.xfail("VMAS_B200_ARITH=fast leaves the 1e-4 contract on this world")
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    worst = 0.0
    for t, state_in, fixed_rot, want in teacher_forced_steps(fix):
        dt = (
            _device_tables(tables, fixed_rot, device, ent_gravity=state_in.get("ent_gravity"))
            if (t == 0 or fixed_rot or "ent_gravity" in state_in)
            else dt
        )
        slab = _Slab(state_in, device)
        n = _native.world_step(lib, dt, slab)
        assert n >= 1
        sens = _ulp_sensitivity(tables, state_in, fixed_rot) if name in JOINT_WORLDS else None
        for k in STATE_KEYS:
            atol = ATOL + (4.0 * sens[k] if sens else 0.0)
            ok, err = _close(slab.t[k], want[k], atol)
            worst = max(worst, err)
            assert ok, f"{name} step {t} field {k}: max |err| {err} (atol {atol:.2e})"
    print(f"{name}: max |err| vs reference {worst:.3e}")


@pytest.mark.parametrize("name", ["balance", "pollock", "flocking", "waterfall"])
def test_world_step_vs_live_oracle(name):
    """Same inputs through the CPU oracle on this box and through the kernels."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    for t, state_in, fixed_rot, _ in teacher_forced_steps(fix):
        if t % 5:
            continue
        cpu = {k: v.clone() for k, v in state_in.items()}
        WS.world_step(tables, cpu, fixed_rot=fixed_rot)
        dt = _device_tables(tables, fixed_rot, device, ent_gravity=state_in.get("ent_gravity"))
        slab = _Slab(state_in, device)
        _native.world_step(lib, dt, slab)
        sens = _ulp_sensitivity(tables, state_in, fixed_rot) if name in JOINT_WORLDS else None
        for k in STATE_KEYS:
            atol = ATOL + (4.0 * sens[k] if sens else 0.0)
            ok, err = _close(slab.t[k], cpu[k], atol)
            assert ok, f"{name} step {t} field {k}: max |err| {err} (atol {atol:.2e})"


@pytest.mark.parametrize("name", ["navigation", "flocking"])
def test_fused_substeps_equal_single_substep_launches(name):
    """Sphere-only worlds fuse all S substeps in one launch; splitting must not change bits."""
    fix, desc, tables = load(name)
    assert tables.spheres_only and desc.substeps > 1
    lib = _native.load()
    device = torch.device("cuda:0")
    _, state_in, _, _ = next(teacher_forced_steps(fix))
    dt = _device_tables(tables, {}, device)
    fused = _Slab(state_in, device)
    assert _native.world_step(lib, dt, fused) == 1
    split = _Slab(state_in, device)
    for s in range(desc.substeps):
        _native.world_substeps(lib, dt, split, s, 1)
    for k in STATE_KEYS:
        assert same_result(fused.t[k], split.t[k])


@pytest.mark.parametrize("name", golden_names())
def test_thread_per_env_and_lanes_per_env_agree_bitwise(name):
    """Two independent thread mappings of the same arithmetic must produce identical bits."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    for t, state_in, fixed_rot, _ in teacher_forced_steps(fix):
        if t % 4:
            continue
        outs = []
        mappings = ["thread_per_env", "lanes_per_env"]
        if _native.DeviceTables(tables, None, device).mapping == "specialized":
            mappings.append("specialized")  # world-specialised, register-resident kernel
        for mapping in mappings:
            dt = _device_tables(tables, fixed_rot, device, mapping=mapping, ent_gravity=state_in.get("ent_gravity"))
            assert dt.mapping == mapping
            slab = _Slab(state_in, device)
            _native.world_step(lib, dt, slab)
            outs.append(slab)
        for other, mapping in zip(outs[1:], mappings[1:]):
            for k in STATE_KEYS:
                assert same_result(outs[0].t[k], other.t[k]), f"{name} step {t} field {k} ({mapping})"


@pytest.mark.parametrize("name", golden_names())
def test_tile_kernel_agrees_bitwise(name):
    """The warp-tile kernel (a warp owns 32 envs, compacted narrow phase; csrc/spec_tile_kernel.cuh)
    against the thread-per-env specialised kernel: identical bits, every golden world that has a
    tile kernel, whole steps (all substeps, broad phase included) and a batch that is not a
    multiple of the tile."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    auto = _native.DeviceTables(tables, None, device)
    if auto.specialization < 0 or not lib.vmas_b200_specialization_has_tile(auto.specialization):
        pytest.skip("no tile kernel for this world")
    for t, state_in, fixed_rot, _ in teacher_forced_steps(fix):
        if t % 3:
            continue
        n_envs = state_in["pos"].shape[0]
        for rows in (None, n_envs - 5):  # the whole batch, and one whose last tile is not full
            state = state_in if rows is None else {k: (v[:rows] if torch.is_tensor(v) else v) for k, v in state_in.items()}
            outs = []
            for mapping in ("specialized", "tile"):
                dt = _device_tables(tables, fixed_rot, device, mapping=mapping, ent_gravity=state.get("ent_gravity"))
                assert dt.mapping == mapping
                if rows is not None:
                    dt.cfg.batch_dim = rows
                slab = _Slab(state, device)
                _native.world_step(lib, dt, slab)
                outs.append(slab)
            for k in STATE_KEYS:
                assert same_result(outs[0].t[k], outs[1].t[k]), f"{name} step {t} field {k} rows {rows}"


@pytest.mark.parametrize("name", ["balance", "transport", "navigation", "flocking"])
def test_env_scheduling_changes_no_bit(name, monkeypatch):
    """Scheduling the envs by contact signature (``vmas_b200_build_env_order``: thread t steps env
    order[t]) must not change any env's result: same inputs stepped with the identity order and with
    the order built from the recorded signatures, bit for bit; and the order is a permutation."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    monkeypatch.setattr(_native, "ENV_REORDER_EVERY", 8)  # opt-in feature: off by default
    steps = list(teacher_forced_steps(fix))
    # a batch of 4096 envs stitched from different golden steps (different contact patterns)
    reps = 4096 // desc.batch_dim
    picks = [steps[(7 * i) % len(steps)][1] for i in range(reps)]
    state = {k: torch.cat([p[k] for p in picks]) for k in STATE_KEYS}
    B = state["pos"].shape[0]
    old = desc.batch_dim
    desc.batch_dim = B
    try:
        dt = _native.DeviceTables(tables, None, device, mapping="specialized")
    finally:
        desc.batch_dim = old
    assert dt.env_order is not None and dt.env_signature is not None
    first = _Slab(state, device)
    _native.world_step(lib, dt, first)  # identity order; records the signatures
    assert _native.build_env_order(lib, dt) == 1
    torch.cuda.synchronize()
    order = dt.env_order.long()
    assert torch.equal(torch.sort(order).values, torch.arange(B, device=device)), "not a permutation"
    if bool((dt.env_signature != dt.env_signature[0]).any()):
        assert not torch.equal(order, torch.arange(B, device=device)), "signatures differ but the order is the identity"
    second = _Slab(state, device)
    _native.world_step(lib, dt, second)
    for k in STATE_KEYS:
        assert torch.equal(first.t[k], second.t[k]), f"{name}: {k} changed under env scheduling"
    # same signatures again: sorting an already grouped batch keeps it a permutation
    _native.build_env_order(lib, dt)
    torch.cuda.synchronize()
    assert torch.equal(torch.sort(dt.env_order.long()).values, torch.arange(B, device=device))


@pytest.mark.parametrize("name", ["give_way", "crafted_clamps", "waterfall", "reverse_transport", "crafted_crowd"])
def test_runtime_specialisation_agrees_bitwise(name):
    """A world without a preset gets its specialised kernels compiled at run time (jit.py: nvcc on this
    box, cached); they must produce the bits of the generic kernels (worlds with joints, hollow boxes,
    action clamps, 70 entities)."""
    from vectorizedmultiagentsimulator_b200 import codegen, jit

    if not jit.available():
        pytest.skip("no nvcc on this box / JIT switched off")
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    job = jit.request(desc)
    if job is None:
        pytest.skip("world is not specialisable")
    assert job.done.wait(timeout=300) and job.error is None, job.error
    assert lib.vmas_b200_find_specialization(codegen.world_hash(desc)) == job.index
    for t, state_in, fixed_rot, _ in teacher_forced_steps(fix):
        if t % 3:
            continue
        outs = []
        for mapping in ("thread_per_env", "specialized"):
            dt = _device_tables(tables, fixed_rot, device, mapping=mapping, ent_gravity=state_in.get("ent_gravity"))
            assert dt.mapping == mapping
            slab = _Slab(state_in, device)
            _native.world_step(lib, dt, slab)
            outs.append(slab)
        for k in STATE_KEYS:
            assert same_result(outs[0].t[k], outs[1].t[k]), f"{name} step {t} field {k}"


def test_config_worlds_have_specialised_kernels():
    lib = _native.load()
    assert lib.vmas_b200_num_specializations() >= 4
    for name in ("balance", "transport", "navigation", "flocking"):
        _, _, tables = load(name)
        dt = _native.DeviceTables(tables, None, torch.device("cuda:0"))
        assert dt.mapping in ("specialized", "tile") and dt.specialization >= 0, name


@pytest.mark.parametrize("mapping", ["thread_per_env", "lanes_per_env"])
def test_world_step_vs_reference_golden_both_mappings(mapping):
    for name in ("balance", "pollock", "waterfall"):
        fix, desc, tables = load(name)
        lib = _native.load()
        device = torch.device("cuda:0")
        for t, state_in, fixed_rot, want in teacher_forced_steps(fix):
            if t > 6:
                break
            dt = _device_tables(tables, fixed_rot, device, mapping=mapping, ent_gravity=state_in.get("ent_gravity"))
            slab = _Slab(state_in, device)
            _native.world_step(lib, dt, slab)
            sens = _ulp_sensitivity(tables, state_in, fixed_rot) if name in JOINT_WORLDS else None
            for k in STATE_KEYS:
                atol = ATOL + (4.0 * sens[k] if sens else 0.0)
                ok, err = _close(slab.t[k], want[k], atol)
                assert ok, f"{mapping} {name} step {t} field {k}: max |err| {err}"


def test_step_is_deterministic_and_mask_is_restored():
    fix, desc, tables = load("pollock")
    lib = _native.load()
    device = torch.device("cuda:0")
    _, state_in, _, _ = next(teacher_forced_steps(fix))
    dt = _device_tables(tables, {}, device)
    outs = []
    for _ in range(3):
        slab = _Slab(state_in, device)
        n = _native.world_step(lib, dt, slab)
        assert n == 2 * desc.substeps  # broad phase + substep kernel per substep
        outs.append({k: slab.t[k].clone() for k in STATE_KEYS})
        assert int(dt.mask.abs().sum()) == 0, "the pair mask must be left cleared"
    for k in STATE_KEYS:
        assert torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k])


def test_broad_phase_mask_matches_oracle():
    fix, desc, tables = load("pollock")
    lib = _native.load()
    device = torch.device("cuda:0")
    _, state_in, _, _ = next(teacher_forced_steps(fix))
    dt = _device_tables(tables, {}, device)
    slab = _Slab(state_in, device)
    _native.broad_phase(lib, dt, slab)
    words = dt.mask.cpu().numpy().astype("uint32")
    dt.mask.zero_()
    for bit, item in enumerate(tables.masked_items):
        want = WS.broad_phase_active(tables, int(item), state_in["pos"])
        got = bool((words[bit // 32] >> (bit % 32)) & 1)
        assert got == want, f"mask bit {bit} (item {item})"


@pytest.mark.parametrize("name", [n for n in golden_names() if load(n)[0]["lidar"]])
def test_lidar_vs_reference_golden(name):
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    worst = 0.0
    for rec in fix["lidar"]:
        st = dict(fix["steps"][rec["step"]]["out"])
        slab = _Slab(st, device)
        targets = torch.tensor(rec["targets"] or [0], dtype=torch.int32, device=device)
        angles = rec["angles"].to(device).contiguous()
        out = torch.empty_like(angles)
        _native.cast_rays(lib, dt, slab, rec["src"], targets, len(rec["targets"]), angles, rec["src"], rec["max_range"], out)
        ok, err = _close(out, rec["out"], 1e-5)
        worst = max(worst, err)
        assert ok, f"{name} lidar src {rec['src']} step {rec['step']}: max |err| {err}"
        # the generic entry (angles already in the world frame) must agree bit for bit
        world_angles = (rec["angles"] + st["rot"][:, rec["src"]].unsqueeze(-1)).to(device).contiguous()
        out2 = torch.empty_like(angles)
        _native.cast_rays(lib, dt, slab, rec["src"], targets, len(rec["targets"]), world_angles, None, rec["max_range"], out2)
        assert same_result(out, out2)
    print(f"{name}: lidar max |err| {worst:.3e}")


@pytest.mark.parametrize("name", golden_names())
def test_queries_vs_reference_golden(name):
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    slab = _Slab(fix["final_state"], device)
    B = desc.batch_dim
    for q in fix["queries"]:
        d = torch.empty(B, device=device)
        _native.pair_query(lib, dt, slab, q["a"], q["b"], 0, d)
        # overlap flips the box-sphere distance to -1: compare where both sides agree on overlap
        o = torch.empty(B, dtype=torch.bool, device=device)
        _native.pair_query(lib, dt, slab, q["a"], q["b"], 1, o)
        near_boundary = (q["distance"].abs() < 1e-5)
        assert torch.equal(o.cpu() | near_boundary, q["overlap"] | near_boundary)
        same = o.cpu() == q["overlap"]
        ok, err = _close(d.cpu()[same], q["distance"][same], 1e-5)
        assert ok, f"{name} distance {q['a']}-{q['b']}: {err}"
        pd = torch.empty(B, device=device)
        _native.point_query(lib, dt, slab, q["a"], q["point"].to(device).contiguous(), pd)
        ok, err = _close(pd, q["point_distance"], 1e-5)
        assert ok, f"{name} point distance {q['a']}: {err}"


def test_large_batch_properties():
    """BASELINE size (balance, 32768 envs): tiling the golden state must reproduce it per tile."""
    fix, desc, tables = load("balance")
    lib = _native.load()
    device = torch.device("cuda:0")
    _, state_in, _, want = next(teacher_forced_steps(fix))
    reps = 32768 // desc.batch_dim
    big = {k: v.repeat(reps, *([1] * (v.dim() - 1))) for k, v in state_in.items()}
    desc.batch_dim = 32768
    dt = _device_tables(tables, {}, device)
    slab = _Slab(big, device)
    _native.world_step(lib, dt, slab)
    small = _Slab(state_in, device)
    desc.batch_dim = 64
    dt_small = _device_tables(tables, {}, device)
    _native.world_step(lib, dt_small, small)
    for k in STATE_KEYS:
        tiles = slab.t[k].reshape(reps, 64, *slab.t[k].shape[1:])
        assert torch.equal(tiles[0], small.t[k])
        assert torch.equal(tiles, tiles[0:1].expand_as(tiles)), f"{k}: envs are not independent"


@pytest.mark.parametrize("name", [n for n in golden_names() if load(n)[0]["lidar"]])
def test_batched_lidar_equals_per_sensor_launches(name):
    """vmas_b200_cast_rays_batched == one vmas_b200_cast_rays per sensor, bit for bit."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    by_step = {}
    for rec in fix["lidar"]:
        by_step.setdefault(rec["step"], []).append(rec)
    checked = 0
    for step, recs in by_step.items():
        n_rays = recs[0]["angles"].shape[-1]
        recs = [r for r in recs if r["angles"].shape[-1] == n_rays and r["targets"]]
        # the batched entry takes sensor-frame angles shared by the whole batch
        recs = [r for r in recs if bool((r["angles"] == r["angles"][:1]).all())]
        if not recs:
            continue
        slab = _Slab(dict(fix["steps"][step]["out"]), device)
        src = torch.tensor([r["src"] for r in recs], dtype=torch.int32, device=device)
        offs, flat = [0], []
        for r in recs:
            flat += r["targets"]
            offs.append(len(flat))
        target_off = torch.tensor(offs, dtype=torch.int32, device=device)
        targets = torch.tensor(flat, dtype=torch.int32, device=device)
        angles = torch.stack([r["angles"][0] for r in recs]).to(device).contiguous()
        max_range = torch.tensor([r["max_range"] for r in recs], dtype=torch.float32, device=device)
        B = desc.batch_dim
        out = torch.empty(len(recs), B, n_rays, device=device)
        _native.cast_rays_batched(lib, dt, slab, src, target_off, targets, angles, max_range, n_rays, out)
        from vectorizedmultiagentsimulator_b200.simulator import plan as P

        if all(int(tables.ent_i32[t, 0]) == P.SHAPE_SPHERE for t in flat):
            hinted = torch.empty_like(out)  # the sphere-only kernel must return the same bits
            _native.cast_rays_batched(lib, dt, slab, src, target_off, targets, angles, max_range, n_rays, hinted,
                                      flags=_native.RAYS_SPHERE_TARGETS)
            assert same_result(hinted, out), f"{name} step {step}: sphere-only LIDAR kernel differs"
        for q, r in enumerate(recs):
            one = torch.empty(B, n_rays, device=device)
            t = torch.tensor(r["targets"], dtype=torch.int32, device=device)
            _native.cast_rays(lib, dt, slab, r["src"], t, len(r["targets"]), r["angles"].to(device).contiguous(),
                              r["src"], r["max_range"], one)
            assert same_result(out[q], one), f"{name} step {step} sensor {q}"
            ok, err = _close(out[q], r["out"], 1e-5)
            assert ok, f"{name} step {step} sensor {q}: max |err| {err}"
            checked += 1
    assert checked > 0


@pytest.mark.parametrize("name", golden_names())
def test_batched_pair_query_equals_per_pair_launches(name):
    fix, desc, tables = load(name)
    if not fix["queries"]:
        pytest.skip("fixture has no pair queries")
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    slab = _Slab(fix["final_state"], device)
    B = desc.batch_dim
    pairs = torch.tensor([[q["a"], q["b"]] for q in fix["queries"]], dtype=torch.int32, device=device)
    K = pairs.shape[0]
    dist = torch.empty(K, B, device=device)
    over = torch.empty(K, B, dtype=torch.bool, device=device)
    centre = torch.empty(K, B, device=device)
    _native.pair_query_batched(lib, dt, slab, pairs, 0, dist)
    _native.pair_query_batched(lib, dt, slab, pairs, 1, over)
    _native.pair_query_batched(lib, dt, slab, pairs, 2, centre)
    pos = slab.t["pos"]
    for k, q in enumerate(fix["queries"]):
        d = torch.empty(B, device=device)
        o = torch.empty(B, dtype=torch.bool, device=device)
        _native.pair_query(lib, dt, slab, q["a"], q["b"], 0, d)
        _native.pair_query(lib, dt, slab, q["a"], q["b"], 1, o)
        assert same_result(dist[k], d) and same_result(over[k], o), f"{name} pair {k}"
        want = torch.linalg.vector_norm(pos[:, q["a"]] - pos[:, q["b"]], dim=-1)
        assert torch.allclose(centre[k], want, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["balance", "pollock"])
def test_gather_observations_equals_torch_expressions(name):
    """vmas_b200_gather_observations: COPY / DIFF / REMAINDER columns == the torch expressions
    a scenario would concatenate, bit for bit; SKIP columns are left untouched."""
    fix, desc, tables = load(name)
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    slab = _Slab(fix["final_state"], device)
    pos, vel, rot, ang_vel = (slab.t[k] for k in ("pos", "vel", "rot", "ang_vel"))
    B, E = pos.shape[0], pos.shape[1]
    rows = min(3, E)
    N = _native
    table, want = [], []
    for r in range(rows):
        e, other = r, (r + 1) % E
        cols, parts = [], []
        for k in range(2):
            cols.append((N.OBS_COPY, (N.OBS_POS << 24) | (2 * e + k), 0, 0))
        parts.append(pos[:, e])
        for k in range(2):
            cols.append((N.OBS_DIFF, (N.OBS_VEL << 24) | (2 * e + k), (N.OBS_VEL << 24) | (2 * other + k), 0))
        parts.append(vel[:, e] - vel[:, other])
        for k in range(2):
            cols.append((N.OBS_DIFF, (N.OBS_POS << 24) | (2 * other + k), (N.OBS_POS << 24) | (2 * e + k), 0))
        parts.append(pos[:, other] - pos[:, e])
        cols.append((N.OBS_SKIP, 0, 0, 0))
        parts.append(torch.full((B, 1), -7.0, device=device))
        cols.append((N.OBS_COPY, (N.OBS_ANG_VEL << 24) | e, 0, 0))
        parts.append(ang_vel[:, e : e + 1])
        for modulus in (torch.pi, -2.0):
            bits = torch.tensor(modulus, dtype=torch.float32).view(torch.int32).item()
            cols.append((N.OBS_REMAINDER, (N.OBS_ROT << 24) | e, 0, bits))
            parts.append(rot[:, e : e + 1] % modulus)
        table.append(cols)
        want.append(torch.cat(parts, dim=-1))
    want = torch.stack(want)
    width = want.shape[-1]
    columns = torch.tensor(table, dtype=torch.int32, device=device).contiguous()
    out = torch.full((rows, B, width), -7.0, device=device)
    _native.gather_observations(lib, dt, slab, columns, rows, width, out)
    assert torch.equal(out, want)


def test_batched_lidar_strided_output_and_range_flip():
    """Readings scattered into columns of a wider block (and max_range - d) == the dense result."""
    fix, desc, tables = load("navigation")
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    recs = [r for r in fix["lidar"] if r["step"] == fix["lidar"][0]["step"]]
    slab = _Slab(dict(fix["steps"][recs[0]["step"]]["out"]), device)
    n_rays = recs[0]["angles"].shape[-1]
    src = torch.tensor([r["src"] for r in recs], dtype=torch.int32, device=device)
    offs, flat = [0], []
    for r in recs:
        flat += r["targets"]
        offs.append(len(flat))
    target_off = torch.tensor(offs, dtype=torch.int32, device=device)
    targets = torch.tensor(flat, dtype=torch.int32, device=device)
    angles = torch.stack([r["angles"][0] for r in recs]).to(device).contiguous()
    max_range = torch.tensor([r["max_range"] for r in recs], dtype=torch.float32, device=device)
    B, Q = desc.batch_dim, len(recs)
    dense = torch.empty(Q, B, n_rays, device=device)
    _native.cast_rays_batched(lib, dt, slab, src, target_off, targets, angles, max_range, n_rays, dense)
    width, col = n_rays + 5, 3
    for flags in (0, _native.RAYS_RANGE_MINUS_DISTANCE):
        block = torch.full((Q, B, width), -1.0, device=device)
        out_off = torch.tensor([q * B * width + col for q in range(Q)], dtype=torch.int64, device=device)
        _native.cast_rays_batched(
            lib, dt, slab, src, target_off, targets, angles, max_range, n_rays, block, out_off, width, flags
        )
        want = max_range.view(-1, 1, 1) - dense if flags else dense
        assert torch.equal(block[:, :, col : col + n_rays], want)
        assert bool((block[:, :, :col] == -1.0).all()) and bool((block[:, :, col + n_rays :] == -1.0).all())


def test_distance_shaping_equals_torch_expressions():
    """vmas_b200_distance_shaping: dist / rew / carried shaping == the torch formulation."""
    fix, desc, tables = load("navigation")
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    slab = _Slab(fix["final_state"], device)
    pos = slab.t["pos"]
    B, E = pos.shape[0], pos.shape[1]
    pairs_list = [(i, (i + 3) % E) for i in range(E)] + [(0, 0)]
    pairs = torch.tensor(pairs_list, dtype=torch.int32, device=device)
    K = len(pairs_list)
    gen = torch.Generator().manual_seed(5)
    prev0 = torch.rand(K, B, generator=gen).to(device)
    prev = prev0.clone()
    dist = torch.empty(K, B, device=device)
    rew = torch.empty(K, B, device=device)
    factor = 0.7
    _native.distance_shaping(lib, dt, slab, pairs, factor, prev, dist, rew)
    want_dist = torch.stack([torch.linalg.vector_norm(pos[:, a] - pos[:, b], dim=-1) for a, b in pairs_list])
    assert torch.allclose(dist, want_dist, rtol=2e-7, atol=0)
    # given the distance, the rest is exact fp32 arithmetic in the reference's order
    assert torch.equal(prev, dist * factor)
    assert torch.equal(rew, prev0 - dist * factor)
    assert bool((dist[-1] == 0).all())
    # dist is optional
    prev2 = prev0.clone()
    rew2 = torch.empty(K, B, device=device)
    _native.distance_shaping(lib, dt, slab, pairs, factor, prev2, None, rew2)
    assert torch.equal(rew2, rew) and torch.equal(prev2, prev)


@pytest.mark.parametrize("name", golden_names())
def test_sphere_only_pair_kernel_equals_general_kernel(name):
    """The VMAS_QUERY_SPHERES fast path returns the bits of the general kernel."""
    from vectorizedmultiagentsimulator_b200.simulator import plan as P

    fix, desc, tables = load(name)
    spheres = [i for i in range(desc.n_entities) if int(tables.ent_i32[i, 0]) == P.SHAPE_SPHERE]
    if len(spheres) < 2:
        pytest.skip("fewer than two spheres")
    lib = _native.load()
    device = torch.device("cuda:0")
    dt = _device_tables(tables, {}, device)
    slab = _Slab(fix["final_state"], device)
    B = desc.batch_dim
    pair_list = [(a, b) for i, a in enumerate(spheres) for b in spheres[i + 1 :]][:40] + [(spheres[0], spheres[0])]
    pairs = torch.tensor(pair_list, dtype=torch.int32, device=device)
    for mode, dtype in ((0, torch.float32), (1, torch.bool), (2, torch.float32)):
        general = torch.empty(len(pair_list), B, dtype=dtype, device=device)
        fast = torch.empty_like(general)
        _native.pair_query_batched(lib, dt, slab, pairs, mode, general)
        _native.pair_query_batched(lib, dt, slab, pairs, mode | _native.QUERY_SPHERES, fast)
        assert same_result(general, fast), f"{name} mode {mode}"
