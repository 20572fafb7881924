"""The CPU oracle against the committed reference roll-outs (runs anywhere, no GPU)."""
import pytest
import torch

from golden_util import golden_names, load, max_abs_err, teacher_forced_steps
from oracle import queries as Q
from oracle import world_step as WS

# The fixtures were produced on an AVX-512 host; on the same ISA the oracle is bit-identical.
# Other hosts may round transcendental ops differently in the last place.
TOL = 2e-6


@pytest.mark.parametrize("name", golden_names())
def test_world_step_matches_reference(name):
    fix, desc, tables = load(name)
    worst = 0.0
    for t, state, fixed_rot, want in teacher_forced_steps(fix):
        WS.world_step(tables, state, fixed_rot=fixed_rot)
        worst = max(worst, max_abs_err(state, want))
    assert worst <= TOL, f"{name}: max |err| {worst}"


@pytest.mark.parametrize("name", [n for n in golden_names() if load(n)[0]["lidar"]])
def test_lidar_matches_reference(name):
    fix, desc, tables = load(name)
    for rec in fix["lidar"]:
        st = fix["steps"][rec["step"]]["out"]
        got = Q.cast_rays(
            tables, st["pos"], st["rot"], rec["src"], rec["targets"],
            rec["angles"] + st["rot"][:, rec["src"]].unsqueeze(-1), rec["max_range"],
        )
        assert float((got - rec["out"]).abs().max()) <= TOL


@pytest.mark.parametrize("name", golden_names())
def test_queries_match_reference(name):
    fix, desc, tables = load(name)
    st = fix["final_state"]
    for q in fix["queries"]:
        d = Q.pair_distance(tables, st["pos"], st["rot"], q["a"], q["b"])
        assert float((d - q["distance"]).abs().max()) <= TOL
        assert torch.equal(Q.pair_overlap(tables, st["pos"], st["rot"], q["a"], q["b"]), q["overlap"])
        pd = Q.distance_from_point(tables, st["pos"], st["rot"], q["a"], q["point"])
        assert float((pd - q["point_distance"]).abs().max()) <= TOL
