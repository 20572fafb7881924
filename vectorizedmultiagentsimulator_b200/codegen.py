"""Generates ``csrc/generated/specializations.cuh``: constexpr world tables for the specialised kernel.

``csrc/spec_kernel.cuh`` is a hand-written template that turns a compile-time world description
into a fully unrolled, register-resident substep kernel.  This module only emits the *data* it
is instantiated with — one ``struct World_<hash>`` per preset world (the BASELINE.json configs
and the shipped scenarios' defaults) — plus the registry the C ABI looks specialisations up in
by a 64-bit hash of the world description.  Worlds without a specialisation run on the generic
table-driven kernels; both produce identical bits (tests/test_cabi_gpu.py).

    python -m vectorizedmultiagentsimulator_b200.codegen        # rewrite the generated header
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np

from .simulator import plan as P

HERE = os.path.dirname(os.path.abspath(__file__))
GENERATED = os.path.join(HERE, "csrc", "generated", "specializations.cuh")

#: worlds that get an ahead-of-time specialisation: (scenario, kwargs[, tuning]).
#: tuning["min_blocks"]: resident blocks per SM the register allocator leaves room for (default: the
#: build's SPEC_MIN_BLOCKS = 8, i.e. <= 128 registers at 64 threads per block).  Measured per world
#: on B200 (profiles/r1g_variant_bench.txt): stock transport runs 11 % faster at 1 Mi envs with 12
#: (80 registers, 24 warps / SM) and the same at 32768; navigation loses 28 % there, balance and
#: flocking gain 3-7 % at 1 Mi envs but lose 9-12 % at 32768, so they stay at the default.
PRESETS: List[Tuple] = [
    ("balance", dict(n_agents=4)),  # BASELINE.json configs[0], [1]
    ("balance", dict()),
    ("transport", dict(n_agents=4), dict(min_blocks=12)),
    ("transport", dict(n_agents=4, n_lines=2, substeps=3)),  # BASELINE.json configs[2] variant
    ("navigation", dict(n_agents=8)),  # configs[3]
    ("navigation", dict()),
    ("flocking", dict(n_agents=5)),  # configs[4]
    ("flocking", dict()),
]

#: specialisation is skipped for worlds whose unrolled code would be unreasonably large
MAX_ENTITIES = 24
MAX_ITEM_COST = 400  # in units of one segment/segment test


def world_hash(desc: P.WorldDescription) -> int:
    """FNV-1a 64 of everything that shapes the kernel (not the batch size, not entity names)."""
    d = json.loads(desc.to_json())
    d.pop("batch_dim")
    for e in d["entities"]:
        e.pop("name")
        if not e.get("gravity_per_env"):
            e.pop("gravity_per_env", None)  # absent in descriptions written before the field existed
    blob = json.dumps(d, sort_keys=True).encode()
    h = 0xCBF29CE484222325
    for byte in blob:
        h ^= byte
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _item_cost(kind: int) -> int:
    return {P.K_JOINT: 1, P.K_SS: 1, P.K_LS: 1, P.K_LL: 1, P.K_BS: 1, P.K_BL: 4, P.K_BB: 32}[kind]


def specializable(desc: P.WorldDescription) -> bool:
    if desc.n_entities > MAX_ENTITIES or desc.n_entities == 0:
        return False
    if any(e.get("gravity_per_env") for e in desc.entities):
        return False
    return sum(_item_cost(it["kind"]) for it in desc.items) <= MAX_ITEM_COST


def _f(x) -> str:
    v = float(np.float32(x))
    if v != v or v in (float("inf"), float("-inf")):
        raise ValueError(f"non-finite constant {x} in world description")
    s = f"{v:.9g}"
    if "e" not in s and "." not in s:
        s += ".0"
    return s + "f"


def emit_world(desc: P.WorldDescription, label: str, tuning: Dict = None) -> Tuple[str, str, int]:
    """C++ text of one world struct.  Returns (struct name, text, hash)."""
    min_blocks = (tuning or {}).get("min_blocks", "SPEC_MIN_BLOCKS")
    tables = P.build_tables(desc)
    h = world_hash(desc)
    name = f"World_{h:016x}"
    E, NI = desc.n_entities, len(desc.items)
    ef, ei = tables.ent_f32, tables.ent_i32
    lines = [f"// {label}: E={E} items={NI} substeps={desc.substeps}", f"struct {name} {{"]
    lines.append(f"  static constexpr int E = {E}, A = {desc.n_agents}, NI = {NI}, N_JOINTS = {tables.n_joints};")
    lines.append(
        f"  static constexpr int MASK_WORDS = {(tables.n_masked + 31) // 32}, BLOCK = SPEC_BLOCK, "
        f"MIN_BLOCKS = {min_blocks};"
    )
    d = desc
    lines.append(
        "  static constexpr CfgC cfg = {"
        f"{d.substeps}, {int(d.x_semidim is not None)}, {int(d.y_semidim is not None)}, "
        f"{int(any(g != 0.0 for g in d.gravity))}, {_f(d.dt / d.substeps)}, {_f(d.x_semidim or 0.0)}, "
        f"{_f(d.y_semidim or 0.0)}, {_f(d.collision_force)}, {_f(d.joint_force)}, "
        f"{_f(d.torque_constraint_force)}, {_f(d.contact_margin)}, {_f(d.gravity[0])}, {_f(d.gravity[1])}}};"
    )
    lines.append(f"  static constexpr EntC ent[{max(E, 1)}] = {{")
    for e in range(E):
        r = ef[e]
        cols = [
            P.EF_D0, P.EF_D1, P.EF_MASS, P.EF_INERTIA, P.EF_DRAG_MULT, P.EF_LIN_FRIC, P.EF_ANG_FRIC, P.EF_GRAV_X,
            P.EF_GRAV_Y, P.EF_MAX_SPEED, P.EF_V_RANGE, P.EF_MAX_F, P.EF_F_RANGE, P.EF_MAX_T, P.EF_T_RANGE,
            P.EF_CIRC_R, P.EF_R_PLUS_LMD,
        ]
        vals = ", ".join(_f(r[c]) for c in cols)
        lines.append(f"      {{{int(ei[e, 0])}, {int(ei[e, 1])}, {int(ei[e, 2])}, {vals}}},  // {desc.entities[e]['name']}")
    lines.append("  };")
    lines.append(f"  static constexpr ItemC item[{max(NI, 1)}] = {{")
    if NI == 0:
        lines.append("      {0, 0, 0, 0, -1, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},")
    for k in range(NI):
        ii, f32 = tables.item_i32[k], tables.item_f32[k]
        flags = int(ii[3]) & 0xFF
        vals = ", ".join(
            _f(f32[c]) for c in (P.IF_DMIN_BASE, P.IF_AX, P.IF_AY, P.IF_BX, P.IF_BY, P.IF_DIST, P.IF_FIXED_ROT, P.IF_BROAD_THR)
        )
        lines.append(
            f"      {{{int(ii[0])}, {int(ii[1])}, {int(ii[2])}, {flags}, {int(tables.mask_slot[k])}, {vals}}},"
            f"  // {P.KIND_NAMES[int(ii[0])]}"
        )
    lines.append("  };")
    lines.append("};")
    return name, "\n".join(lines), h


def post_hash(cols, instrs, acts=()) -> int:
    """FNV-1a 64 of what a whole-step kernel does around the substeps: the observation plan's column table
    (int32 ``[rows, width, 4]`` or None), the step program's instructions ``[(op, dst, a, b, arg, imm)]`` with
    entity indices resolved, and the action ingest ``[(agent row, u_range x 2, u_multiplier x 2)]`` of the
    policy agents (empty: actions are ingested by a launch of their own)."""
    blob = json.dumps(
        [None if cols is None else [list(cols.shape), [int(x) for x in cols.reshape(-1)]],
         [[int(op), int(dst), int(a), int(b), int(arg), _f(imm)] for op, dst, a, b, arg, imm in instrs],
         [[int(agent)] + [_f(v) for v in rest] for agent, *rest in acts]]
    ).encode()
    h = 0xCBF29CE484222325
    for byte in blob:
        h ^= byte
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def fuse_value_columns(cols, buffer_sources, instrs):
    """The column table of an observation plan for a whole-step kernel: value columns (``observe.value`` of a
    program output; OP_BUFFER = 4, index into ``buffer_sources``) become reads of the program register that
    output stores (OP_REG = 5) — program and observation rows run in one thread there."""
    OP_BUFFER, OP_REG, STORES = 4, 5, (20, 21)
    if cols is None or not buffer_sources:
        return cols
    reg_of = {b: a for op, _, a, b, _, _ in instrs if op in STORES}
    cols = cols.copy()
    for row in cols.reshape(-1, 4):
        if row[0] == OP_BUFFER:
            row[0], row[1] = OP_REG, reg_of[buffer_sources[int(row[1])]._slot]
    return cols


def emit_post(cols, instrs, acts=()) -> Tuple[str, str, int]:
    """C++ text of one epilogue (+ ingest prologue) struct (``spec_epilogue`` / ``spec_ingest`` in
    csrc/spec_kernel.cuh).  Returns (name, text, hash)."""
    h = post_hash(cols, instrs, acts)
    name = f"Post_{h:016x}"
    rows, width = (0, 0) if cols is None else (int(cols.shape[0]), int(cols.shape[1]))
    lines = [f"struct {name} {{"]
    lines.append(f"  static constexpr int N_PROG = {len(instrs)}, OBS_ROWS = {rows}, OBS_WIDTH = {width}, N_ACT = {len(acts)};")
    lines.append(f"  static constexpr ActC act[{max(len(acts), 1)}] = {{")
    for agent, r0, r1, m0, m1 in acts:
        lines.append(f"      {{{int(agent)}, {_f(r0)}, {_f(r1)}, {_f(m0)}, {_f(m1)}}},")
    if not acts:
        lines.append("      {0, 0.f, 0.f, 0.f, 0.f},")
    lines.append("  };")
    lines.append(f"  static constexpr ProgC prog[{max(len(instrs), 1)}] = {{")
    for op, dst, a, b, arg, imm in instrs:
        lines.append(f"      {{{int(op)}, {int(dst)}, {int(a)}, {int(b)}, {int(arg)}, {_f(imm)}}},")
    if not instrs:
        lines.append("      {0, 0, 0, 0, 0, 0.f},")
    lines.append("  };")
    lines.append(f"  static constexpr ObsColC obs[{max(rows * width, 1)}] = {{")
    if rows * width == 0:
        lines.append("      {0, 0, 0, 0.f},")
    else:
        flat = cols.reshape(-1, 4)
        for op, src, src2, par in flat:
            par_f = float(np.array([par], dtype=np.int32).view(np.float32)[0])
            lines.append(f"      {{{int(op)}, {int(src)}, {int(src2)}, {_f(par_f)}}},")
    lines.append("  };")
    lines.append("};")
    return name, "\n".join(lines), h


def preset_descriptions() -> List[Tuple[str, P.WorldDescription, Dict]]:
    """Builds every preset world on the CPU (construction only, no physics) and describes it."""
    import torch

    from . import scenarios

    out = []
    for scenario, kwargs, *tuning in PRESETS:
        sc = scenarios.load(scenario + ".py").Scenario()
        world = sc.env_make_world(1, torch.device("cpu"), **dict(kwargs))
        label = scenario + "(" + ", ".join(f"{k}={v}" for k, v in kwargs.items()) + ")"
        out.append((label, P.describe_world(world), tuning[0] if tuning else None))
    return out


def generate(path: str = GENERATED) -> List[Tuple[str, int]]:
    worlds, seen = [], set()
    for label, desc, tuning in preset_descriptions():
        if not specializable(desc):
            continue
        name, text, h = emit_world(desc, label, tuning)
        if h in seen:
            continue
        seen.add(h)
        worlds.append((label, name, text, h, desc))
    parts = [
        "// GENERATED by vectorizedmultiagentsimulator_b200/codegen.py — do not edit.",
        "// constexpr world tables the specialised substep kernel (spec_kernel.cuh) is instantiated with.",
        "#pragma once",
        '#include "../spec_kernel.cuh"',
        '#include "../spec_tile_kernel.cuh"',
        "",
        "namespace vmas {",
        "",
    ]
    for _, _, text, _, _ in worlds:
        parts += [text, ""]
    parts.append("#ifdef __CUDACC__")
    parts.append("static const SpecEntry kSpecs[] = {")
    for label, name, _, h, desc in worlds:
        parts.append(
            f'    {{0x{h:016x}ull, "{label}", {desc.n_entities}, {len(desc.items)}, &launch_spec<{name}>, '
            f"&launch_tile<{name}>, TileLayout<{name}>::SUPPORTED}},"
        )
    if not worlds:
        parts.append('    {0ull, "", 0, 0, nullptr, nullptr, false},')
    parts.append("};")
    parts.append(f"static const int kNumSpecs = {len(worlds)};")
    parts.append("#endif")
    # the same worlds as a type list, for code that instantiates a template per world (tests/hostsim)
    parts.append("#define VMAS_FOR_EACH_SPEC_WORLD(X) \\")
    parts.append(" \\\n".join(f"  X({i}, {name}, 0x{h:016x}ull)" for i, (_, name, _, h, _) in enumerate(worlds)))
    parts += ["", "}  // namespace vmas", ""]
    text = "\n".join(parts)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as fh:
            fh.write(text)
    return [(label, h) for label, _, _, h, _ in worlds]


if __name__ == "__main__":
    for label, h in generate():
        print(f"{h:016x}  {label}")
