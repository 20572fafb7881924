#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>gpurun_out/r2z_bench.err > gpurun_out/r2z_bench.json
VMAS_BENCH_PINNED_ACTIONS=1 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2z_bench_pinned.json
python - <<'PY'
import json
for f in ("r2z_bench", "r2z_bench_pinned"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e ms %.4f e2e %.3e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]), d["config"].get("bracket_us"))
    except Exception as e:
        print(f, "failed", e)
PY
