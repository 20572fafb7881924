#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for i in 1 2 3 4; do
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2ab_bench$i.json
done
python - <<'PY'
import json
for f in ("r2ab_bench1", "r2ab_bench2", "r2ab_bench3", "r2ab_bench4"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e e2e %.3e (%.4f ms)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"].get("bracket_us"), d["e2e"].get("bracket_us"))
    except Exception as e:
        print(f, "failed", e)
PY
