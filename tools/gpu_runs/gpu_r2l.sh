#!/bin/bash
mkdir -p gpurun_out
python tools/debug_discrete.py 2>&1 | tail -6 | tee gpurun_out/r2l_debug.txt
python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r2l_env_tests.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2l_bench20.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2l_bench20b.json
python - <<'PY' | tee gpurun_out/r2l_summary.txt
import json
for f in ("r2l_bench20", "r2l_bench20b"):
    d = json.loads(open("gpurun_out/" + f + ".json").read())
    print(f, "%.3e" % d["value"], "%.3e" % d["e2e"]["value"], d["e2e"]["bracket_ms_first5_last5"])
PY
