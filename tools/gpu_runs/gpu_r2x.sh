#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_reset_gpu.py tests/test_program_gpu.py tests/test_jit.py -q -x -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r2x_tests.txt
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>gpurun_out/r2x_bench.err > gpurun_out/r2x_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2x_bench20.json
for c in transport3 navigation flocking; do timeout 300 python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2x_bench_$c.json; done
python - <<'PY'
import json
for f in ("r2x_bench", "r2x_bench20", "r2x_bench_transport3", "r2x_bench_navigation", "r2x_bench_flocking"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e ms %.4f e2e %.3e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]), d["config"].get("bracket_us"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -v "^frame" gpurun_out/r2x_bench.err | tail -3
