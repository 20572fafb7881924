#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_cabi_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -8
VMAS_B200_ARITH=fast timeout 900 python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --config transport3 --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2ae_bench_transport3.json
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2ae_bench.json
timeout 300 python tools/step_timeline.py transport 16384 n_agents=4 n_lines=2 substeps=3 > gpurun_out/r2ae_timeline_transport3.txt 2>&1
python - <<'PY'
import json
for f in ("r2ae_bench_transport3","r2ae_bench"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e ms %.4f e2e %.3e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]), d["config"].get("bracket_us"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -v "Warning\|_warn" gpurun_out/r2ae_timeline_transport3.txt | cut -c1-150
