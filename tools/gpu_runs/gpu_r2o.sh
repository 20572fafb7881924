#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests/test_program_gpu.py tests/test_env_gpu.py tests/test_reset_gpu.py -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r2o_tests.txt
for i in 1 2; do python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2o_bench$i.json; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2o_bench20.json
python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2o_timeline_balance.txt 2>&1
python - <<'PY' | tee gpurun_out/r2o_summary.txt
import json
for f in ("r2o_bench1", "r2o_bench2", "r2o_bench20"):
    d = json.loads(open("gpurun_out/" + f + ".json").read())
    print(f, "value %.3e ms %.4f e2e %.3e launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
tail -10 gpurun_out/r2o_timeline_balance.txt | cut -c1-130
