#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_env_gpu.py tests/test_reset_gpu.py tests/test_program_gpu.py tests/test_jit.py -q -x -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r2w_tests.txt
timeout 300 python tools/host_profile.py balance n_agents=4 2>&1 | head -3 > gpurun_out/r2w_host_profile_balance.txt
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>gpurun_out/r2w_bench.err > gpurun_out/r2w_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2w_bench20.json
VMAS_B200_INGEST_IN_KERNEL=0 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2w_bench_ingest_launch.json
timeout 300 python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2w_timeline_balance.txt 2>&1
cat gpurun_out/r2w_host_profile_balance.txt
python - <<'PY'
import json
for f in ("r2w_bench", "r2w_bench20", "r2w_bench_ingest_launch"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e ms %.4f e2e %.3e launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), d["config"].get("bracket_us"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2w_bench.err
tail -8 gpurun_out/r2w_timeline_balance.txt | cut -c1-130
