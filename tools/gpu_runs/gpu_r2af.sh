#!/bin/bash
# after the epilogue-input prefetch: env tests in both builds, the bench lines, and the JIT cache of the box
mkdir -p gpurun_out/jit
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
VMAS_B200_ARITH=fast timeout 900 python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2af_bench.json
timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null > gpurun_out/r2af_bench_driver_style.json
timeout 300 python bench.py --config transport3 --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/r2af_bench_transport3.json
timeout 200 python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2af_timeline_balance.txt 2>&1
rm -f gpurun_out/jit/*; cp vectorizedmultiagentsimulator_b200/csrc/generated/jit/*.so gpurun_out/jit/ 2>/dev/null; ls gpurun_out/jit | wc -l
python - <<'PY'
import json
for f in ("r2af_bench", "r2af_bench_driver_style", "r2af_bench_transport3"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        print(f, "value %.3e ms %.4f e2e %.3e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]), d["config"].get("bracket_us"), d["config"].get("remeasured"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -v "Warning\|_warn" gpurun_out/r2af_timeline_balance.txt | cut -c1-140 | tail -4
