#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests/test_env_gpu.py tests/test_reset_gpu.py tests/test_program_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r2r_tests.txt
python tools/host_profile.py balance n_agents=4 > gpurun_out/r2r_host_profile_balance.txt 2>&1
VMAS_B200_ONE_CALL_STEP=0 python tools/host_profile.py balance n_agents=4 2>&1 | head -3 > gpurun_out/r2r_host_profile_balance_3calls.txt
python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/r2r_bench.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2r_bench20.json
python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2r_timeline_balance.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:"post_step_kernel|ingest_broad_kernel" -s 8 -c 2 \
  -o gpurun_out/r2r_small_kernels -f python tools/run_steps.py balance 32768 n_agents=4 > gpurun_out/r2r_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2r_small_kernels.ncu-rep > gpurun_out/r2r_small_kernels_ncu.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2r_small_kernels.ncu-rep > gpurun_out/r2r_small_kernels_regions.txt 2>&1
head -4 gpurun_out/r2r_host_profile_balance.txt; cat gpurun_out/r2r_host_profile_balance_3calls.txt
python - <<'PY'
import json
for f in ("r2r_bench", "r2r_bench20"):
    d = json.loads(open("gpurun_out/" + f + ".json").read())
    print(f, "value %.3e ms %.4f e2e %.3e launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
tail -8 gpurun_out/r2r_timeline_balance.txt | cut -c1-130
tail -2 gpurun_out/r2r_ncu.log
