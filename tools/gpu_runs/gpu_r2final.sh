#!/bin/bash
# Round-2 closing run: the full GPU suite in both arithmetic builds, smoke, the driver-style bench lines (both
# arms), and the ncu evidence for the one-kernel step (launch list of the bench command + one --set full capture).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2final_stages.log; }
stamp "start"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/r2final_stages.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2final_gpu_tests_exact.log 2>&1
stamp "gpu suite (exact) rc=$?"; tail -4 gpurun_out/r2final_gpu_tests_exact.log
VMAS_B200_ARITH=fast timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2final_gpu_tests_fast.log 2>&1
stamp "gpu suite (fast) rc=$?"; tail -6 gpurun_out/r2final_gpu_tests_fast.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2final_smoke.log 2>&1
stamp "smoke rc=$?"; tail -3 gpurun_out/r2final_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench_driver.json 2> gpurun_out/r2final_bench_driver.err
stamp "bench (driver-style) rc=$?"
timeout 300 python bench.py --steps 300 --warmup 20 > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err
stamp "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2final_bench_reference.json 2>/dev/null
stamp "bench reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2final_under_ncu.log 2>&1
stamp "ncu launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_env_kernel -s 2 -c 1 -o gpurun_out/r2final_step_env_kernel -f python tools/run_graph_steps.py balance 32768 n_agents=4 > gpurun_out/r2final_ncu.log 2>&1
stamp "ncu --set full rc=$?"
python tools/ncu_summary.py gpurun_out/r2final_step_env_kernel.ncu-rep > gpurun_out/r2final_step_env_kernel_ncu_full.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2final_step_env_kernel.ncu-rep 0x1000 >> gpurun_out/r2final_step_env_kernel_ncu_full.txt 2>&1
# the whole-step kernels compiled on this box: back into the snapshot's JIT cache (same toolchain, same source stamp)
mkdir -p gpurun_out/jit && cp vectorizedmultiagentsimulator_b200/csrc/generated/jit/*.so gpurun_out/jit/ 2>/dev/null
ls gpurun_out/jit | wc -l
python - <<'PY'
import json
for f in ("r2final_bench_driver", "r2final_bench", "r2final_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read())
        e = d.get("e2e", {})
        print(f, "value %.3e ms %.4f e2e %.3e launches %s frac %s" % (d["value"], d["ms_per_step"], e.get("value", 0), d.get("gpu_launches"), d.get("roofline", {}).get("frac")))
    except Exception as err:
        print(f, "failed", err)
PY
head -30 gpurun_out/r2final_step_env_kernel_ncu_full.txt
