#!/bin/bash
# round 2, call G: the whole GPU suite (exact build) after device dynamics / JIT / forked observations;
# bench with the re-ordered e2e loop; step timeline.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2g_stages.log; }
stamp start
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2g_gpu_tests.log 2>&1
stamp "gpu suite rc=$?"; tail -25 gpurun_out/r2g_gpu_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/r2g_smoke.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
stamp "bench rc=$?"; cat gpurun_out/r2g_bench.json; tail -5 gpurun_out/r2g_bench.err
for cfg in transport3 navigation; do
timeout 400 python bench.py --config $cfg --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/r2g_bench_$cfg.json 2> gpurun_out/r2g_bench_$cfg.err
stamp "bench $cfg rc=$?"; cut -c1-600 gpurun_out/r2g_bench_$cfg.json; tail -3 gpurun_out/r2g_bench_$cfg.err
done
timeout 400 python bench.py --config flocking --total-envs 262144 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2g_bench_flocking.json 2> gpurun_out/r2g_bench_flocking.err
stamp "bench flocking rc=$?"; cut -c1-600 gpurun_out/r2g_bench_flocking.json; tail -3 gpurun_out/r2g_bench_flocking.err
timeout 100 python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2g_timeline_balance.txt 2>&1
timeout 100 python tools/step_timeline.py navigation 8192 n_agents=8 > gpurun_out/r2g_timeline_navigation.txt 2>&1
stamp "timelines"; cat gpurun_out/r2g_timeline_balance.txt | cut -c1-140
