#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2m_stages.log; }
stamp start
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2m_gpu_tests.log 2>&1
stamp "gpu suite rc=$?"; tail -15 gpurun_out/r2m_gpu_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log
timeout 400 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
stamp "bench rc=$?"; cut -c1-300 gpurun_out/r2m_bench.json; tail -3 gpurun_out/r2m_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench20.json 2> gpurun_out/r2m_bench20.err
stamp "bench20 rc=$?"; cut -c1-300 gpurun_out/r2m_bench20.json
timeout 100 python tools/step_timeline.py balance 32768 n_agents=4 > gpurun_out/r2m_timeline_balance.txt 2>&1
stamp "timeline"; cat gpurun_out/r2m_timeline_balance.txt | cut -c1-140
