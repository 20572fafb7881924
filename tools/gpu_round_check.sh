#!/bin/bash
# One gpurun call at the end of a session: the full GPU suite (parity of every kernel against its
# oracle), smoke, the substep kernel alone, the reset cost and the bench line.  Every stage writes
# its own log under gpurun_out/ so a cut-off call still leaves the earlier results.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/stages.log; }
stamp "start"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/stages.log
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
stamp "gpu suite rc=$?"; tail -15 gpurun_out/gpu_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/smoke.log
KB_MAPPINGS=specialized timeout 100 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/kernel_bench.txt 2>&1
stamp "kernel bench rc=$?"; cat gpurun_out/kernel_bench.txt
timeout 200 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
stamp "bench rc=$?"; cat gpurun_out/bench.json
