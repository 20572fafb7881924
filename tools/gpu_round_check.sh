#!/bin/bash
# One gpurun call at the end of a session: the reset kernels against their oracle, the reset cost,
# the bench line, smoke, then as much of the full GPU suite as the time allows.  Every stage writes
# its own log under gpurun_out/ so a cut-off call still leaves the earlier results.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/stages.log; }
stamp "start"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/stages.log
timeout 200 python -m pytest tests/test_reset_gpu.py -q -x -p no:cacheprovider > gpurun_out/reset_tests.log 2>&1
stamp "reset tests rc=$?"; tail -5 gpurun_out/reset_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 120 python tools/reset_bench.py > gpurun_out/reset_bench.jsonl 2> gpurun_out/reset_bench.err
stamp "reset bench rc=$?"; cat gpurun_out/reset_bench.jsonl
timeout 200 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
stamp "bench rc=$?"; cat gpurun_out/bench.json
timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_reset_gpu.py > gpurun_out/gpu_tests.log 2>&1
stamp "gpu suite rc=$?"; tail -5 gpurun_out/gpu_tests.log
