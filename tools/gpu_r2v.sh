#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/run_graph_steps.py balance 64 n_agents=4 2>&1 | grep -v "^frame" | tail -15
echo ==== INGEST_IN_KERNEL=0
VMAS_B200_INGEST_IN_KERNEL=0 timeout 300 python tools/run_graph_steps.py balance 64 n_agents=4 2>&1 | grep -v "^frame" | tail -12
echo ==== sanitizer
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_graph_steps.py balance 64 n_agents=4 2>&1 | grep -v "^frame" | grep -v "^=========     Host Frame\|^=========         in " | head -60 | tee gpurun_out/r2v_sanitizer.txt
