#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 80 python -m pytest tests/test_cabi_gpu.py -q -x -p no:cacheprovider -k cooperative > gpurun_out/coop_tests.log 2>&1
echo "coop tests rc=$?"; tail -6 gpurun_out/coop_tests.log
KB_MAPPINGS=specialized,cooperative timeout 60 python tools/kernel_bench.py balance transport navigation flocking 32768 131072 1048576 > gpurun_out/coop_kernel_bench.txt 2>&1
cat gpurun_out/coop_kernel_bench.txt
