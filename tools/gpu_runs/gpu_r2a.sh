#!/bin/bash
# round 2, call A: the fast-arithmetic build on the GPU (suite + kernel timings), the reference's
# torch op chain on the B200 (J1), the new bench line, and a source-level ncu capture of the balance kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2a_stages.log; }
stamp start; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/r2a_stages.log
VMAS_B200_ARITH=fast timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_gpu_tests_fast.log 2>&1
stamp "fast suite rc=$?"; tail -30 gpurun_out/r2a_gpu_tests_fast.log
KB_MAPPINGS=specialized timeout 150 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2a_kernel_bench_exact.txt 2>&1
stamp "kb exact rc=$?"; cat gpurun_out/r2a_kernel_bench_exact.txt
VMAS_B200_ARITH=fast KB_MAPPINGS=specialized timeout 150 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2a_kernel_bench_fast.txt 2>&1
stamp "kb fast rc=$?"; cat gpurun_out/r2a_kernel_bench_fast.txt
timeout 400 python bench.py --impl reference --ref-device cuda --steps 20 --warmup 5 > gpurun_out/r2a_ref_cuda.json 2> gpurun_out/r2a_ref_cuda.err
stamp "ref cuda rc=$?"; cat gpurun_out/r2a_ref_cuda.json; tail -5 gpurun_out/r2a_ref_cuda.err
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
stamp "bench rc=$?"; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
KB_MAPPINGS=specialized timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_spec -c 1 -f -o gpurun_out/r2a_balance_1M python tools/kernel_bench.py balance 1048576 > gpurun_out/r2a_ncu.log 2>&1
stamp "ncu rc=$?"; tail -3 gpurun_out/r2a_ncu.log
