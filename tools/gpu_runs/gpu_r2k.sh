#!/bin/bash
# round 2, call K: full GPU suite (both builds), kernel bench after the scheduling code left the default kernel, bench lines.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2k_stages.log; }
stamp start
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_gpu_tests.log 2>&1
stamp "gpu suite exact rc=$?"; tail -12 gpurun_out/r2k_gpu_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/r2k_smoke.log
for arith in exact fast; do
KB_MAPPINGS=specialized,tile VMAS_B200_ARITH=$arith timeout 200 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2k_kernel_bench_$arith.txt 2>&1
stamp "kb $arith rc=$?"; cat gpurun_out/r2k_kernel_bench_$arith.txt
done
VMAS_B200_ARITH=fast timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2k_gpu_tests_fast.log 2>&1
stamp "gpu suite fast rc=$?"; tail -8 gpurun_out/r2k_gpu_tests_fast.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_driver_like.json 2> gpurun_out/r2k_bench_driver_like.err
stamp "bench (driver-like) rc=$?"; cut -c1-400 gpurun_out/r2k_bench_driver_like.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_ref_driver_like.json 2> gpurun_out/r2k_ref_driver_like.err
stamp "reference arm rc=$?"; cut -c1-400 gpurun_out/r2k_ref_driver_like.json
timeout 400 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
stamp "bench rc=$?"; cut -c1-400 gpurun_out/r2k_bench.json
