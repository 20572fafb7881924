#!/bin/bash
# round 2, call E: chunk-local env scheduling; run-time specialisation; bench with scheduling on.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2e_stages.log; }
stamp start
timeout 600 python -m pytest tests/test_cabi_gpu.py tests/test_env_gpu.py -q -p no:cacheprovider > gpurun_out/r2e_gpu_tests.log 2>&1
stamp "cabi+env tests rc=$?"; tail -8 gpurun_out/r2e_gpu_tests.log
for arith in exact fast; do
  VMAS_B200_ARITH=$arith KB_MAPPINGS=specialized,specialized_ordered timeout 200 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2e_kernel_bench_$arith.txt 2>&1
  stamp "kb $arith rc=$?"; cat gpurun_out/r2e_kernel_bench_$arith.txt
done
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2e_bench_exact.json 2> gpurun_out/r2e_bench_exact.err
stamp "bench rc=$?"; cat gpurun_out/r2e_bench_exact.json; tail -5 gpurun_out/r2e_bench_exact.err
VMAS_B200_ENV_REORDER_EVERY=0 timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2e_bench_exact_noorder.json 2> gpurun_out/r2e_bench_exact_noorder.err
stamp "bench no order rc=$?"; cat gpurun_out/r2e_bench_exact_noorder.json
KB_MAPPINGS=specialized_ordered timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_spec -s 1 -c 1 -f -o gpurun_out/r2e_balance_spec_ordered_1M_exact python tools/kernel_bench.py balance 1048576 > gpurun_out/r2e_ncu.log 2>&1
stamp "ncu rc=$?"; tail -2 gpurun_out/r2e_ncu.log
KB_MAPPINGS=specialized_ordered timeout 400 ncu --set full --import-source on --clock-control none -k regex:order_sort -c 1 -f -o gpurun_out/r2e_order_sort_1M python tools/kernel_bench.py balance 1048576 > gpurun_out/r2e_ncu2.log 2>&1
stamp "ncu sort rc=$?"
