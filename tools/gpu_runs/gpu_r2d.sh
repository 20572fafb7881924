#!/bin/bash
# round 2, call D: env scheduling by contact signature; crafted worlds + stock-style + full-size env tests.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2d_stages.log; }
stamp start
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2d_gpu_tests.log 2>&1
stamp "gpu suite exact rc=$?"; tail -15 gpurun_out/r2d_gpu_tests.log
for arith in exact fast; do
  VMAS_B200_ARITH=$arith KB_MAPPINGS=specialized,specialized_ordered timeout 200 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2d_kernel_bench_$arith.txt 2>&1
  stamp "kb $arith rc=$?"; cat gpurun_out/r2d_kernel_bench_$arith.txt
done
for arith in exact fast; do
VMAS_B200_ARITH=$arith timeout 400 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/r2d_bench_$arith.json 2> gpurun_out/r2d_bench_$arith.err
stamp "bench $arith rc=$?"; cat gpurun_out/r2d_bench_$arith.json; tail -5 gpurun_out/r2d_bench_$arith.err
done
VMAS_B200_ARITH=exact KB_MAPPINGS=specialized_ordered timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_spec -s 1 -c 1 -f -o gpurun_out/r2d_balance_spec_ordered_1M_exact python tools/kernel_bench.py balance 1048576 > gpurun_out/r2d_ncu.log 2>&1
stamp "ncu rc=$?"; tail -2 gpurun_out/r2d_ncu.log
