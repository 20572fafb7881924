#!/bin/bash
# round 2, call C: tile kernel v2 (contact ring) — bit equality, timings in both arithmetic builds,
# parity report of both builds against the reference goldens, ncu of the tile kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2c_stages.log; }
stamp start
timeout 300 python -m pytest tests/test_cabi_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2c_cabi_tests.log 2>&1
stamp "cabi tests exact rc=$?"; tail -5 gpurun_out/r2c_cabi_tests.log
VMAS_B200_ARITH=fast VMAS_B200_SPEC_MAPPING=tile timeout 300 python -m pytest tests/test_cabi_gpu.py tests/test_env_gpu.py -q -p no:cacheprovider > gpurun_out/r2c_tests_fast_tile.log 2>&1
stamp "cabi+env tests fast/tile rc=$?"; tail -5 gpurun_out/r2c_tests_fast_tile.log
for arith in exact fast; do
  VMAS_B200_ARITH=$arith KB_MAPPINGS=specialized,tile timeout 200 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2c_kernel_bench_$arith.txt 2>&1
  stamp "kb $arith rc=$?"; cat gpurun_out/r2c_kernel_bench_$arith.txt
  VMAS_B200_ARITH=$arith timeout 200 python tools/gpu_parity_report.py > gpurun_out/r2c_parity_$arith.txt 2>&1
  stamp "parity $arith rc=$?"; tail -25 gpurun_out/r2c_parity_$arith.txt
done
for arith in exact fast; do
VMAS_B200_ARITH=$arith KB_MAPPINGS=tile timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_tile -c 1 -f -o gpurun_out/r2c_balance_tile_1M_$arith python tools/kernel_bench.py balance 1048576 > gpurun_out/r2c_ncu_$arith.log 2>&1
stamp "ncu $arith rc=$?"; tail -2 gpurun_out/r2c_ncu_$arith.log
done
VMAS_B200_ARITH=fast KB_MAPPINGS=specialized timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_spec -c 1 -f -o gpurun_out/r2c_balance_spec_1M_fast python tools/kernel_bench.py balance 1048576 > gpurun_out/r2c_ncu_spec_fast.log 2>&1
stamp "ncu spec fast rc=$?"
