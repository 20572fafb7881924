#!/bin/bash
# round 2, call B: the warp-tile kernel (bit equality with the thread-per-env kernel, timings in both
# arithmetic builds), the reference's torch op chain on the B200, PCIe probe, ncu of the tile kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2b_stages.log; }
stamp start
timeout 300 python -m pytest tests/test_cabi_gpu.py -q -x -p no:cacheprovider > gpurun_out/r2b_cabi_tests.log 2>&1
stamp "cabi tests rc=$?"; tail -5 gpurun_out/r2b_cabi_tests.log
for arith in exact fast; do
  VMAS_B200_ARITH=$arith KB_MAPPINGS=specialized,tile timeout 200 python tools/kernel_bench.py balance transport navigation flocking > gpurun_out/r2b_kernel_bench_$arith.txt 2>&1
  stamp "kb $arith rc=$?"; cat gpurun_out/r2b_kernel_bench_$arith.txt
done
timeout 60 python tools/pcie_probe.py > gpurun_out/r2b_pcie.txt 2>&1; cat gpurun_out/r2b_pcie.txt
timeout 400 python bench.py --impl reference --ref-device cuda --steps 20 --warmup 5 > gpurun_out/r2b_ref_cuda.json 2> gpurun_out/r2b_ref_cuda.err
stamp "ref cuda rc=$?"; cat gpurun_out/r2b_ref_cuda.json; tail -5 gpurun_out/r2b_ref_cuda.err
VMAS_B200_SPEC_MAPPING=tile timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_tile.json 2> gpurun_out/r2b_bench_tile.err
stamp "bench tile rc=$?"; cat gpurun_out/r2b_bench_tile.json; tail -5 gpurun_out/r2b_bench_tile.err
KB_MAPPINGS=tile timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_tile -c 1 -f -o gpurun_out/r2b_balance_tile_1M python tools/kernel_bench.py balance 1048576 > gpurun_out/r2b_ncu.log 2>&1
stamp "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
