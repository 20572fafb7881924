#!/bin/bash
# Times the specialised substep kernel for every library variant under tools/variants/ (built with
# different SPEC_BLOCK / SPEC_MIN_BLOCKS; same source, so same results) — one gpurun call.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 KB_MAPPINGS=specialized
for lib in tools/variants/lib_*.so; do
  echo "== $lib" | tee -a gpurun_out/variant_bench.txt
  VMAS_B200_LIB=$PWD/$lib timeout 100 python tools/kernel_bench.py balance flocking navigation transport 32768 1048576 2>&1 | tee -a gpurun_out/variant_bench.txt
done
