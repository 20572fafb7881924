#!/bin/bash
# round 2, call J: occupancy variants (registers per thread) x env scheduling, tiled-golden and real states;
# the reference's torch op chain on the GPU for the other BASELINE configs.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2j_stages.log; }
stamp start
: > gpurun_out/r2j_variants.txt
for lib in default tools/variants/lib_b64m10.so tools/variants/lib_b64m12.so tools/variants/lib_b128m4.so tools/variants/lib_b128m6.so; do
  echo "== $lib" >> gpurun_out/r2j_variants.txt
  if [ $lib = default ]; then unset VMAS_B200_LIB; else export VMAS_B200_LIB=$PWD/$lib; fi
  KB_MAPPINGS=specialized,specialized_ordered timeout 150 python tools/kernel_bench.py balance flocking 1048576 >> gpurun_out/r2j_variants.txt 2>&1
  VMAS_B200_ENV_REORDER_EVERY=8 timeout 200 python bench.py --steps 60 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('real state balance: value %.3e kernel_us %.1f 1Mi: %s'%(d['value'], r['kernel_us'], json.dumps({k:r['at_1Mi_envs'].get(k) for k in ('kernel_us','kernel_us_identity_order','frac')})))" >> gpurun_out/r2j_variants.txt
done
unset VMAS_B200_LIB
stamp "variants done"; cat gpurun_out/r2j_variants.txt
for cfg in transport3 navigation; do
timeout 300 python bench.py --impl reference --ref-device cuda --config $cfg --steps 10 --warmup 3 > gpurun_out/r2j_ref_cuda_$cfg.json 2> gpurun_out/r2j_ref_cuda_$cfg.err
stamp "ref cuda $cfg rc=$?"; cut -c1-260 gpurun_out/r2j_ref_cuda_$cfg.json
done
timeout 300 python bench.py --impl reference --ref-device cuda --config flocking --total-envs 32768 --steps 10 --warmup 3 > gpurun_out/r2j_ref_cuda_flocking.json 2> gpurun_out/r2j_ref_cuda_flocking.err
stamp "ref cuda flocking rc=$?"; cut -c1-260 gpurun_out/r2j_ref_cuda_flocking.json
