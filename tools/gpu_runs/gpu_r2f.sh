#!/bin/bash
# round 2, call F: env-scheduling sweep (sort chunk x L2 fetch granularity), real-state 1Mi numbers, e2e timeline
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2f_stages.log; }
stamp start
timeout 120 python tools/e2e_timeline.py > gpurun_out/r2f_e2e_timeline.txt 2>&1
stamp "e2e timeline rc=$?"; tail -60 gpurun_out/r2f_e2e_timeline.txt
: > gpurun_out/r2f_sweep.txt
for l2 in 0 128; do for chunk in 256 512 1024 2048; do
  echo "## chunk=$chunk l2_fetch=$l2" >> gpurun_out/r2f_sweep.txt
  VMAS_B200_L2_FETCH=${l2/#0/} VMAS_B200_ENV_REORDER_CHUNK=$chunk KB_MAPPINGS=specialized_ordered timeout 100 python tools/kernel_bench.py balance flocking 1048576 >> gpurun_out/r2f_sweep.txt 2>&1
done; done
stamp "sweep done"; cat gpurun_out/r2f_sweep.txt
: > gpurun_out/r2f_real_state.txt
for chunk in 256 1024 2048; do for cfg in balance flocking; do
  VMAS_B200_ENV_REORDER_CHUNK=$chunk timeout 200 python bench.py --config $cfg --steps 60 --warmup 30 --no-cpu-baseline $( [ $cfg = flocking ] && echo "--total-envs 32768" ) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('chunk=$chunk', '$cfg', 'value %.3e kernel_us %.1f 1Mi: %s'%(d['value'], r['kernel_us'], json.dumps({k:r['at_1Mi_envs'].get(k) for k in ('kernel_us','kernel_us_identity_order','frac')})))" >> gpurun_out/r2f_real_state.txt
done; done
stamp "real state done"; cat gpurun_out/r2f_real_state.txt
