#!/bin/bash
# round 2, call H: fixed tests, smoke, A/B of the forked observation stream.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2h_stages.log; }
stamp start
timeout 600 python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider > gpurun_out/r2h_env_tests.log 2>&1
stamp "env tests rc=$?"; tail -12 gpurun_out/r2h_env_tests.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/r2h_smoke.log
: > gpurun_out/r2h_ab.txt
for rep in 1 2; do for fork in 1 0; do
  VMAS_B200_FORK_OBS=$fork timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fork=$fork', 'balance value %.3e ms %.4f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))" >> gpurun_out/r2h_ab.txt
done; done
for fork in 1 0; do
  VMAS_B200_FORK_OBS=$fork timeout 300 python bench.py --config navigation --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fork=$fork', 'navigation value %.3e ms %.4f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))" >> gpurun_out/r2h_ab.txt
  VMAS_B200_FORK_OBS=$fork timeout 300 python bench.py --config flocking --total-envs 32768 --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fork=$fork', 'flocking32768 value %.3e ms %.4f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))" >> gpurun_out/r2h_ab.txt
done
stamp "A/B done"; cat gpurun_out/r2h_ab.txt
