#!/bin/bash
# round 2, call I (8 GPUs): weak scaling of the default config at N = 1, 8 (NUMA pinning on), strong scaling of flocking
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/r2i_stages.log; }
stamp start; nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
timeout 200 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2i_n1.json 2> gpurun_out/r2i_n1.err
stamp "N=1 rc=$?"
for n in 2 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/r2i_n$n.json 2> gpurun_out/r2i_n$n.err
stamp "N=$n rc=$?"; tail -3 gpurun_out/r2i_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --config flocking --steps 50 --warmup 10 > gpurun_out/r2i_flocking_n8.json 2> gpurun_out/r2i_flocking_n8.err
stamp "flocking N=8 rc=$?"
for f in r2i_n1 r2i_n2 r2i_n8 r2i_flocking_n8; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
    print('$f', 'value %.3e ms %.4f e2e %.3e'%(d['value'],d['ms_per_step'],d['e2e']['value']), 'per-rank ms', [round(x,4) for x in d['per_rank_ms_per_step']], d['config'].get('host_affinity'))
except Exception as e:
    print('$f', 'ERR', e)
PY
done | tee gpurun_out/r2i_summary.txt
