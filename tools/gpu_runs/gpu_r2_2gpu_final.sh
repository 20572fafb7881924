#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ag_2gpu_bench.json 2> gpurun_out/r2ag_2gpu_bench.err
echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ag_2gpu_bench.json").read().strip().splitlines()[-1])
print("n_gpus", d.get("n_gpus"), "value %.3e ms %.4f e2e %.3e launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")), d.get("per_rank_ms_per_step"), d["config"].get("bracket_us"), d["config"].get("remeasured"))
PY
