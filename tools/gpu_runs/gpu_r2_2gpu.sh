#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_2gpu_bench.json 2> gpurun_out/r2_2gpu_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2_2gpu_bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config flocking --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_2gpu_bench_flocking.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_2gpu_bench_reference.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_2gpu_bench", "r2_2gpu_bench_flocking", "r2_2gpu_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f, "n_gpus", d.get("n_gpus"), "value %.3e ms %.4f e2e %.3e launches %s" % (d["value"], d["ms_per_step"], e.get("value", 0), d.get("gpu_launches")), d.get("per_rank_ms_per_step"), d.get("scaling"))
    except Exception as err:
        print(f, "failed", err)
PY
