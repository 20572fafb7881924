#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_4gpu_bench.json 2> gpurun_out/r2_4gpu_bench.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_4gpu_bench.err | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --config flocking --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_4gpu_bench_flocking.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_4gpu_bench", "r2_4gpu_bench_flocking"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f, "n_gpus", d.get("n_gpus"), "value %.3e ms %.4f e2e %.3e launches %s" % (d["value"], d["ms_per_step"], e.get("value", 0), d.get("gpu_launches")), d.get("per_rank_ms_per_step"), d.get("scaling"), d["config"].get("host_affinity"))
    except Exception as err:
        print(f, "failed", err)
PY
