#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
VMAS_B200_ARITH=fast timeout 900 python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider 2>&1 | tail -6
timeout 600 python -m pytest tests/test_env_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_env_kernel -s 2 -c 1 -o gpurun_out/r2final_step_env_kernel -f python tools/run_graph_steps.py balance 32768 n_agents=4 > gpurun_out/r2final_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2final_step_env_kernel.ncu-rep > gpurun_out/r2final_step_env_kernel_ncu_full.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2final_step_env_kernel.ncu-rep 0x1000 >> gpurun_out/r2final_step_env_kernel_ncu_full.txt 2>&1
head -60 gpurun_out/r2final_step_env_kernel_ncu_full.txt
