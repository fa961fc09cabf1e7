#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_debug.py > gpurun_out/dbg.log 2>&1; cut -c1-200 gpurun_out/dbg.log | tail -20
timeout 600 python scripts/gemm_sweep.py > gpurun_out/sweep.log 2>&1; grep -E "^---|auto|single persist|bn256|gen1" gpurun_out/sweep.log
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|rel err" gpurun_out/pytest_gpu.log | tail -10
timeout 600 python scripts/unet_once.py > gpurun_out/unet_once.log 2>&1; cat gpurun_out/unet_once.log
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -1 gpurun_out/stages.log
REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_once.py > gpurun_out/ncu_unet.log 2>&1
