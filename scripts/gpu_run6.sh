#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "rel err|FAILED|passed|failed|image max" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python scripts/unet_once.py > gpurun_out/unet_once.log 2>&1; cat gpurun_out/unet_once.log
RFB_GEMM_PAIR=0 timeout 600 python scripts/unet_once.py > gpurun_out/unet_once_nopair.log 2>&1; tail -2 gpurun_out/unet_once_nopair.log
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -2 gpurun_out/stages.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_once.py > gpurun_out/ncu_unet.log 2>&1
tail -2 gpurun_out/ncu_unet.log
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 130 -c 6 -o gpurun_out/prof_gemm_pair python scripts/unet_once.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
