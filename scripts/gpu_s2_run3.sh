#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes.log 2>&1; tail -70 gpurun_out/gemm_shapes.log
REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_unet_s2.csv python scripts/unet_once.py > gpurun_out/ncu_unet.log 2>&1
python scripts/agg_launches.py gpurun_out/launches_unet_s2.csv > gpurun_out/launches_unet_s2.txt; head -40 gpurun_out/launches_unet_s2.txt
