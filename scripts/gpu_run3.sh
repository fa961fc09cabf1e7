#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "rel err|FAILED|passed|failed|^vae|^clip|^arcface|^cond|^c |^z_inpaint|^samples|^image" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python scripts/unet_once.py > gpurun_out/unet_once.log 2>&1; cat gpurun_out/unet_once.log
RFB_ATTN_FLASH=0 timeout 600 python scripts/unet_once.py > gpurun_out/unet_once_noflash.log 2>&1; tail -2 gpurun_out/unet_once_noflash.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
(time timeout 600 python bench.py --impl reference --steps 1 --warmup 1) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_once.py > gpurun_out/ncu_unet.log 2>&1
tail -3 gpurun_out/ncu_unet.log
