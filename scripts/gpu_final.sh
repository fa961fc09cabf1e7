#!/bin/bash
# end-of-round capture: gpu tests, bench (both arms), ncu launch list + --set full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "rel err|FAILED|passed|failed|image max|vae|clip|arcface|cond" gpurun_out/pytest_gpu.log | tail -24
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -1 gpurun_out/stages.log
REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_once.py > gpurun_out/ncu_unet.log 2>&1
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -s 186 -c 10 -o gpurun_out/prof_gemm python scripts/unet_once.py > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_flash -s 10 -c 2 -o gpurun_out/prof_flash python scripts/unet_once.py > gpurun_out/ncu_flash.log 2>&1
tail -1 gpurun_out/ncu_flash.log
