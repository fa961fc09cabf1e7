#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2a.json 2> gpurun_out/bench_s2a.err; tail -c 2000 gpurun_out/bench_s2a.json; tail -3 gpurun_out/bench_s2a.err
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -1 gpurun_out/stages.log
