#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --size 1024 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; tail -c 900 gpurun_out/bench_1024.json; tail -2 gpurun_out/bench_1024.err
timeout 900 python bench.py --ddim-steps 30 --scale 3.0 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_video.json 2> gpurun_out/bench_video.err; head -c 300 gpurun_out/bench_video.json; tail -2 gpurun_out/bench_video.err
