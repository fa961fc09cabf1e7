#!/bin/bash
# full check: all GPU tests, smoke, bench (ours), per-stage times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2b.json 2> gpurun_out/bench_s2b.err; tail -c 1500 gpurun_out/bench_s2b.json; tail -3 gpurun_out/bench_s2b.err
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -1 gpurun_out/stages.log
