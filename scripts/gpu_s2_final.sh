#!/bin/bash
# end-of-session capture: all GPU tests, smoke, bench (both arms), per-stage times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2600 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 700 gpurun_out/bench_ref_final.json
timeout 600 python scripts/stage_times.py > gpurun_out/stages.log 2>&1; tail -1 gpurun_out/stages.log
