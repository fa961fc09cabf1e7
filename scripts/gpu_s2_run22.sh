#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40300 -c 11500 --csv --log-file gpurun_out/launches_bench_tail.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_tail.log 2>&1
python scripts/agg_launches.py gpurun_out/launches_bench_tail.csv > gpurun_out/launches_bench_tail.txt; head -32 gpurun_out/launches_bench_tail.txt
rm -f gpurun_out/launches_bench_tail.csv
