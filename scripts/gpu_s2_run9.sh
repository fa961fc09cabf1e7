#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -k "geglu or linear or ddim_update" --timeout 300 --timeout-method=thread 2>&1 | tail -3
timeout 600 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes3.log 2>&1; head -12 gpurun_out/gemm_shapes3.log; tail -1 gpurun_out/gemm_shapes3.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 52000 --csv --log-file gpurun_out/launches_bench_s2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/agg_launches.py gpurun_out/launches_bench_s2.csv > gpurun_out/launches_bench_s2.txt; head -45 gpurun_out/launches_bench_s2.txt
rm -f gpurun_out/launches_bench_s2.csv
