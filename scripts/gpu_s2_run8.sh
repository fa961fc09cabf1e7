#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "groupnorm or layernorm or attention" --timeout 300 --timeout-method=thread > gpurun_out/pytest_norm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_norm.log
tail -4 gpurun_out/pytest_norm.log
ONLY=gn timeout 600 python scripts/micro_bench.py 2>&1 | tail -22 | head -14
