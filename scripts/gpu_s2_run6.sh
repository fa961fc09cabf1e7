#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k attention --timeout 300 --timeout-method=thread > gpurun_out/pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_attn.log
tail -12 gpurun_out/pytest_attn.log
ONLY=attn timeout 600 python scripts/micro_bench.py > gpurun_out/micro_attn.log 2>&1; tail -20 gpurun_out/micro_attn.log
