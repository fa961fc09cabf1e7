#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gemm_sweep.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | tail -10
timeout 600 python scripts/unet_once.py > gpurun_out/unet_once.log 2>&1; cat gpurun_out/unet_once.log
