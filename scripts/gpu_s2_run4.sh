#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/pytest_ops.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_ops.log
tail -5 gpurun_out/pytest_ops.log
timeout 600 python scripts/unet_ab.py "gemm_wave_bn=0" > gpurun_out/unet_ab.log 2>&1; tail -4 gpurun_out/unet_ab.log
timeout 600 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes2.log 2>&1; head -45 gpurun_out/gemm_shapes2.log
