#!/bin/bash
timeout 800 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x --timeout 600 --timeout-method=thread 2>&1 | tail -3
timeout 600 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes5.log 2>&1; grep -E " 65536 +(320|960|2560) +(320|1280) |16384 +640 +(640|2560) |total" gpurun_out/gemm_shapes5.log | head -20
REPS=10 timeout 600 python scripts/unet_ab.py 2>&1 | tail -2
