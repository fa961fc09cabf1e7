#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x --timeout 600 --timeout-method=thread 2>&1 | tail -3
timeout 600 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes4.log 2>&1; grep -E " 320 +(320|1280) |  960 +320|2560 +320|total" gpurun_out/gemm_shapes4.log | head -20
WHAT=lin timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -c 1 -o gpurun_out/s2_lin2 -f python scripts/ncu_ops.py > gpurun_out/ncu_lin2.log 2>&1; tail -1 gpurun_out/ncu_lin2.log
