#!/bin/bash
# session 2, run 1: parity of the new kernels (attention v2, fused GroupNorm, vectorised LayerNorm) + A/B timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/pytest_ops.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_ops.log
tail -15 gpurun_out/pytest_ops.log
timeout 600 python scripts/micro_bench.py > gpurun_out/micro.log 2>&1; cat gpurun_out/micro.log | tail -40
timeout 600 python scripts/unet_ab.py "attn_flash=1" "gn_fused=0" "ln_vec=0" "attn_flash=1,gn_fused=0,ln_vec=0" > gpurun_out/unet_ab.log 2>&1; tail -8 gpurun_out/unet_ab.log
