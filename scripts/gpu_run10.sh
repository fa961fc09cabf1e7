#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_debug2.py > gpurun_out/dbg2.log 2>&1; cat gpurun_out/dbg2.log | cut -c1-330
timeout 600 python scripts/unet_once.py > gpurun_out/unet_once.log 2>&1; cat gpurun_out/unet_once.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -2
