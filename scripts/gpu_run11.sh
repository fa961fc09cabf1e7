#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -4
timeout 300 python scripts/gemm_debug2.py > gpurun_out/dbg2.log 2>&1; cat gpurun_out/dbg2.log | cut -c1-330
for km in 1 2; do RFB_GEMM_KMERGE=$km timeout 600 python scripts/unet_once.py 2>&1 | tail -2; done
RFB_GEMM_PAIR=0 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
