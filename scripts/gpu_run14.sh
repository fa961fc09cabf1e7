#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -3
timeout 300 python scripts/attn_bench.py 2>&1 | tail -3
timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_PAIR_MIN_NK=0 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_PAIR=0 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_PAIR_MIN_NK=24 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
