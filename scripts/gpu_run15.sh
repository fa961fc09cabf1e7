#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -3
timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_EPI3_MAX_NK=0 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_EPI3_MAX_NK=24 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
RFB_GEMM_EPI3_MAX_NK=5 timeout 600 python scripts/unet_once.py 2>&1 | tail -2
