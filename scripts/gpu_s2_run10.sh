#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_unet_gpu.py -m gpu -q -x --timeout 600 --timeout-method=thread 2>&1 | tail -4
for v in 1 0 1 0; do RFB_CFG_SHARE=$v timeout 600 python scripts/stage_times.py 2>&1 | tail -1; done
