#!/bin/bash
timeout 800 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "split_k or linear_tcgen05 or conv" --timeout 600 --timeout-method=thread 2>&1 | tail -3
REPS=10 timeout 600 python scripts/unet_ab.py "gemm_splitk=1" "gemm_splitk=0" "gemm_splitk=1" 2>&1 | tail -5
RFB_GEMM_SPLITK=1 timeout 600 python scripts/gemm_shapes.py 2>&1 | grep -E "^ +0 +1024 " | head
