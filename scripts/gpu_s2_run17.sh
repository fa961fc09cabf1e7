#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "conv_generic or groupnorm" --timeout 120 --timeout-method=thread 2>&1 | tail -12
