#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k groupnorm --timeout 120 --timeout-method=thread 2>&1 | tail -3
ONLY=gn timeout 600 python scripts/micro_bench.py 2>&1 | grep groupnorm | head -13
