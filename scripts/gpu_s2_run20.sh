#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -3
for i in 1 2; do timeout 600 python scripts/stage_times.py 2>&1 | tail -1; done
