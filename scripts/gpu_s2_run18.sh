#!/bin/bash
REPS=10 timeout 600 python scripts/unet_ab.py "conv_tma_stride2=1" "conv_tma_stride2=0" "conv_tma_stride2=1" "conv_tma_stride2=0" 2>&1 | tail -6
for v in 1 0; do RFB_CONV_TMA_STRIDE2=$v S=4 timeout 600 python scripts/stage_times.py 2>&1 | tail -1; done
