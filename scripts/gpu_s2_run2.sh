#!/bin/bash
mkdir -p gpurun_out
WHAT=attn timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_flash2 -c 1 -o gpurun_out/s2_attn2 -f python scripts/ncu_ops.py > gpurun_out/ncu_attn2.log 2>&1; tail -2 gpurun_out/ncu_attn2.log
WHAT=gn timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_fused -c 4 -o gpurun_out/s2_gnf -f python scripts/ncu_ops.py > gpurun_out/ncu_gnf.log 2>&1; tail -2 gpurun_out/ncu_gnf.log
WHAT=gn RFB_GN_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ -c 6 -o gpurun_out/s2_gns -f python scripts/ncu_ops.py > gpurun_out/ncu_gns.log 2>&1; tail -2 gpurun_out/ncu_gns.log
