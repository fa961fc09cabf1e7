#!/bin/bash
mkdir -p gpurun_out
WHAT=attn timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_flash4 -c 1 -o gpurun_out/s2_attn4 -f python scripts/ncu_ops.py > gpurun_out/ncu_a4.log 2>&1; tail -1 gpurun_out/ncu_a4.log
WHAT=lin timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -c 1 -o gpurun_out/s2_lin -f python scripts/ncu_ops.py > gpurun_out/ncu_lin.log 2>&1; tail -1 gpurun_out/ncu_lin.log
WHAT=conv timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_persist -c 1 -o gpurun_out/s2_conv -f python scripts/ncu_ops.py > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
WHAT=gn timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_fused -c 1 -o gpurun_out/s2_gn16 -f python scripts/ncu_ops.py > gpurun_out/ncu_gn16.log 2>&1; tail -1 gpurun_out/ncu_gn16.log
