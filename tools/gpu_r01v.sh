#!/bin/bash
# r01v: ncu --set full of the fused decoder tail with the warp-MMA head (source-level stall reasons)
AESR_HEAD_MMA=1 timeout 120 python tools/head_one.py > gpurun_out/head_one_r01v.log 2>&1; echo "plain rc $?"
AESR_HEAD_MMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel" -s 2 -c 1 -o gpurun_out/prof_head_mma_r01v python tools/head_one.py > gpurun_out/ncu_head_r01v.log 2>&1; echo "ncu rc $?"
