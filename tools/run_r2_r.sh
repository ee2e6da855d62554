#!/bin/bash
# round-2 GPU pass R (1 GPU): ncu --set full of the residual and diagonal kernels (hyperFS p=4, 64^3)
mkdir -p gpurun_out
CMD="python tools/kernel_time.py"
timeout 300 $CMD > gpurun_out/r2r_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_diag|k_fused_apply" -s 8 -c 40 -f -o gpurun_out/prof_r2r $CMD > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log; cat gpurun_out/r2r_plain.log | tail -1
