#!/bin/bash
# round-2 GPU pass R (1 GPU): ncu --set full of ONE residual and ONE diagonal launch (hyperFS p=4, 64^3)
# (a 40-launch capture with source exceeded the 64 MiB gpurun_out limit and was lost: one launch per kernel)
mkdir -p gpurun_out
CMD="python tools/kernel_time.py"
timeout 300 $CMD > gpurun_out/r2r_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_apply -s 10 -c 1 -f -o gpurun_out/prof_r2r_residual $CMD > gpurun_out/r2r_ncu1.log 2>&1
tail -1 gpurun_out/r2r_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_diag -s 3 -c 1 -f -o gpurun_out/prof_r2r_diag $CMD > gpurun_out/r2r_ncu2.log 2>&1
tail -1 gpurun_out/r2r_ncu2.log; ls -la gpurun_out/prof_r2r*
