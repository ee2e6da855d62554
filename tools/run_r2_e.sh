#!/bin/bash
# round-2 GPU pass E (8 GPUs): multi-GPU tests, the driver's default line, exchange variants, weak + strong
mkdir -p gpurun_out
( python tools/kernel_time.py; CEED_B200_LIB=ceedpetscsolid_b200/variants/libceed_b200_noalloc.so python tools/kernel_time.py ) 2>&1 | grep jacobian | tee gpurun_out/r2e_kernel_time.txt
bash tools/run_r2_multi.sh 8 2>&1 | tee gpurun_out/r2e_multi.log
