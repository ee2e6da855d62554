#!/bin/bash
# round-2 GPU pass T (1 GPU): final build (diagonal with compile-time strides, residual unrolled): timing + full suite
mkdir -p gpurun_out
( python tools/kernel_time.py; python tools/kernel_time.py ) 2>&1 | grep jacobian | tee gpurun_out/r2t_kernel_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2t_pytest.log
timeout 600 python tools/kernel_table.py --configs hyperSS:3:64,hyperFS:4:64 2>/dev/null | grep "diagonal\|residual" | tee gpurun_out/r2t_table.txt
