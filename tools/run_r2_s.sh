#!/bin/bash
# round-2 GPU pass S (1 GPU): rolled point loop (residual default, Jacobian as a variant): timing + parity
mkdir -p gpurun_out
( python tools/kernel_time.py; CEED_B200_LIB=ceedpetscsolid_b200/variants/libceed_b200_jacrolled.so python tools/kernel_time.py; python tools/kernel_time.py ) 2>&1 | grep jacobian | tee gpurun_out/r2s_rolled.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2s_pytest.log
