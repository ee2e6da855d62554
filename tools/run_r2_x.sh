#!/bin/bash
# round-2 GPU pass X (1 GPU): transfer kernels with L-vector-ordered gather / scatter sweeps: parity, then timing A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deterministic.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2x_pytest.log
for v in "" xfer8 xfer1; do
  if [ -n "$v" ]; then export CEED_B200_LIB=ceedpetscsolid_b200/variants/libceed_b200_$v.so; fi
  timeout 300 python tools/transfer_time.py 2>&1 | grep "ms" 
done | tee gpurun_out/r2x_transfer_ab3.txt
