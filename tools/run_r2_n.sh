#!/bin/bash
# round-2 GPU pass N (1 GPU): diagonal kernel with the shared-memory (M, kappa) store -- parity + timing
mkdir -p gpurun_out
python tools/kernel_time.py 2>&1 | grep jacobian | tee gpurun_out/r2n_kernel_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2n_pytest.log
timeout 600 python tools/kernel_table.py 2>/dev/null | grep "diagonal" | tee gpurun_out/r2n_diag.txt
