#!/bin/bash
# round-2 GPU pass V (1 GPU): reference host code incl. ViewDiagnosticQuantities + borrowed-host write-through
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_host_code_on_gpu.py tests/test_gpu_edge.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2v_pytest.log
