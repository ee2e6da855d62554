#!/bin/bash
# round-2 GPU pass Q (1 GPU): the reference's own host code (setuplibceed.c, matops.c, misc.c) driving /gpu/b200
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_host_code_on_gpu.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2q_pytest.log
