#!/bin/bash
# round-2 GPU pass O (1 GPU): where the solve spends its time (per-Newton-step set-up pieces, V-cycle, shares)
mkdir -p gpurun_out
python tools/setup_bench.py 64 2>&1 | tail -20 | tee gpurun_out/r2o_setup.txt
python tools/solve_profile.py 64 2 2>&1 | tail -24 | tee gpurun_out/r2o_solve_profile.txt
