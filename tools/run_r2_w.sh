#!/bin/bash
# round-2 GPU pass W (1 GPU): full GPU suite and the default bench line on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2w_pytest.log
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 3000 gpurun_out/r2w_bench.json
