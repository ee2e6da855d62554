#!/bin/bash
# round-2 GPU pass B (1 GPU): GPU test-suite with the new kernels / deterministic mode, kernel table, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/r2b_pytest.log
timeout 600 python tools/kernel_table.py > gpurun_out/r2b_kernel_table.md 2> gpurun_out/r2b_kernel_table.err; cat gpurun_out/r2b_kernel_table.md; tail -3 gpurun_out/r2b_kernel_table.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'])"
