#!/bin/bash
# round-2 GPU pass M (1 GPU): last build -- kernel times (diagonal with the bulk prefetch), full GPU suite, default line
mkdir -p gpurun_out
python tools/kernel_time.py 2>&1 | grep jacobian | tee gpurun_out/r2m_kernel_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2m_pytest.log
timeout 600 python tools/kernel_table.py --configs hyperFS:4:64 2>/dev/null | grep diagonal
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['snes_solve']['time_s'])"
