#!/bin/bash
# round-2 GPU pass P (1 GPU): the 64^3 solve in deterministic mode (same iteration counts expected), tube mesh solve
mkdir -p gpurun_out
timeout 600 python bench.py --solve --deterministic 2>/dev/null | tail -1 > gpurun_out/r2p_solve_det.json
python -c "import json; d=json.load(open('gpurun_out/r2p_solve_det.json')); print('deterministic solve', d['value'], d['snes_its'], d['ksp_its'], d['converged'], d['deterministic'])"
timeout 600 python bench.py --solve --mesh tube:4,32,24 --load-steps 4 2>/dev/null | tail -1 > gpurun_out/r2p_solve_tube.json
python -c "import json; d=json.load(open('gpurun_out/r2p_solve_tube.json')); print('tube solve', d['value'], d['snes_its'], d['ksp_its'], d['converged'])"
