#!/bin/bash
# multi-GPU regression + measurement pass: bash tools/run_r1_multi.sh N [box] [load-steps] [quick]   (through gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "[$N-" 2>&1 | tail -3
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale${N}_masked.json
$TR --master-port 29512 bench.py --gpus $N --no-e2e --overlap 2>/dev/null | tail -1 > gpurun_out/scale${N}_masked_overlap.json
VARIANTS="masked masked_overlap"
if [ -z "$4" ]; then
  $TR --master-port 29513 bench.py --gpus $N --no-e2e --dm compressed 2>/dev/null | tail -1 > gpurun_out/scale${N}_compressed.json
  VARIANTS="$VARIANTS compressed"
fi
for f in $VARIANTS; do python -c "import json; d=json.load(open('gpurun_out/scale${N}_$f.json')); print('$f', round(d['value'],2), round(d['ms_per_step'],4), d['clocks'])"; done
$TR --master-port 29514 bench.py --gpus $N --solve --box ${2:-32} --load-steps ${3:-4} 2>gpurun_out/solve_multi.err | tail -1 > gpurun_out/solve${N}.json
python -c "import json; d=json.load(open('gpurun_out/solve${N}.json')); print('solve', d['value'], d['snes_its'], d['ksp_its'], d['coarse_pcg_its'], d['dofs_unconstrained'])"
tail -3 gpurun_out/solve_multi.err
