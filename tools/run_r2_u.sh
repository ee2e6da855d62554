#!/bin/bash
# round-2 GPU pass U (2 GPUs): final build, the driver's default line at N = 2 (insurance after the last kernel changes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29511 bench.py --gpus 2 > gpurun_out/r2u_bench2.json 2> gpurun_out/r2u_bench2.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/r2u_bench2.json'))
print(d['value'], d['ms_per_step'], d['parity']['ok'], d['strong_c4']['value'], d['snes_solve']['time_s'], d['snes_solve']['snes_its'], d['snes_solve']['ksp_its'])" || tail -5 gpurun_out/r2u_bench2.err
