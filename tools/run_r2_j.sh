#!/bin/bash
# round-2 GPU pass J (8 GPUs): final default line with the forked interface kernel + strong/weak MatMult-only legs
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "peer_memory_halo_matches_oracle[8" 2>&1 | tail -2
( time timeout 900 $TR --master-port 29511 bench.py --gpus $N > gpurun_out/r2j_bench8.json 2> gpurun_out/r2j_bench8.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench8.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
for k in ('parity','value_compressed_dm','strong_c4','snes_solve','e2e'): print(k, d.get(k))
PY
for v in "--no-overlap" ""; do
  tag=$(echo "$v" | tr -d ' -'); tag=${tag:-p2poverlap}
  timeout 300 $TR --master-port 29512 bench.py --gpus $N --no-extras --no-e2e --no-cpu $v 2>gpurun_out/r2j_$tag.err | tail -1 > gpurun_out/r2j_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2j_$tag.json')); print('weak $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2j_$tag.err
  timeout 300 $TR --master-port 29513 bench.py --gpus $N --no-extras --no-e2e --no-cpu --scaling strong --box 80 $v 2>gpurun_out/r2j_s_$tag.err | tail -1 > gpurun_out/r2j_s_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2j_s_$tag.json')); print('strong80 $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2j_s_$tag.err
done
