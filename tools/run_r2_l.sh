#!/bin/bash
# round-2 GPU pass L (4 GPUs): final default lines at N = 2 and N = 4, new single-GPU test, deterministic-mode bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -x -q -k partitioned_apply 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --no-extras --no-e2e --deterministic 2>/dev/null | tail -1 > gpurun_out/r2l_det1.json
python -c "import json; d=json.load(open('gpurun_out/r2l_det1.json')); print('deterministic N=1', round(d['value'],2), round(d['ms_per_step'],4), d['config']['scatter'], d['gpu_launches'])"
for N in 2 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  ( time timeout 900 $TR --master-port 2951$N bench.py --gpus $N > gpurun_out/r2l_bench$N.json 2> gpurun_out/r2l_bench$N.err ) 2>&1 | grep real
  python - <<PY
import json
d=json.load(open('gpurun_out/r2l_bench$N.json'))
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])
for k in ('value_compressed_dm','strong_c4','snes_solve','e2e'): print(k, json.dumps(d.get(k))[:400])
PY
done
