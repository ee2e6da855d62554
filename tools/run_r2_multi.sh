#!/bin/bash
# round-2 multi-GPU pass: bash tools/run_r2_multi.sh N [quick]     (through gpurun --gpus N)
# quick = only the default bench line the driver runs (no multi-GPU pytest, no exchange variants)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -12
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "[$N-masked or peer_memory_halo_matches_oracle[$N" 2>&1 | tail -4 | tee gpurun_out/r2m${N}_pytest.log
fi
# the default line the driver will run (parity, weak MatMult, compressed DM, strong C4, SNES solve)
( time timeout 900 $TR --master-port 29511 bench.py --gpus $N > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err ) 2>&1 | grep real
tail -4 gpurun_out/r2m${N}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2m${N}_bench.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
    for k in ('parity','value_compressed_dm','strong_c4','snes_solve','e2e'):
        print(k, d.get(k))
except Exception as e:
    print('bench line unreadable:', e); print(open('gpurun_out/r2m${N}_bench.json').read()[-2000:])
PY
# variants of the exchange, MatMult only (skipped in quick mode)
if [ -n "$2" ]; then exit 0; fi
VARIANTS=("--halo nccl" "--halo nccl --no-overlap" "--no-overlap" "")
if [ "$N" -ge 4 ]; then VARIANTS=("--halo nccl --no-overlap" "--no-overlap" ""); fi
for v in "${VARIANTS[@]}"; do
  tag=$(echo "$v" | tr -d ' -'); tag=${tag:-p2poverlap}
  timeout 300 $TR --master-port 29512 bench.py --gpus $N --no-extras --no-e2e --no-cpu $v 2>gpurun_out/r2m${N}_$tag.err | tail -1 > gpurun_out/r2m${N}_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2m${N}_$tag.json')); print('weak $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2m${N}_$tag.err
  timeout 300 $TR --master-port 29513 bench.py --gpus $N --no-extras --no-e2e --no-cpu --scaling strong --box 80 $v 2>gpurun_out/r2m${N}_s_$tag.err | tail -1 > gpurun_out/r2m${N}_s_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2m${N}_s_$tag.json')); print('strong80 $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2m${N}_s_$tag.err
done
