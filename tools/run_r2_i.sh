#!/bin/bash
# round-2 GPU pass I (2 GPUs): prefetch A/B (1 GPU), 2-GPU tests + bench legs with the forked interface kernel
mkdir -p gpurun_out
bash tools/run_r2_h.sh
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "[2-" 2>&1 | tail -3
for v in "--no-overlap" ""; do
  tag=$(echo "$v" | tr -d ' -'); tag=${tag:-p2poverlap}
  timeout 300 $TR --master-port 29512 bench.py --gpus $N --no-extras --no-e2e --no-cpu $v 2>gpurun_out/r2i_$tag.err | tail -1 > gpurun_out/r2i_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2i_$tag.json')); print('weak $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2i_$tag.err
  timeout 300 $TR --master-port 29513 bench.py --gpus $N --no-extras --no-e2e --no-cpu --scaling strong --box 40 $v 2>gpurun_out/r2i_s_$tag.err | tail -1 > gpurun_out/r2i_s_$tag.json
  python -c "import json; d=json.load(open('gpurun_out/r2i_s_$tag.json')); print('strong40 $tag', round(d['value'],2), round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))" || tail -3 gpurun_out/r2i_s_$tag.err
done
