#!/bin/bash
# single-GPU regression + measurement pass (run on the GPU box through gpurun)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for dm in compressed masked; do
  timeout 300 python bench.py --no-cpu --dm $dm 2>/dev/null | tail -1 > gpurun_out/bench_$dm.json
  python -c "import json; d=json.load(open('gpurun_out/bench_$dm.json')); print('$dm', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])"
done
timeout 300 python bench.py --solve --box 64 --load-steps 10 > gpurun_out/solve64_masked.json 2> gpurun_out/solve64_masked.err
python -c "import json; d=json.load(open('gpurun_out/solve64_masked.json')); print('solve', d['value'], d['snes_its'], d['ksp_its'], d['coarse_pcg_its'])"
