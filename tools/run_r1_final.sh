#!/bin/bash
# final single-GPU pass of the round: tests, smoke, both bench arms, launch list under ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print('b200', d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['cpu_baseline'])"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --box 32 > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --box 32 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
