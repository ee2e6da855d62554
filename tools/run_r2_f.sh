#!/bin/bash
# round-2 GPU pass F (1 GPU): occupancy variants of residual / Jacobian / diagonal, final kernel table, default bench line
mkdir -p gpurun_out
V=ceedpetscsolid_b200/variants
( python tools/kernel_time.py
  for v in jac4 res4a res3a res3 diag2 diag4; do CEED_B200_LIB=$V/libceed_b200_$v.so python tools/kernel_time.py; done
  python tools/kernel_time.py ) 2>&1 | grep jacobian | tee gpurun_out/r2f_variants.txt
timeout 600 python tools/kernel_table.py > gpurun_out/r2f_kernel_table.md 2> gpurun_out/r2f_kernel_table.err; cat gpurun_out/r2f_kernel_table.md; tail -3 gpurun_out/r2f_kernel_table.err
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r2f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])
for k in ('value_compressed_dm','strong_c4','snes_solve','e2e','cpu_baseline','jcache_build_ms','parity'): print(k, d.get(k))"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err ) 2>&1 | grep real
cat gpurun_out/r2f_bench_ref.json | cut -c1-700
