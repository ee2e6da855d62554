#!/bin/bash
# round-2 GPU pass K (1 GPU): what the driver runs at round end -- full GPU suite, smoke, default bench, reference arm --
# plus the final kernel table and the ncu launch list of the default bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2k_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
for k in ('value_compressed_dm','strong_c4','snes_solve','e2e','cpu_baseline','jcache_build_ms'): print(k, d.get(k))"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err ) 2>&1 | grep real
timeout 600 python tools/kernel_table.py > gpurun_out/r2k_kernel_table.md 2> gpurun_out/r2k_kernel_table.err; cat gpurun_out/r2k_kernel_table.md
CMD="python bench.py --no-cpu --no-extras --steps 20 --warmup 5"
timeout 300 $CMD > gpurun_out/r2k_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv $CMD > gpurun_out/r2k_ncu_list.log 2>&1
tail -1 gpurun_out/r2k_ncu_list.log | cut -c1-200
