#!/bin/bash
# round-2 first GPU pass (1 GPU): full GPU test-suite incl. C2/C3 oracle parity, bench line, kernel table,
# ncu launch list + one full capture of the SHIPPED fused Jacobian kernel at 64^3.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
timeout 600 python tools/kernel_table.py > gpurun_out/r2a_kernel_table.md 2> gpurun_out/r2a_kernel_table.err; cat gpurun_out/r2a_kernel_table.md
CMD="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3"
timeout 300 $CMD > gpurun_out/r2a_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_apply -s 4 -c 2 -f -o gpurun_out/prof_r2a $CMD > gpurun_out/r2a_ncu_full.log 2>&1
tail -3 gpurun_out/r2a_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv $CMD > gpurun_out/r2a_ncu_list.log 2>&1
tail -2 gpurun_out/r2a_ncu_list.log
ls -la gpurun_out | grep r2a
