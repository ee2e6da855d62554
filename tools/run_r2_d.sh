#!/bin/bash
# round-2 GPU pass D (2 GPUs): final kernel timing, full GPU suite (incl. 2-GPU tests), 2-GPU bench legs, ncu capture
mkdir -p gpurun_out
python tools/kernel_time.py 2>&1 | tail -1 | tee gpurun_out/r2d_kernel_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2d_pytest.log
bash tools/run_r2_multi.sh 2 quick 2>&1 | tee gpurun_out/r2d_multi.log
CMD="python bench.py --no-cpu --no-e2e --no-extras --steps 3 --warmup 3"
timeout 300 $CMD > gpurun_out/r2d_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_apply -s 4 -c 2 -f -o gpurun_out/prof_r2d $CMD > gpurun_out/r2d_ncu_full.log 2>&1
tail -2 gpurun_out/r2d_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/r2d_ncu_list.log 2>&1
tail -1 gpurun_out/r2d_ncu_list.log | cut -c1-300
