#!/bin/bash
# round-2 GPU pass C (1 GPU): A/B of the fused-kernel tuning switches, new bench line, solver tests
mkdir -p gpurun_out
V=ceedpetscsolid_b200/variants
( python tools/kernel_time.py
  B200_OFFSETS_AHEAD=0 python tools/kernel_time.py
  B200_OFFSETS_AHEAD=2048 python tools/kernel_time.py
  for v in noscpad nobatch nofull oldjp r1like; do CEED_B200_LIB=$V/libceed_b200_$v.so python tools/kernel_time.py; done
  CEED_B200_LIB=$V/libceed_b200_r1like.so B200_OFFSETS_AHEAD=0 python tools/kernel_time.py
  python tools/kernel_time.py ) 2>&1 | grep -v "^$" | tee gpurun_out/r2c_variants.txt
timeout 600 python -m pytest tests/test_solver.py tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2c_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
for k in ('value','ms_per_step','gpu_launches','jcache_build_ms','parity','value_compressed_dm','strong_c4','snes_solve','e2e','cpu_baseline'):
    print(k, d.get(k))
print(d['roofline'])
PY
