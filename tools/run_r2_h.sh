#!/bin/bash
# round-2 GPU pass H (1 GPU): prefetch-distance A/B of the fused kernel (run-time switches)
mkdir -p gpurun_out
( python tools/kernel_time.py
  for x in 185 370 740; do B200_X_AHEAD=$x python tools/kernel_time.py; done
  for s in 148 370 740; do B200_SLAB_AHEAD=$s python tools/kernel_time.py; done
  B200_X_AHEAD=370 B200_SLAB_AHEAD=370 python tools/kernel_time.py
  B200_X_AHEAD=740 B200_SLAB_AHEAD=740 python tools/kernel_time.py
  python tools/kernel_time.py ) 2>&1 | grep jacobian | tee gpurun_out/r2h_prefetch.txt
