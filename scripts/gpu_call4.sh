#!/bin/bash
# round 2, call 4: ncu --set full of one c2 step (aggregation kernels) with the hsweep + f32x2 build
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"sgm_" -s 12 -c 4 -f -o gpurun_out/r2_c2_hs python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c2_hs_ncu.log 2>&1
ls -la gpurun_out
ncu -i gpurun_out/r2_c2_hs.ncu-rep --page raw --csv > gpurun_out/r2_c2_hs_raw.csv
