#!/bin/bash
# ncu --set full of the five per-chunk launches of the guided filter (after the three guide-image launches)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_|box_epi' -s 3 -c 5 -f -o /tmp/r2_gfilter python scripts/bench_gfilter.py > gpurun_out/r2_gfilter_ncu.log 2>&1
ncu -i /tmp/r2_gfilter.ncu-rep --page raw --csv > gpurun_out/r2_gfilter_raw.csv 2>/dev/null
ls -la /tmp/r2_gfilter.ncu-rep gpurun_out/r2_gfilter_raw.csv
