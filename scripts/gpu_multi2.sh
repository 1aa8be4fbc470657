#!/bin/bash
# multi-GPU lease: wave-limited strip sweeps (ROO_TUNE_STRIP_CTAS_PER_SM) for the single-pair split
mkdir -p gpurun_out
for c in 0 2 3 5; do
  echo "== strip CTAs per SM: $c"
  timeout 600 python scripts/c5_split.py --reps 5 --strip-ctas $c --out gpurun_out/r2_c5_split_ctas$c.json 2>&1 | grep strips | cut -c1-200
done
