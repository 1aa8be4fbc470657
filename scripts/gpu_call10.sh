#!/bin/bash
mkdir -p gpurun_out
STEPS=10 scripts/run_variants.sh 'vg_spin*' 2>&1 | tee gpurun_out/r2_call10_spin.txt
