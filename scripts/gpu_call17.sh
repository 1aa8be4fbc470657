#!/bin/bash
mkdir -p gpurun_out
STEPS=20 scripts/run_variants.sh 'vg_*' 2>&1 | tee gpurun_out/r2_call17_c2.txt
STEPS=20 scripts/run_variants.sh 'vg_*' 2>&1 | tee -a gpurun_out/r2_call17_c2.txt
