#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or in_sweep or 256 or c4_full or random_shapes or degenerate" > gpurun_out/r2_call16_tests.log 2>&1
tail -4 gpurun_out/r2_call16_tests.log
STEPS=10 scripts/run_variants.sh 'vg_*' 2>&1 | tee gpurun_out/r2_call16_c2.txt
