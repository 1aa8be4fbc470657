#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "bulk_copy or engine_ieee or c2_full or degenerate or random_shapes or golden_pipeline" > gpurun_out/r2_call5_tests.log 2>&1
tail -3 gpurun_out/r2_call5_tests.log
STEPS=10 scripts/run_variants.sh 'hs_*' 2>&1 | tee gpurun_out/r2_call5_hs.txt
