#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_call11_tests.log 2>&1
tail -6 gpurun_out/r2_call11_tests.log
STEPS=10 scripts/run_variants.sh 'vg_*' 2>&1 | tee gpurun_out/r2_call11_relacq.txt
