#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "in_sweep or fused or engine_ieee" > gpurun_out/r2_call8_tests.log 2>&1
tail -3 gpurun_out/r2_call8_tests.log
STEPS=10 scripts/run_variants.sh 'vg_*' 2>&1 | tee gpurun_out/r2_call8_c2.txt
BENCH_ARGS="--workload c4_1920x1080x256_8path_subpix_lr" STEPS=10 scripts/run_variants.sh 'vg_fc16_8fc8' 2>&1 | tee gpurun_out/r2_call8_c4.txt
