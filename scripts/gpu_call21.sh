#!/bin/bash
mkdir -p gpurun_out
BENCH_ARGS="--workload c4_1920x1080x256_8path_subpix_lr" STEPS=10 scripts/run_variants.sh 'v8_*' 2>&1 | tee gpurun_out/r2_call21_c4.txt
BENCH_ARGS="--workload c5_3840x2160x256_8path_subpix_lr_single_gpu" STEPS=5 scripts/run_variants.sh 'v8_*' 2>&1 | tee gpurun_out/r2_call21_c5.txt
