#!/bin/bash
# round 2, call 2: packed-f32x2 build -- parity (engine + sgm tests), bench c2 / c4, 256-disparity band-shape variants
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "sgm or engine or fused" > gpurun_out/r2_call2_tests.log 2>&1
tail -3 gpurun_out/r2_call2_tests.log
BENCH_ARGS="" scripts/run_variants.sh none > gpurun_out/r2_call2_c2.txt 2>&1
cat gpurun_out/r2_call2_c2.txt
BENCH_ARGS="--workload c4_1920x1080x256_8path_subpix_lr" STEPS=10 scripts/run_variants.sh 'v8_*' > gpurun_out/r2_call2_c4.txt 2>&1
cat gpurun_out/r2_call2_c4.txt
