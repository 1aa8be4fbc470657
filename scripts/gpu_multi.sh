#!/bin/bash
# multi-GPU lease: row-strip split of one pair (real peer hand-off), in-library pair sharding, c5 / c3 numbers
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "row_strip or multi_gpu" > gpurun_out/r2_multi_tests_${1:-n}.log 2>&1
tail -5 gpurun_out/r2_multi_tests_${1:-n}.log
timeout 900 python scripts/c5_split.py --reps 5 --out gpurun_out/r2_c5_split_${1:-n}.json 2>&1 | tail -6
timeout 900 python scripts/multi_engine_c3.py --out gpurun_out/r2_multi_engine_${1:-n}.json 2>&1 | tail -5
