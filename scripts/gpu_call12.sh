#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "row_strip or different_pitches or filtgrad_stage" > gpurun_out/r2_call12_tests.log 2>&1
tail -8 gpurun_out/r2_call12_tests.log
