#!/bin/bash
mkdir -p gpurun_out/golden
python tests/golden/make_golden_n4b.py gpurun_out/golden 2>&1 | tail -2
cp gpurun_out/golden/sqpen.npz gpurun_out/golden/filtgrad.npz tests/golden/
python -m pytest tests/test_oracle_golden.py -x -q -k "filter_disp or square_penalty" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_frontback.py tests/test_gpu_cpp_shim.py -x -q > gpurun_out/r2_call9_tests.log 2>&1
tail -5 gpurun_out/r2_call9_tests.log
STEPS=10 scripts/run_variants.sh 'vg_spin*' 2>&1 | tee gpurun_out/r2_call9_spin.txt
