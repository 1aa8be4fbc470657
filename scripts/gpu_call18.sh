#!/bin/bash
mkdir -p gpurun_out/golden
python tests/golden/make_golden_bilateral.py gpurun_out/golden 2>&1 | tail -2
cp gpurun_out/golden/bilateral.npz tests/golden/
python -m pytest tests/test_oracle_golden.py -x -q -k "bilateral" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_frontback.py -x -q -k "bilateral" 2>&1 | tail -15
