#!/bin/bash
ROO_STRESS_SEEDS=80 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "split_engine_random or random_shapes" 2>&1 | tail -8
