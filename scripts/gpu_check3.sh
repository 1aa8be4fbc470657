#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "more_than_256 or invalid_arguments or row_strip or sgm_ieee or degenerate" 2>&1 | tail -12
