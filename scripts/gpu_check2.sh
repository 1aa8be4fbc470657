#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "row_strip" 2>&1 | tail -3
