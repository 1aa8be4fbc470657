#!/bin/bash
# guided filter: parity tests incl. the chunk loop, sanitizers on the small script
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontback.py -q -x 2>&1 | tail -5 > gpurun_out/r2_gfilter_tests.log
cat gpurun_out/r2_gfilter_tests.log
for tool in memcheck initcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2_san_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_san_$tool.log | tail -1)"
done
