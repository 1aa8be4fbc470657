#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_solo.py > gpurun_out/r2_san_solo_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2_san_solo_$tool.log | tail -1) $(grep -c done gpurun_out/r2_san_solo_$tool.log)"
done
