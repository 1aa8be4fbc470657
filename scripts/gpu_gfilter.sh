#!/bin/bash
# guided filter: parity tests, timing, ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frontback.py tests/test_gpu_cpp_shim.py -q -x 2>&1 | tail -15 > gpurun_out/r2_gfilter_tests.log
cat gpurun_out/r2_gfilter_tests.log
timeout 300 python scripts/bench_gfilter.py 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'scan_|box_epi' -c 40 --csv --log-file gpurun_out/r2_gfilter_launches.csv python scripts/bench_gfilter.py > /dev/null 2>&1
tail -5 gpurun_out/r2_gfilter_launches.csv
