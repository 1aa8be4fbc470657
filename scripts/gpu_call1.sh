#!/bin/bash
# round 2, call 1: new full-size parity tests, baseline bench lines of the round-1 build, ncu of the 256-disparity kernels
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "live_reference or c4_full or c5_full" > gpurun_out/r2_call1_tests.log 2>&1
tail -3 gpurun_out/r2_call1_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_base_c2.json 2> gpurun_out/r2_base_c2.err
python bench.py --workload c4_1920x1080x256_8path_subpix_lr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_base_c4.json 2>> gpurun_out/r2_base_c2.err
python bench.py --workload c5_3840x2160x256_8path_subpix_lr_single_gpu --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_base_c5.json 2>> gpurun_out/r2_base_c2.err
python bench.py --workload c3_kitti_1242x375x128_4path --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_base_c3.json 2>> gpurun_out/r2_base_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -s 33 -c 11 -f -o gpurun_out/r2_c4_base python bench.py --workload c4_1920x1080x256_8path_subpix_lr --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c4_ncu.log 2>&1
cat gpurun_out/r2_base_c2.json | cut -c1-400
