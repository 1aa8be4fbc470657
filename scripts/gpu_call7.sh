#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "in_sweep or bulk_copy or engine or fused" > gpurun_out/r2_call7_tests.log 2>&1
tail -8 gpurun_out/r2_call7_tests.log
for flag in "" "--materialised-cost"; do
for wl in c2_1280x720x128_8path_wta c4_1920x1080x256_8path_subpix_lr c1_640x480x64_4path; do
  echo -n "$wl $flag: "
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $flag 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>0})"
done; done 2>&1 | tee gpurun_out/r2_call7_ab.txt
