#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "engine or bulk_copy or in_sweep or row_strip or fused or sgm" > gpurun_out/r2_call22_tests.log 2>&1
tail -4 gpurun_out/r2_call22_tests.log
for flag in "" "--materialised-cost"; do
for wl in c1_640x480x64_4path c3_kitti_1242x375x128_4path; do
  echo -n "$wl $flag: "
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $flag 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>0}, [round(q['ms'],2) for q in d['roofline_passes']])"
done; done
timeout 600 python scripts/c5_split.py --reps 3 --gpus 1 2>&1 | tail -2 | cut -c1-300
