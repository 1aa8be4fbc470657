#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "engine or bulk_copy or in_sweep or row_strip or c4_full" > gpurun_out/r2_call20_tests.log 2>&1
tail -4 gpurun_out/r2_call20_tests.log
for wl in c4_1920x1080x256_8path_subpix_lr c5_3840x2160x256_8path_subpix_lr_single_gpu c2_1280x720x128_8path_wta; do
  echo -n "$wl: "
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), [round(q['ms'],2) for q in d['roofline_passes']])"
done
