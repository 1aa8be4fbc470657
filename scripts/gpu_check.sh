#!/bin/bash
# parity subset + bench lines of the main configs (development check)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/check_tests.log 2>&1
tail -4 gpurun_out/check_tests.log
for wl in c2_1280x720x128_8path_wta c4_1920x1080x256_8path_subpix_lr c3_kitti_1242x375x128_4path c1_640x480x64_4path; do
  echo -n "$wl: "
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), [round(q['ms'],2) for q in d['roofline_passes']], d['clocks']['sm_mhz'])"
done
