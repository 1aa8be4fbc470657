#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, bench lines of every BASELINE config, ncu launch list and
# --set full summaries (raw-page CSV only; the reports stay on the box).
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1
tail -3 $O/gpu_tests.log
ROO_STRESS_SEEDS=60 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "random_shapes" > $O/gpu_stress.log 2>&1
tail -2 $O/gpu_stress.log
python bench.py --steps 20 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench_c2.err
for wl in c1_640x480x64_4path c3_kitti_1242x375x128_4path c4_1920x1080x256_8path_subpix_lr c5_3840x2160x256_8path_subpix_lr_single_gpu; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${wl%%_*}.json 2>> $O/bench_c2.err
done
python bench.py --window 16x16 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2_16x16.json 2>> $O/bench_c2.err
python bench.py --materialised-cost --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c2_materialised_cost.json 2>> $O/bench_c2.err
# launch list of one c2 step (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 7 --csv --log-file $O/launches_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
prof() {  # tag, then bench args
  tag=$1; shift
  timeout 600 ncu --set full --clock-control none -k regex:sgm_ -s 12 -c 4 -f -o /tmp/$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $O/ncu_$tag.log 2>&1
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/ncu_${tag}_raw.csv 2>/dev/null
}
prof c2
prof c2_16x16 --window 16x16
prof c3 --workload c3_kitti_1242x375x128_4path
prof c4 --workload c4_1920x1080x256_8path_subpix_lr
prof c5 --workload c5_3840x2160x256_8path_subpix_lr_single_gpu
for tool in memcheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -h 'ERROR SUMMARY' $O/sanitizer_$tool.log | tail -1)" | tee -a $O/sanitizer_summary.txt
done
ls -la $O | head -40
cut -c1-300 $O/bench_c2.json
