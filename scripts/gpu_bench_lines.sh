#!/bin/bash
# bench lines of every BASELINE config (after profiles/r2_traffic.json has been regenerated for the current kernel sources)
mkdir -p gpurun_out/final2
O=gpurun_out/final2
python bench.py --steps 20 --warmup 5 > $O/bench_c2.json 2> $O/bench.err
for wl in c1_640x480x64_4path c3_kitti_1242x375x128_4path c4_1920x1080x256_8path_subpix_lr c5_3840x2160x256_8path_subpix_lr_single_gpu; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${wl%%_*}.json 2>> $O/bench.err
done
python bench.py --window 16x16 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2_16x16.json 2>> $O/bench.err
python bench.py --materialised-cost --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c2_materialised_cost.json 2>> $O/bench.err
cut -c1-200 $O/bench_c2.json; tail -3 $O/bench.err
