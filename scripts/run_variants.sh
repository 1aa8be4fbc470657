#!/bin/bash
# on the GPU box: bench every scripts/variants/*.so in turn (the box's copy of the repo is scratch)
# usage: BENCH_ARGS="--workload ..." scripts/run_variants.sh [pattern]
cd "$(dirname "$0")/.."
cp kangaroo_b200/lib/libroo_b200.so /tmp/base.so
for v in /tmp/base.so scripts/variants/${1:-*}.so; do
  cp $v kangaroo_b200/lib/libroo_b200.so
  echo -n "$(basename $v): "
  timeout 180 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>0}, [round(q['ms'],2) for q in d['roofline_passes']], round(d['aggregation']['frac_of_peak'],4))"
done
cp /tmp/base.so kangaroo_b200/lib/libroo_b200.so
