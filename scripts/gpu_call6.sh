#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "in_sweep or bulk_copy or engine_ieee or c2_full or random_shapes" > gpurun_out/r2_call6_tests.log 2>&1
tail -5 gpurun_out/r2_call6_tests.log
echo -n "materialised: "; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --materialised-cost | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>0})"
STEPS=10 scripts/run_variants.sh 'hs_*' 2>&1 | tee gpurun_out/r2_call6_hs.txt
