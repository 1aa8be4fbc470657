#!/bin/bash
# whole GPU suite, smoke(), the default bench line, front/back-end operator timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
cat gpurun_out/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-900 gpurun_out/r2_bench_default.json
timeout 300 python scripts/bench_frontback.py > gpurun_out/frontback_bench.log 2>&1; tail -3 gpurun_out/frontback_bench.log | cut -c1-300
