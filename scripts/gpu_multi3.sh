#!/bin/bash
# 8-GPU lease: split engine (tests + config 5 numbers) and the bench contract under torchrun at N = 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "row_strip or multi_gpu" 2>&1 | tail -3
timeout 600 python scripts/c5_split.py --reps 5 --out gpurun_out/r2_c5_split_final.json 2>&1 | grep '"strips"' | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_c2_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_8gpu.json')); print('N=8', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))"
