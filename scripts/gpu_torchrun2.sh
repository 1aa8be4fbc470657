#!/bin/bash
# 2-GPU check of the bench contract under torchrun (NCCL only for the barrier / max-over-ranks reduction)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_2gpu.json 2> gpurun_out/r2_bench_c2_2gpu.err
tail -c 600 gpurun_out/r2_bench_c2_2gpu.json; tail -3 gpurun_out/r2_bench_c2_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
