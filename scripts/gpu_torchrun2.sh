#!/bin/bash
# 2-GPU check of the bench contract under torchrun (NCCL only for the barrier / max-over-ranks reduction):
# stdout must be exactly ONE JSON line (NCCL's version banner and everything else goes to stderr)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_2gpu.json 2> gpurun_out/r2_bench_c2_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/r2_bench_c2_2gpu.json)"; head -c 300 gpurun_out/r2_bench_c2_2gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null > gpurun_out/r2_bench_reference_2gpu.json
echo "reference stdout lines: $(wc -l < gpurun_out/r2_bench_reference_2gpu.json)"; head -c 200 gpurun_out/r2_bench_reference_2gpu.json; echo
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | wc -l
