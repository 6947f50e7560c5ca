#!/bin/bash
# Round 2, GPU call H (8 GPUs): GPU suite, gathered == single-GPU at world 8, 8-GPU bench lines (C2, C3) with the
# peer-interleaved pull kernel and the SM-partitioned overlap variants.
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    tests/dist_gather_check.py 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C2_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --workload C3 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C3_n8.json
