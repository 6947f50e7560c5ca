#!/bin/bash
# Round 2, GPU call F (2 GPUs): re-tiled spectrum kernels, gather with a reserved SM, op timings.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 300 python tools/bench_ops.py 2>&1 | tail -9
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C2_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload C3 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C3_n2.json
