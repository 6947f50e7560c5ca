#!/bin/bash
# Round 2, GPU call E (2 GPUs): whole GPU suite again (tcgen05 pipeline, in-kernel gather barrier, constant-bank DCT),
# 2-GPU bench lines with the three gather variants, op timings.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 10 2>&1 | tail -2 | tee gpurun_out/r2_bench_C2_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload C3 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/r2_bench_C3_n2.json
timeout 300 python tools/bench_ops.py 2>&1 | tail -9
timeout 300 python tools/bench_wave_ops.py 2>&1 | tail -6
