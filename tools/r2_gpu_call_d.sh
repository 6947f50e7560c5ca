#!/bin/bash
# Round 2, GPU call D (2 GPUs): tcgen05 mel GEMM parity, the 2-rank gather test, the 2-GPU bench line, op timings.
set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k tcgen05 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q -k "not tcgen05" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 10 2>&1 | tail -3 | tee gpurun_out/r2_bench_C2_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload C3 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/r2_bench_C3_n2.json
timeout 300 python tools/bench_ops.py 2>&1 | tail -3
