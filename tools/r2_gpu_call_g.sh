#!/bin/bash
# Round 2, GPU call G (8 GPUs): gathered == single-GPU at world 8, 8-GPU bench lines (C2, C3) with the gather variants
# and NUMA-bound end-to-end copies; the fused pre-emphasis test on the way.
set -x
nvidia-smi topo -m 2>&1 | head -14
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "preemphasis or fast_and_generic" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    tests/dist_gather_check.py 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C2_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --workload C3 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_C3_n8.json
