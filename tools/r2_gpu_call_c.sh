#!/bin/bash
# Round 2, GPU call C (1 GPU): whole GPU suite, bench line (e2e through b200mel_forward_host + PCIe probe), the C3/C4/C5
# shard benches, ncu launch list + one --set full capture of the shipped fused kernel, op timings.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 400 --warmup 20 | tee gpurun_out/r2_bench_C2.json
for w in C3 C4 C5; do python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2_bench_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logmel_fast -s 10 -c 1 -f -o gpurun_out/prof_r2_fast \
    python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/bench_ops.py 2>&1 | tail -20
python tools/bench_wave_ops.py 2>&1 | tail -20
