#!/bin/bash
# usage (on the GPU box, from the repo root): tools/gpu_check.sh <label> [ncu]
# parity tests -> bench line -> (optional) one ncu --set full capture of the fused kernel
label=${1:-run}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${label}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('VALUE %.1f h/s  %.2f us/launch  roofline frac %.4f  e2e %.1f h/s  clocks %s' % (d['value'], r['avg_launch_us'], r['frac'], d['e2e']['value'], d['clocks']))"
if [ "$2" == "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:logmel -s 10 -c 1 -o gpurun_out/prof_${label} -f python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${label}.log 2>&1
  ls -la gpurun_out/prof_${label}.ncu-rep
fi
