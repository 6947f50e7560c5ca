#!/bin/bash
# End-of-round evidence on one B200: GPU suite, bench lines of every workload, launch list and one full ncu capture of the
# product kernel.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -x
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for w in C2 C3 C4 C5; do
  timeout 300 python bench.py --workload $w $([ $w != C2 ] && echo --no-cpu-baseline) 2>&1 | tail -1 > gpurun_out/r2_final_bench_$w.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench_$w.json').read()); r=d['roofline']
print('$w VALUE %.1f h/s  %.2f us/launch  frac %.4f  e2e %.1f h/s  clocks %s' % (d['value'], r['avg_launch_us'], r['frac'], d['e2e']['value'], d['clocks']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel_fast -s 10 -c 1 -o gpurun_out/prof_r2_final -f python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_final.log 2>&1
ls -la gpurun_out/prof_r2_final.ncu-rep
