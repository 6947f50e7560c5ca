for w in 16 20 24; do echo "== warps $w"; B200MEL_WARPS=$w python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('VALUE %.1f h/s  %.2f us/launch  frac %.4f' % (d['value'], r['avg_launch_us'], r['frac']))"; done
