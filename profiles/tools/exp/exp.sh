timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name lib workload extra
  if [ -n "$2" ]; then export DTOF_LIB=$PWD/mitsuba3dopplertof_b200/$2; else unset DTOF_LIB; fi
  timeout 200 python bench.py --workload $3 $4 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/exp_$1.json
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/exp_$1.json').read().strip().split(chr(10))[-1]); print('$1', round(d['value'],1), round(d['e2e']['value'],1), d['roofline'].get('traversal_mode'))
except Exception as e: print('$1', 'ERR', open('gpurun_out/exp_$1.json').read()[-300:])
"
}
