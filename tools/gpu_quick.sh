#!/bin/bash
# quick perf/parity iteration: gpu parity tests, then the headline bench without the CPU leg
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('value %.0f systems/s  ms/step %.2f  e2e %.0f  ok %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['systems_ok']))
    print('counters', {k: round(v, 2) for k, v in d['counters_per_system'].items()})
"
[ -n "$1" ] && python tools/time_configs.py --nsys 65536 --reps 2 2>&1 | tail -6
exit 0
