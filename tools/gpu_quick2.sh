#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g strict %.4g' % (d['value'], d['e2e']['value'], d['strict']['value']))
for o in d['other_configs']: print(o['workload'][:12], '%.4g e2e %.4g' % (o['value'], o['e2e']['value']))
"
