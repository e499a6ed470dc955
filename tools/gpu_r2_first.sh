#!/bin/bash
# round 2, first GPU pass: parity tests (old + new), then the new bench line
tag=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g strict %.4g roof %.3f cpu %s" % (d["value"], d["e2e"]["value"], d.get("strict", {}).get("value", 0), d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value")))
    print("parity", json.dumps(d.get("parity"))[:600])
    for o in d.get("other_configs", []): print(o["workload"][:30], "%.4g" % o["value"], "e2e %.4g" % o["e2e"]["value"], "roof %.3f" % o["roofline"]["frac"], o["systems_ok"])
    print("strong", [(s["total_systems"], "%.4g" % s["value"]) for s in d.get("strong_scaling", [])])
except Exception as e:
    print("bench parse failed", e)
PY
