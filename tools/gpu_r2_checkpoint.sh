#!/bin/bash
# One 1-GPU pass for the record: tests, bench line, ncu launch list, ncu --set full of the three kernels.
tag=${1:-r2g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-strict > gpurun_out/${tag}_b2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-strict > gpurun_out/${tag}_ncu_launch.log 2>&1
python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vode_thread_kernel -s 1 -c 1 \
    -o gpurun_out/${tag}_thread -f python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_ncu_thread.log 2>&1
python tools/time_configs.py --nsys 65536 --reps 1 > gpurun_out/${tag}_warp_plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_advection -f python tools/time_configs.py --nsys 65536 --reps 1 --which advection > gpurun_out/${tag}_ncu_adv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
}
cat gpurun_out/${tag}_prof_plain.log gpurun_out/${tag}_warp_plain.log
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g strict %.4g roof %.3f cpu %s" % (d["value"], d["e2e"]["value"], d.get("strict", {}).get("value", 0), d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value")))
for o in d.get("other_configs", []): print(o["workload"][:30], "%.4g" % o["value"], "e2e %.4g" % o["e2e"]["value"], "roof %.3f" % o["roofline"]["frac"], o["systems_ok"])
PY
