#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu full capture of the headline kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_b2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vode_thread_kernel -s 1 -c 1 \
    -o gpurun_out/${tag}_thread -f python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
cat gpurun_out/${tag}_prof_plain.log
python tools/time_configs.py --nsys 65536 --reps 2 2>&1 | tail -6
