#!/bin/bash
# ncu --set full capture of the headline thread kernel (after a plain run of the same command exits 0)
tag=${1:-rXX}
python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vode_thread_kernel -s 1 -c 1 \
    -o gpurun_out/${tag}_thread -f python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
cat gpurun_out/${tag}_prof_plain.log
