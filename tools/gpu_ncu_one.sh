#!/bin/bash
# one ncu --set full capture of a warp kernel, exported as CSV on the box.  usage: gpu_ncu_one.sh <tag> <advection|mech32> [nsys]
tag=$1; which=$2; n=${3:-65536}
mkdir -p gpurun_out
python tools/time_configs.py --nsys $n --reps 2 --which $which 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o /tmp/${tag}_${which} -f python tools/time_configs.py --nsys $n --reps 1 --which $which > gpurun_out/${tag}_ncu_${which}.log 2>&1
ncu -i /tmp/${tag}_${which}.ncu-rep --page raw --csv > gpurun_out/${tag}_${which}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_${which}.ncu-rep --page source --csv > gpurun_out/${tag}_${which}_src.csv 2>/dev/null
ls -la gpurun_out | tail -4
