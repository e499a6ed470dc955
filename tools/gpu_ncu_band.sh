#!/bin/bash
# ncu --set full capture of the banded warp kernel (after a plain run exits 0)
tag=${1:-rXX}
mkdir -p gpurun_out
python tools/time_configs.py --nsys 65536 --reps 2 --which advection > gpurun_out/${tag}_band_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_advection -f python tools/time_configs.py --nsys 65536 --reps 1 --which advection > gpurun_out/${tag}_ncu_adv.log 2>&1
cat gpurun_out/${tag}_band_plain.log
