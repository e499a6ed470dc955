#!/bin/bash
# ncu --set full captures of the two warp-per-system kernels (after a plain run exits 0)
tag=${1:-rXX}
python tools/time_configs.py --nsys 65536 --reps 1 > gpurun_out/${tag}_warp_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_advection -f python tools/time_configs.py --nsys 65536 --reps 1 --which advection > gpurun_out/${tag}_ncu_adv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
cat gpurun_out/${tag}_warp_plain.log
