#!/bin/bash
# ncu --set full capture of the two-components-per-lane kernel (diffusion48, mf 22, 2^14 systems), CSV pages exported on the box
tag=${1:-rXX}
mkdir -p gpurun_out
python tools/time_diffusion48.py 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o /tmp/${tag}_d48 -f python tools/time_diffusion48.py > gpurun_out/${tag}_ncu_d48.log 2>&1
ncu -i /tmp/${tag}_d48.ncu-rep --page raw --csv > gpurun_out/${tag}_d48_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_d48.ncu-rep --page source --csv > gpurun_out/${tag}_d48_src.csv 2>/dev/null
ls -la gpurun_out | tail -3
