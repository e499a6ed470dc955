#!/bin/bash
# round 2 verification pass: tests, bench, launch list, ncu full captures of the three hot kernels
tag=${1:-r2j}
bash tools/gpu_round.sh $tag
bash tools/gpu_ncu_warp.sh $tag
python tools/time_diffusion48.py 2>&1 | tail -4
