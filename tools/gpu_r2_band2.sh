#!/bin/bash
# blocked band solve: LU known-answer tests, band parity tests, timing, ncu capture
tag=${1:-r2l}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "band or advection or lu or bandwidth" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
python tools/time_configs.py --nsys 262144 --reps 2 --which advection 2>&1 | tail -4
bash tools/gpu_ncu_band.sh $tag
