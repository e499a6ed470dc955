#!/bin/bash
# dense warp kernel after a change: LU known answers + mech32 parity tests, timing, ncu capture
tag=${1:-r2n}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "lu or mech32 or dense or band or advection" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
python tools/time_configs.py --nsys 65536 --reps 2 2>&1 | tail -6
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
