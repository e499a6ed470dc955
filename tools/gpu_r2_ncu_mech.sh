#!/bin/bash
# full gpu test-suite, then an ncu --set full capture of the dense warp kernel (mech32 mf=22, 32768 systems)
tag=${1:-r2d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -15 gpurun_out/${tag}_pytest.log
python tools/time_configs.py --nsys 32768 --reps 2 --which mech32 > gpurun_out/${tag}_warp_plain.log 2>&1 || { cat gpurun_out/${tag}_warp_plain.log; exit 1; }
cat gpurun_out/${tag}_warp_plain.log
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
tail -2 gpurun_out/${tag}_ncu_mech.log
