#!/bin/bash
# head-of-session check: dense/band parity subset, warp timings, ncu full captures of the thread kernel and the dense warp kernel
tag=${1:-r2k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "lu or mech32 or dense" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
python tools/time_configs.py --nsys 65536 --reps 2 2>&1 | tail -6
python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vode_thread_kernel -s 1 -c 1 \
    -o gpurun_out/${tag}_thread -f python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
cat gpurun_out/${tag}_prof_plain.log
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o gpurun_out/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
ls -la gpurun_out
