#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2e_pytest.log
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib timeout 300 python tools/time_configs.py --nsys 262144 --reps 3 --which advection 2>&1 | grep "rep 2"
done
BVODE_LIB=$PWD/dvode_b200/libbvode.so timeout 300 python tools/time_configs.py --nsys 65536 --reps 2 --which mech32 2>&1 | grep "rep 1"
