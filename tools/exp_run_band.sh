#!/bin/bash
# time the in-tree library and every experimental one on the banded config (2^18 systems)
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib timeout 300 python tools/time_configs.py --nsys 262144 --reps 2 --which advection 2>&1 | grep "rep 1"
done
