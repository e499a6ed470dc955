#!/bin/bash
# time advection (2^16 systems, mf 24 / 25) with every experimental library under exp_libs/ (and the in-tree one)
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib python tools/time_configs.py --nsys 65536 --reps 2 --which advection 2>&1 | grep "rep 1"
done
