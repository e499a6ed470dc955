#!/bin/bash
# time the headline thread kernel (Robertson, 2^20 systems, device-resident) with every library under exp_libs/ and the in-tree one
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib python tools/prof_run.py --nsys 1048576 --reps 3 2>&1 | tail -2
done
