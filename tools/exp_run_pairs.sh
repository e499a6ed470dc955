#!/bin/bash
# libraries named p2* carry a rebuilt band unit (time advection), p3* a rebuilt dense unit (time mech32)
echo "== in-tree"; python tools/time_configs.py --nsys 65536 --reps 2 2>&1 | grep "rep 1"
for lib in exp_libs/*.so; do
  echo "== $lib"
  case $lib in *p2*) w=advection;; *) w=mech32;; esac
  BVODE_LIB=$PWD/$lib python tools/time_configs.py --nsys 65536 --reps 2 --which $w 2>&1 | grep "rep 1"
done
