#!/bin/bash
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib timeout 300 python tools/time_diffusion48.py 2>&1 | grep "2^16"
done
