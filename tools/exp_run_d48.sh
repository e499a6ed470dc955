#!/bin/bash
# time diffusion48 (mf 22) with every experimental library under exp_libs/ (and the in-tree one)
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib python tools/time_diffusion48.py 2>&1 | tail -1
done
