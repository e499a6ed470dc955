#!/bin/bash
# round 2: warp-kernel experiments: the relevant parity tests with the in-tree library, then timings of every library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "${2:-lu_ or mech32 or singular or per_system or adapter or reference_form}" 2>&1 | tail -8
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  echo "== $lib"
  BVODE_LIB=$PWD/$lib timeout 300 python tools/time_configs.py --nsys 65536 --reps 2 --which ${1:-mech32} 2>&1 | grep "rep 1"
done
