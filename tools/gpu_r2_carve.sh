#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "adapter or reference_form" 2>&1 | tail -3
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  for cv in -1 default 100; do
    echo "== $lib carve=$cv"
    if [ $cv = default ]; then
      BVODE_LIB=$PWD/$lib timeout 300 python tools/time_configs.py --nsys 65536 --reps 2 --which mech32 2>&1 | grep "rep 1"
    else
      BVODE_CARVEOUT=$cv BVODE_LIB=$PWD/$lib timeout 300 python tools/time_configs.py --nsys 65536 --reps 2 --which mech32 2>&1 | grep "rep 1"
    fi
  done
done
