#!/bin/bash
# kernel-only duration (ncu gpu__time_duration) of the compacting thread kernel for each library under exp_libs/
# on the homogeneous headline batch (tail 0) and on heavy-tailed batches (one system in TAIL runs to 4e10, the rest to 40)
for lib in dvode_b200/libbvode.so exp_libs/*.so; do
  for tl in 0 16 64 512; do
    for f in "" "--compact"; do
      [ "$lib" != dvode_b200/libbvode.so ] && [ -z "$f" ] && continue
      ns=$(BVODE_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vode_thread_kernel --csv python tools/prof_run.py --nsys 1048576 --reps 2 --tail $tl $f 2>/dev/null | grep vode_thread | tail -1 | awk -F'","' '{print $NF}' | tr -d '"')
      echo "$lib tail=$tl $f : $ns ns"
    done
  done
done
