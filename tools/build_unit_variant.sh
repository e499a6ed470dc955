#!/bin/bash
# Rebuild ONE fast translation unit with extra flags and link it with the in-tree objects of the
# other units into exp_libs/libbvode_<name>.so (the in-tree library is left alone).
# usage: tools/build_unit_variant.sh <name> <unit 0..4> <extra nvcc flags...>
name=$1; unit=$2; shift 2
mkdir -p exp_libs /tmp/bv_$name
B=dvode_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -DBVODE_STRICT=0 -DBVODE_UNIT=$unit "$@" -Xptxas -v -c dvode_b200/csrc/bvode_registry.cu -o /tmp/bv_$name/reg.o 2> /tmp/bv_$name/ptxas.log || { tail -20 /tmp/bv_$name/ptxas.log; exit 1; }
objs=""
for o in $B/*.o; do
  case "$o" in *32*) continue;; esac  # the REAL32 build's objects belong to libbvode32.so
  if [ "$o" = "$B/reg_fast_$unit.o" ]; then objs="$objs /tmp/bv_$name/reg.o"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o exp_libs/libbvode_$name.so $objs -ldl && echo built exp_libs/libbvode_$name.so
