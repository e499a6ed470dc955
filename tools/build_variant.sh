#!/bin/bash
# build an experimental variant of the library into exp_libs/libbvode_<name>.so
# usage: tools/build_variant.sh <name> <extra nvcc flags...>
name=$1; shift
mkdir -p exp_libs
cp dvode_b200/libbvode.so /tmp/libbvode_keep.so
BVODE_EXTRA_NVCC="$*" python -m dvode_b200.build --force -v 2>&1 | grep -A3 "fast18vode_thread_kernelINS_12ThreadPolicyINS_9RobertsonELi1ELb1" | grep -E "spill|registers"
cp dvode_b200/libbvode.so exp_libs/libbvode_$name.so
cp /tmp/libbvode_keep.so dvode_b200/libbvode.so
