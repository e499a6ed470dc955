#!/bin/bash
# usage: tools/ncu_report.sh <tag> <object.o> <title>  -> profiles/r01_ncu_thread_kernel_<tag>.txt (+ by-region)
tag=$1; obj=$2; title=$3
K='18vode_thread_kernelINS_12ThreadPolicyINS_9RobertsonELi1ELb1ELi2EEELb0ELb0E'
python tools/ncu_summary.py gpurun_out/${tag}_thread.ncu-rep "$title" > profiles/r01_ncu_thread_kernel_${tag}.txt
ncu -i gpurun_out/${tag}_thread.ncu-rep --page source --csv > /tmp/${tag}_src.csv 2>/dev/null
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $(realpath $obj) > /dev/null && nvdisasm -g -c *.cubin > /tmp/${tag}.dis)
python tools/ncu_by_region.py /tmp/${tag}_src.csv /tmp/${tag}.dis $K > profiles/r01_ncu_thread_kernel_${tag}_by_region.txt
