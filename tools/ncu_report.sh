#!/bin/bash
# usage: tools/ncu_report.sh <tag> <object.o> <title>  -> profiles/${PREFIX:-r02}_ncu_thread_kernel_<tag>.txt (+ by-region)
# from gpurun_out/<tag>_thread.ncu-rep, or from the CSV exports gpurun_out/<tag>_thread_{raw,src}.csv (tools/gpu_final.sh)
tag=$1; obj=$2; title=$3
K='18vode_thread_kernelINS_12ThreadPolicyINS_9RobertsonELi1ELb1ELi2EEELb0ELb0E'
P=profiles/${PREFIX:-r02}_ncu_thread_kernel_${tag}
if [ -f gpurun_out/${tag}_thread_raw.csv ]; then
  python tools/ncu_summary.py gpurun_out/${tag}_thread_raw.csv "$title" > $P.txt
  cp gpurun_out/${tag}_thread_src.csv /tmp/${tag}_src.csv
else
  python tools/ncu_summary.py gpurun_out/${tag}_thread.ncu-rep "$title" > $P.txt
  ncu -i gpurun_out/${tag}_thread.ncu-rep --page source --csv > /tmp/${tag}_src.csv 2>/dev/null
fi
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $(realpath $obj) > /dev/null && nvdisasm -g -c *.cubin > /tmp/${tag}.dis)
python tools/ncu_by_region.py /tmp/${tag}_src.csv /tmp/${tag}.dis $K > ${P}_by_region.txt
