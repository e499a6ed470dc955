#!/bin/bash
# usage: tools/ncu_report_warp.sh <tag> <mech32|advection> <title>
#   -> profiles/r01_ncu_warp_<which>_<tag>.txt (summary + by-region), from gpurun_out/<tag>_<which>.ncu-rep
tag=$1; which=$2; title=$3
if [ "$which" = mech32 ]; then obj=$PWD/dvode_b200/build/reg_fast_3.o; K='16vode_warp_kernelINS_15WarpDensePolicyINS_6Mech32ELi2ELb1ELi2EEELb0E'
else obj=$PWD/dvode_b200/build/reg_fast_2.o; K='16vode_warp_kernelINS_14WarpBandPolicyINS_11Advection25ELi4ELb1ELi2EEELb0E'; fi
out=profiles/${PREFIX:-r02}_ncu_warp_${which}_${tag}.txt
if [ -f gpurun_out/${tag}_${which}_raw.csv ]; then  # CSV exports made on the GPU box (tools/gpu_final.sh)
  python tools/ncu_summary.py gpurun_out/${tag}_${which}_raw.csv "$title" > $out
  cp gpurun_out/${tag}_${which}_src.csv /tmp/${tag}_${which}_src.csv
else
  python tools/ncu_summary.py gpurun_out/${tag}_${which}.ncu-rep "$title" > $out
  ncu -i gpurun_out/${tag}_${which}.ncu-rep --page source --csv > /tmp/${tag}_${which}_src.csv 2>/dev/null
fi
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $obj > /dev/null && nvdisasm -g -c *.cubin > /tmp/${tag}_${which}.dis)
python tools/ncu_by_region.py /tmp/${tag}_${which}_src.csv /tmp/${tag}_${which}.dis $K >> $out
