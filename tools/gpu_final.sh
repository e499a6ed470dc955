#!/bin/bash
# Final pass of a round on one GPU box: parity tests, bench, ncu launch list, ncu --set full captures of the three hot
# kernels.  The .ncu-rep files stay on the box (gpurun brings back at most 64 MiB): their raw and source pages are
# exported as CSV there.  usage (from the repo root, under gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_b2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
export_rep() {  # <name>: gpurun_out/<name>.ncu-rep -> <name>_raw.csv, <name>_src.csv
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_src.csv 2>/dev/null
}
python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vode_thread_kernel -s 1 -c 1 \
    -o /tmp/${tag}_thread -f python tools/prof_run.py --nsys 131072 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
export_rep ${tag}_thread
cat gpurun_out/${tag}_prof_plain.log
python tools/time_configs.py --nsys 65536 --reps 2 > gpurun_out/${tag}_warp_plain.log 2>&1 || exit 1
cat gpurun_out/${tag}_warp_plain.log
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o /tmp/${tag}_advection -f python tools/time_configs.py --nsys 65536 --reps 1 --which advection > gpurun_out/${tag}_ncu_adv.log 2>&1
export_rep ${tag}_advection
ncu --set full --clock-control none --import-source on -k regex:vode_warp_kernel -c 1 \
    -o /tmp/${tag}_mech32 -f python tools/time_configs.py --nsys 32768 --reps 1 --which mech32 > gpurun_out/${tag}_ncu_mech.log 2>&1
export_rep ${tag}_mech32
python tools/time_diffusion48.py 2>&1 | tail -3
du -sh gpurun_out
