#!/bin/bash
# blocked band solve: LU known-answer tests, band parity tests, timing of variants, ncu capture
tag=${1:-r2l}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "band or advection or lu or bandwidth" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
bash tools/exp_run_band.sh
bash tools/gpu_ncu_band.sh $tag
