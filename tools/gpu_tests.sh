#!/bin/bash
# the whole GPU test-suite (log kept under gpurun_out/<tag>_pytest.log)
tag=${1:-r2x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q ${2:--x} > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log; tail -25 gpurun_out/${tag}_pytest.log
