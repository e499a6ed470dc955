#!/bin/bash
# closing pass of a round: all GPU tests, smoke(), the bench line, the batch-size sweep (no profiler)
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench.json
python tools/sweep.py > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"
tail -c 600 gpurun_out/${tag}_sweep.json
