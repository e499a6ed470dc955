#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/gather_bench.py 2>/dev/null | grep world; }
run X=1 | tee gpurun_out/r2_gather_n$N.jsonl
run NCCL_MIN_P2P_NCHANNELS=16 | tee -a gpurun_out/r2_gather_n$N.jsonl
run NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 | tee -a gpurun_out/r2_gather_n$N.jsonl
run NCCL_P2P_LEVEL=NVL NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216 | tee -a gpurun_out/r2_gather_n$N.jsonl
