#!/usr/bin/env python3
"""Bandwidth of the final gather (bvode_gather_device) alone, under torchrun: every rank holds the results of
2^20 Robertson systems (solved once), then the gather to rank 0 is timed.  NCCL tunables come from the environment.
usage: torchrun --nproc-per-node N tools/gather_bench.py [--nsys 1048576] [--reps 20]"""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dvode_b200 as B
from dvode_b200 import batches, sharding

ap = argparse.ArgumentParser()
ap.add_argument("--nsys", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = sharding.make_comm(dist, rank, world, local)
nsys = a.nsys
s = B.dvode_t().initialize("robertson", nsys, batches.robertson_batch(nsys, first=rank * nsys), device=local)
y = torch.zeros(3, nsys, dtype=torch.float64, device=dev); y[0].fill_(1.0)
t = torch.zeros(nsys, dtype=torch.float64, device=dev)
ist = torch.ones(nsys, dtype=torch.int32, device=dev)
tol = batches.ROBERTSON_TOL_HEADLINE
st = torch.cuda.current_stream()
s.solve_schedule_device(3, y.data_ptr(), t.data_ptr(), batches.robertson_touts()[:2], 2, tol["rtol"], tol["atol"], 1,
                        ist.data_ptr(), 0, None, None, 21, 0, st.cuda_stream)
n = nsys * world
if rank == 0:
    bufs = [torch.empty(n, 3, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev),
            torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, 4, dtype=torch.float64, device=dev),
            torch.empty(n, 12, dtype=torch.int32, device=dev)]
    ptrs = [b.data_ptr() for b in bufs]
else:
    ptrs = [1] * 5
for _ in range(3):
    s.gather_device(comm, 0, n, *ptrs, stream=st.cuda_stream)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ms = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(st)
    s.gather_device(comm, 0, n, *ptrs, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
tt = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    byt = (n - nsys) * 116
    ok = bool((bufs[2] == 2).all().item()) and bool(torch.equal(bufs[0][:nsys], y.t().contiguous()))
    print(json.dumps({"world": world, "nsys_per_rank": nsys, "gather_ms_median": float(tt.item()), "bytes_into_rank0": byt,
                      "GBs_into_rank0": byt / (float(tt.item()) * 1e-3) / 1e9, "ok": ok, "nccl": B.nccl_version(),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}))
comm.close()
dist.destroy_process_group()
