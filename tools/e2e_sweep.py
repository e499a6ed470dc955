#!/usr/bin/env python3
"""End-to-end (host buffers, headers) timing of the headline batch for chunk / stream settings from the environment."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dvode_b200 as B
from dvode_b200 import batches
nsys = 1 << 20
tol = batches.ROBERTSON_TOL_HEADLINE; touts = batches.robertson_touts()
par = torch.from_numpy(batches.robertson_batch(nsys)).pin_memory().numpy()
s = B.dvode_t().initialize("robertson", nsys, par)
pin = lambda *sh, dt=torch.float64: torch.empty(*sh, dtype=dt).pin_memory().numpy()
y, t, ist = pin(nsys, 3), pin(nsys), pin(nsys, dt=torch.int32)
yo, rw, iw = pin(12, nsys, 3), pin(nsys, 20), pin(nsys, 30, dt=torch.int32)
rw[:] = 0; iw[:] = 0
ts = []
for rep in range(5):
    y[:] = batches.ROBERTSON_Y0; t[:] = 0; ist[:] = 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.set_params(par)
    s.solve_schedule(3, y, t, touts, 2, tol["rtol"], tol["atol"], 1, ist, 0, rw, 20, iw, 30, 21, yo)
    ts.append(time.perf_counter() - t0)
print("chunk %s streams %s: e2e best %.2f ms median %.2f ms -> %.3f M systems/s, ok=%d" % (
    os.environ.get("BVODE_CHUNK_LOG2", "17"), os.environ.get("BVODE_HOST_STREAMS", "2"), min(ts[1:]) * 1e3,
    float(np.median(ts[1:])) * 1e3, nsys / float(np.median(ts[1:])) / 1e6, int((ist == 2).sum())))
