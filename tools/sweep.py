#!/usr/bin/env python3
"""BASELINE.json configs[4]: batch-size sweep of the device-resident kernels on one GPU.
Robertson mf=21 (rtol 1e-6, 12 decades) 2^16..2^24, advection mf=24 and mech32 mf=22 2^14..2^20.
Prints one JSON object; usage: python tools/sweep.py [--max-rob 24] [--max-warp 20] > profiles/r01_sweep.json"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dvode_b200 as B
from dvode_b200 import batches

ap = argparse.ArgumentParser()
ap.add_argument("--max-rob", type=int, default=24)
ap.add_argument("--max-warp", type=int, default=20)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()


def run(name, mf, lg, par_fn, y0, touts, tol, iopt, ml, mu, mxstep):
    nsys = 1 << lg
    par = par_fn(nsys)
    s = B.dvode_t().initialize(name, nsys, par)
    neq = s.neq
    y0d = torch.tensor(np.ascontiguousarray(np.broadcast_to(np.asarray(y0)[:, None], (neq, nsys))), dtype=torch.float64, device=dev)
    yd = torch.empty_like(y0d)
    td = torch.zeros(nsys, dtype=torch.float64, device=dev)
    isd = torch.ones(nsys, dtype=torch.int32, device=dev)
    iw = np.zeros(30, dtype=np.int32); iw[0], iw[1], iw[5] = ml, mu, mxstep
    rw = np.zeros(20)
    best = None
    for r in range(a.reps + 1):  # the first pass is the warm-up
        yd.copy_(y0d); td.zero_(); isd.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        s.solve_schedule_device(neq, yd.data_ptr(), td.data_ptr(), touts, tol["itol"], tol["rtol"], tol["atol"], 1,
                                isd.data_ptr(), iopt, rw, iw, mf, 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if r > 0:
            best = ms if best is None else min(best, ms)
    ok = int((isd == 2).sum().item())
    s.close()
    del y0d, yd, td, isd
    torch.cuda.empty_cache()
    return {"problem": name, "mf": mf, "log2_nsys": lg, "ms": round(best, 3), "systems_per_s": round(nsys / best * 1e3, 1), "ok": ok == nsys}


out = {"what": "device-resident schedule launch, best of %d after one warm-up, CUDA events; no output points stored" % a.reps,
       "gpu": torch.cuda.get_device_name(0), "rows": []}
for lg in range(16, a.max_rob + 1, 2):
    out["rows"].append(run("robertson", 21, lg, batches.robertson_batch, batches.ROBERTSON_Y0, batches.robertson_touts(),
                           batches.ROBERTSON_TOL_HEADLINE, 0, 0, 0, 0))
    print(out["rows"][-1], file=sys.stderr)
for lg in range(14, a.max_warp + 1, 2):
    out["rows"].append(run("advection", 24, lg, batches.advection_batch, batches.advection_y0(), batches.advection_touts(),
                           batches.ADVECTION_TOL, 0, 5, 0, 0))
    print(out["rows"][-1], file=sys.stderr)
    out["rows"].append(run("mech32", 22, lg, batches.mech32_batch, batches.mech32_y0(), batches.mech32_touts(),
                           batches.MECH32_TOL, 1, 0, 0, 100000))
    print(out["rows"][-1], file=sys.stderr)
print(json.dumps(out, indent=1))
