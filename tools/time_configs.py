#!/usr/bin/env python3
"""Device-resident timing of the non-headline configs (3: banded advection, 4: mech32).
usage: python tools/time_configs.py [--nsys N] [--reps R] [--which advection|mech32|both]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dvode_b200 as B
from dvode_b200 import batches

ap = argparse.ArgumentParser()
ap.add_argument("--nsys", type=int, default=1 << 18)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--which", default="both")
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfgs = []
if a.which in ("both", "advection"):
    for mf in (24, 25):
        cfgs.append(("advection", mf, batches.advection_batch(a.nsys), batches.advection_y0(),
                     batches.advection_touts(), batches.ADVECTION_TOL, 0, (5, 0, 0)))
if a.which in ("both", "mech32", "mech32_21"):
    cfgs.append(("mech32", 21 if a.which == "mech32_21" else 22, batches.mech32_batch(a.nsys), batches.mech32_y0(),
                 batches.mech32_touts(), batches.MECH32_TOL, 1, (0, 0, 100000)))
for name, mf, par, y0, touts, tol, iopt, (ml, mu, mxstep) in cfgs:
    s = B.dvode_t().initialize(name, a.nsys, par)
    neq = s.neq
    y0d = torch.tensor(np.tile(y0[:, None], (1, a.nsys)), dtype=torch.float64, device=dev)
    yd = torch.empty_like(y0d)
    td = torch.zeros(a.nsys, dtype=torch.float64, device=dev)
    isd = torch.ones(a.nsys, dtype=torch.int32, device=dev)
    iw = np.zeros(30, dtype=np.int32); iw[0], iw[1], iw[5] = ml, mu, mxstep
    rw = np.zeros(20)
    st = torch.cuda.current_stream()
    for r in range(a.reps):
        yd.copy_(y0d); td.zero_(); isd.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        s.solve_schedule_device(neq, yd.data_ptr(), td.data_ptr(), touts, tol["itol"], tol["rtol"],
                                tol["atol"], 1, isd.data_ptr(), iopt, rw, iw, mf, 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s mf=%d nsys=%d rep %d: %.1f ms, %.0f systems/s, ok=%d" %
              (name, mf, a.nsys, r, ms, a.nsys / ms * 1e3, int((isd == 2).sum().item())))
    s.close()
