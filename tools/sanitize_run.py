#!/usr/bin/env python3
"""Small batches of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvode_b200 as B
from dvode_b200 import batches as Bt

def run(problem, nsys, par, y0, touts, itol, rtol, atol, mf, strict, iw0=None, general=False):
    s = B.dvode_t().initialize(problem, nsys, par, strict=strict)
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, s.neq))).copy()
    t = np.zeros(nsys); ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20)); iw = np.zeros((nsys, 30), dtype=np.int32)
    if iw0 is not None: iw[0, :len(iw0)] = iw0
    yo = np.zeros((len(touts), nsys, s.neq))
    if general:  # itask = 1 call by call, then an istate = 3 restart: the full-driver kernel
        for k, to in enumerate(touts):
            if k == 1: ist[:] = 3
            s.solve(s.neq, y, t, to, itol, rtol, atol, 1, ist, 1, rw, 20, iw, 30, mf)
    else:
        s.solve_schedule(s.neq, y, t, np.asarray(touts), itol, rtol, atol, 1, ist, 1, rw, 20, iw, 30, mf, yo)
    ok = int((ist == 2).sum())
    s.close()
    print("%-12s mf=%3d strict=%d general=%d ok=%d/%d" % (problem, mf, strict, general, ok, nsys), flush=True)

for strict in (False, True):
    for general in (False, True):
        run("robertson", 64, Bt.robertson_batch(64), Bt.ROBERTSON_Y0, Bt.robertson_touts()[:4], 2, [1e-4], [1e-8, 1e-14, 1e-6], 21, strict, general=general)
        run("advection", 8, Bt.advection_batch(8), Bt.advection_y0(), Bt.advection_touts()[:3], 1, [0.0], [1e-6], 24, strict, [5, 0, 0, 0, 0, 5000], general)
        run("advection", 8, Bt.advection_batch(8), Bt.advection_y0(), Bt.advection_touts()[:3], 1, [0.0], [1e-6], 25, strict, [5, 5, 0, 0, 0, 5000], general)
        run("mech32", 8, Bt.mech32_batch(8), Bt.mech32_y0(), [1e-4, 1e-3, 1e-2], 1, [1e-6], [1e-12], 22, strict, [0, 0, 0, 0, 0, 20000], general)
        run("mech32", 8, Bt.mech32_batch(8), Bt.mech32_y0(), [1e-4, 1e-3, 1e-2], 1, [1e-6], [1e-12], -21, strict, [0, 0, 0, 0, 0, 20000], general)
        run("diffusion48", 4, Bt.diffusion48_batch(4), Bt.diffusion48_y0(), [1e-3, 1e-2, 1e-1], 1, [1e-6], [1e-9], 22, strict, [0, 0, 0, 0, 0, 20000], general)
        run("mech32", 8, Bt.mech32_batch(8), Bt.mech32_y0(), [1e-4, 1e-3], 1, [1e-6], [1e-12], 10, strict, [0, 0, 0, 0, 0, 50000], general)
