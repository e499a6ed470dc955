#!/usr/bin/env python3
"""One warm-up + one measured schedule launch of a batch, for ncu.
usage: python tools/prof_run.py [--problem robertson] [--nsys N] [--tol headline|ref] [--reps R]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvode_b200 as B  # noqa: E402
from dvode_b200 import batches  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nsys", type=int, default=1 << 17)
ap.add_argument("--tol", default="headline")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mf", type=int, default=21)
ap.add_argument("--sort", default="none", help="none | lex | morton: order the systems by their parameters")
a = ap.parse_args()
tol = batches.ROBERTSON_TOL_HEADLINE if a.tol == "headline" else batches.ROBERTSON_TOL_REF
touts = batches.robertson_touts()
par = batches.robertson_batch(a.nsys)
if a.sort != "none":
    q = (par - par.min(axis=0)) / (par.max(axis=0) - par.min(axis=0) + 1e-300)
    if a.sort == "lex":
        order = np.lexsort((q[:, 2], q[:, 1], q[:, 0]))
    else:
        b = np.minimum((q * 1024).astype(np.uint64), 1023)
        key = np.zeros(a.nsys, dtype=np.uint64)
        for bit in range(10):
            for d in range(3):
                key |= ((b[:, d] >> np.uint64(bit)) & np.uint64(1)) << np.uint64(3 * bit + d)
        order = np.argsort(key, kind="stable")
    par = np.ascontiguousarray(par[order])
s = B.dvode_t().initialize("robertson", a.nsys, par)
for r in range(a.reps):
    y = np.tile(np.array(batches.ROBERTSON_Y0), (a.nsys, 1))
    t = np.zeros(a.nsys)
    ist = np.ones(a.nsys, dtype=np.int32)
    t0 = time.perf_counter()
    s.solve_schedule(3, y, t, touts, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 0, None, 0, None,
                     0, a.mf, None)
    dt = time.perf_counter() - t0
    print("rep %d: %.1f ms, %.0f systems/s, ok=%d" % (r, dt * 1e3, a.nsys / dt, int((ist == 2).sum())))
