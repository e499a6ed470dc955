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
ap.add_argument("--sort", default="none", help="none | lex | morton: order the systems by their parameters; attempts: by measured step attempts")
ap.add_argument("--compact", action="store_true", help="BVODE_FLAG_COMPACT: finished threads take the next system")
ap.add_argument("--tail", type=int, default=0, help="heavy-tailed batch: one system in TAIL integrates to 4e10, the others to 40 (per-system tout, one call)")
a = ap.parse_args()
tol = batches.ROBERTSON_TOL_HEADLINE if a.tol == "headline" else batches.ROBERTSON_TOL_REF
touts = batches.robertson_touts()
par = batches.robertson_batch(a.nsys)
if a.sort in ("lex", "morton"):
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
if a.sort == "attempts":
    # order the systems by the number of step attempts they take (measured by a first run): what a
    # perfect refill / compaction of finished systems could gain at most
    s0 = B.dvode_t().initialize("robertson", a.nsys, par)
    y = np.tile(np.array(batches.ROBERTSON_Y0), (a.nsys, 1)); t = np.zeros(a.nsys)
    ist = np.ones(a.nsys, dtype=np.int32); iw = np.zeros((a.nsys, 30), dtype=np.int32); rw = np.zeros((a.nsys, 20))
    s0.solve_schedule(3, y, t, touts, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 0, rw, 20, iw, 30, a.mf, None)
    s0.close()
    att = iw[:, 10].astype(np.int64) + iw[:, 20] + iw[:, 21]
    print("attempts: mean %.1f std %.1f max %d; mean/max-in-warp %.4f" % (
        att.mean(), att.std(), att.max(), att.mean() / att.reshape(-1, 32).max(axis=1).mean()))
    par = np.ascontiguousarray(par[np.argsort(att, kind="stable")])
s = B.dvode_t().initialize("robertson", a.nsys, par, compact=a.compact)
for r in range(a.reps):
    y = np.tile(np.array(batches.ROBERTSON_Y0), (a.nsys, 1))
    t = np.zeros(a.nsys)
    ist = np.ones(a.nsys, dtype=np.int32)
    t0 = time.perf_counter()
    if a.tail:
        tps = np.full(a.nsys, 40.0)
        tps[(np.arange(a.nsys) * 2654435761 % a.nsys) % a.tail == 0] = 4e10
        iw = np.zeros((a.nsys, 30), dtype=np.int32); iw[0, 5] = 100000
        rw = np.zeros((a.nsys, 20))
        t0 = time.perf_counter()
        s.solve(3, y, t, tps, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 1, rw, 20, iw, 30, a.mf)
    else:
        s.solve_schedule(3, y, t, touts, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 0, None, 0, None,
                         0, a.mf, None)
    dt = time.perf_counter() - t0
    print("rep %d: %.1f ms, %.0f systems/s, ok=%d" % (r, dt * 1e3, a.nsys / dt, int((ist == 2).sum())))
