#!/usr/bin/env python3
"""debug: one diffusion48 system, istate = 3 with a smaller neq, then step by step (itask = 2) against the oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dvode_b200 as B
import oracle_lib as O
mf, neq2, which, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
par = B.batches.diffusion48_batch(1, first=which)
y0 = B.batches.diffusion48_y0()
t12 = list(B.batches.diffusion48_touts())[:2]
touts = t12 + [1.0] * M
rtol, atol, nsplit = [1e-6], [1e-9], 2
ref = O.batch_solve2("diffusion48", par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1, neq2=neq2,
                     iopt=1, mxstep=50000, iopt2=1, mxstep2=50000, itask2=2)
s = B.dvode_t().initialize("diffusion48", 1, par, strict=True)
n0 = s.neq
y = np.asarray(y0, dtype=np.float64).reshape(1, n0).copy()
t = np.zeros(1); ist = np.ones(1, dtype=np.int32); rw = np.zeros((1, 20)); iwk = np.zeros((1, 30), dtype=np.int32); iwk[0, 5] = 50000
prev = None
for k, to in enumerate(touts):
    neq = n0 if k < nsplit else neq2
    if k == nsplit:
        ist[:] = 3
        y = np.ascontiguousarray(y[:, :neq2])
    s.solve(neq, y, t, to, 1, rtol, atol, 1 if k < nsplit else 2, ist, 1, rw, 20, iwk, 30, mf)
    same = np.array_equal(y[0], ref["y"][0, k, :neq]) and t[0] == ref["t"][0, k] and np.array_equal(iwk[0, 10:22], ref["istats"][0, k])
    if not same or k < nsplit + 1:
        print("k", k, "same", same, "t", t[0], ref["t"][0, k])
        print("  gpu", iwk[0, 10:22], rw[0, 10:14]); print("  ref", ref["istats"][0, k], ref["rstats"][0, k])
        if prev is not None: print("  previous step (equal): ", prev)
        if not same:
            d = np.abs(y[0] - ref["y"][0, k, :neq]); print("  ydiff", np.array2string(d, precision=1, max_line_width=250))
            break
    prev = (iwk[0, 10:22].copy(), rw[0, 10:14].copy())
else:
    print("all", len(touts), "calls equal")
