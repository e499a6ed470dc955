#!/usr/bin/env python3
"""debug: istate = 3 with a smaller neq on diffusion48 (two components per lane) against the oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dvode_b200 as B
import oracle_lib as O
mf, neq2 = int(sys.argv[1]), int(sys.argv[2])
nsys = int(sys.argv[3]) if len(sys.argv) > 3 else 4
par, y0, touts = B.batches.diffusion48_batch(nsys), B.batches.diffusion48_y0(), list(B.batches.diffusion48_touts())[:3]
rtol, atol, nsplit = [1e-6], [1e-9], 2
kw = dict(iopt=1, mxstep=50000, iopt2=1, mxstep2=50000)
iw = np.zeros(30, dtype=np.int32); iw[5] = 50000
ref = O.batch_solve2("diffusion48", par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1, neq2=neq2, **kw)
s = B.dvode_t().initialize("diffusion48", nsys, par, strict=True)
n0 = s.neq
y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, n0))).copy()
t = np.zeros(nsys); ist = np.ones(nsys, dtype=np.int32); rw = np.zeros((nsys, 20)); iwk = np.zeros((nsys, 30), dtype=np.int32); iwk[0] = iw
for k, to in enumerate(touts):
    neq = n0 if k < nsplit else neq2
    if k == nsplit:
        ist[:] = 3
        y = np.ascontiguousarray(y[:, :neq2])
    s.solve(neq, y, t, to, 1, rtol, atol, 1, ist, 1, rw, 20, iwk, 30, mf)
    print("k", k, "ist", ist, "eq", np.array_equal(y, ref["y"][:, k, :neq]))
    print("  gpu istats", iwk[0, 10:22]); print("  ref istats", ref["istats"][0, k])
    print("  gpu rstats", rw[0, 10:14]); print("  ref rstats", ref["rstats"][0, k])
    d = np.abs(y - ref["y"][:, k, :neq]); bad = np.nonzero(d.max(axis=1) > 0)[0]
    print("  systems that differ:", bad)
    for b in bad[:3]:
        print("   sys", b, "gpu", iwk[b, 10:22], rw[b, 10:14]); print("   sys", b, "ref", ref["istats"][b, k], ref["rstats"][b, k])
        print("   diff per comp:", np.array2string(d[b], precision=1, max_line_width=250))
