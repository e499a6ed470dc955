"""The REAL32 build (the reference's kind switch, src/dvode_kinds_module.F90:29-31): libbvode32.so is
the same CUDA source compiled with -DBVODE_REAL32 -- every `bvode_real` of the C ABI a float -- and liboracle32.so
the same C restatement with -DDVO_REAL32.  Floating literals are typed in the working precision on both sides
(like the reference's `_wp` suffix), so the two agree bit for bit.  The reference ships no REAL32 golden output:
this parity is oracle-to-GPU only ("unpinned"); the REAL64 oracle bounds it from the other side (same solution
to single-precision accuracy)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_real32_library_exports_the_whole_c_abi():
    from dvode_b200 import build
    lib = ctypes.CDLL(build.build32())
    hdr = open(os.path.join(ROOT, "include", "bvode.h")).read()
    for n in re.findall(r"BVODE_API[^;(]*?\b(bvode_\w+)\s*\(", hdr):
        assert hasattr(lib, n), n
    import dvode_b200 as B
    assert B.api32.REAL is np.float32
    assert B.api32.problem_names() == ["robertson", "vdp", "vdp_maxnorm", "vdp_ewtdrop"]


def test_real32_oracle_agrees_with_the_real64_oracle_to_single_precision():
    import dvode_b200 as B
    par = B.batches.robertson_batch(64)
    touts = B.batches.robertson_touts()[:8]
    a = O.batch_solve("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, [1e-4], [1e-6, 1e-10, 1e-6], 21, itol=2, variant="32")
    b = O.batch_solve("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, [1e-4], [1e-6, 1e-10, 1e-6], 21, itol=2)
    assert a["y"].dtype == np.float32 and (a["istate"] == 2).all() and (b["istate"] == 2).all()
    w = np.abs(a["y"] - b["y"]) / (1e-4 * np.abs(b["y"]) + np.array([1e-6, 1e-10, 1e-6]))
    assert np.median(w) < 2.0 and w.max() < 100.0
    assert abs(a["istats"][:, -1, 0].mean() / b["istats"][:, -1, 0].mean() - 1.0) < 0.1


def _gpu32(problem, par, y0, touts, rtol, atol, mf, itol):
    import dvode_b200 as B
    A = B.api32
    nsys = par.shape[0]
    s = A.dvode_t().initialize(problem, nsys, par.astype(np.float32), strict=True)
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float32), (nsys, s.neq))).copy()
    t = np.zeros(nsys, dtype=np.float32)
    ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20), dtype=np.float32)
    iw = np.zeros((nsys, 30), dtype=np.int32)
    yo = np.zeros((len(touts), nsys, s.neq), dtype=np.float32)
    s.solve_schedule(s.neq, y, t, np.asarray(touts, dtype=np.float32), itol, rtol, atol, 1, ist, 0, rw, 20, iw, 30, mf, yo)
    msg = s.message(0)
    s.close()
    return yo.transpose(1, 0, 2), t, ist, rw[:, 10:14].copy(), iw[:, 10:22].copy(), msg


@pytest.mark.gpu
@pytest.mark.parametrize("problem,mf", [("robertson", 21), ("robertson", 22), ("robertson", -21), ("vdp", 21), ("vdp", 10),
                                        ("vdp", 23), ("vdp", 12), ("vdp_maxnorm", 21), ("vdp_ewtdrop", 21)])
def test_real32_gpu_equals_real32_oracle_bit_for_bit(problem, mf):
    import dvode_b200 as B
    nsys = 96
    if problem == "robertson":
        par, y0, touts, itol = B.batches.robertson_batch(nsys), [1.0, 0.0, 0.0], B.batches.robertson_touts()[:8], 2
        rtol, atol = [1e-4], [1e-6, 1e-10, 1e-6]
    else:
        par, y0, touts, itol = B.batches.perturbed((3.0,), nsys), [2.0, 0.0], [1.0, 2.0, 4.0, 8.0, 16.0], 1
        rtol, atol = [1e-4], [1e-5]
    ref = O.batch_solve(problem, par, y0, 0.0, touts, rtol, atol, mf, itol=itol, pow_mode=1, variant="32")
    y, t, ist, rst, istt, msg = _gpu32(problem, par, y0, touts, rtol, atol, mf, itol)
    assert y.dtype == np.float32
    assert np.array_equal(ist, ref["istate"][:, -1])
    assert np.array_equal(y, ref["y"])
    assert np.array_equal(t, ref["t"][:, -1])
    assert np.array_equal(istt, ref["istats"][:, -1])
    assert np.array_equal(rst, ref["rstats"][:, -1])
    if problem == "vdp_ewtdrop":
        assert (ist == -6).all() and "ewt(i1) has become r2 <= 0" in msg
    else:
        assert (ist == 2).all()
