"""A user problem registered at run time (bvode_register_problem): the f / jac procedure arguments
of initialize (M:377-430) as a functor compiled into a library of its own (examples/decay_chain.cu).

CPU part: the library builds (nvcc cross-compiles without a GPU), loads, registers, and the
registry reports it.  GPU part: the integration against scipy's BDF/Radau at tight tolerance,
mass conservation, strict vs fast build."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "decay_chain.cu")
OUT = os.path.join(ROOT, "examples", "_build", "libdecay_chain.so")


def _registered():
    import dvode_b200 as B
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    so = B.plugins.build_problem(SRC, OUT)
    return B.plugins.register_problem(so)


def test_plugin_builds_and_registers():
    import dvode_b200 as B
    pid, name = _registered()
    assert name == "decay_chain" and pid >= 5
    assert "decay_chain" in B.problem_names()
    assert B.problem_info("decay_chain") == (pid, 4, 3)
    # registering the same library again is idempotent on the Python side
    assert B.plugins.register_problem(OUT)[0] == pid
    # a stale ABI tag is refused
    import ctypes as C
    lib = B.api._lib
    assert lib.bvode_register_problem(C.c_void_p(1), None, lib.bvode_plugin_abi() + 1) < 0


@pytest.mark.gpu
def test_plugin_problem_against_scipy():
    from scipy.integrate import solve_ivp
    import dvode_b200 as B
    _registered()
    nsys = 257
    k = np.array([0.05, 1.0e4, 2.0]) * (1.0 + 0.2 * B.batches.uniform_pm1(nsys, 3, 7, 0))
    y0 = np.array([1.0, 0.0, 0.0, 0.0])
    touts = np.array([0.1, 1.0, 10.0, 100.0])
    rtol, atol = [1e-7], [1e-12]
    res = {}
    for strict in (False, True):
        for mf in (21, 22):
            s = B.dvode_t().initialize("decay_chain", nsys, k, strict=strict)
            y = np.tile(y0, (nsys, 1)); t = np.zeros(nsys); ist = np.ones(nsys, dtype=np.int32)
            rw = np.zeros((nsys, 20)); iw = np.zeros((nsys, 30), dtype=np.int32); iw[0, 5] = 5000
            yo = np.zeros((touts.size, nsys, 4))
            s.solve_schedule(4, y, t, touts, 1, rtol, atol, 1, ist, 1, rw, 20, iw, 30, mf, yo)
            assert (ist == 2).all()
            assert np.abs(yo.sum(axis=2) - 1.0).max() < 1e-6
            res[(strict, mf)] = yo
            s.close()
    # the two builds and the analytic / finite-difference Jacobians agree to integration accuracy
    for key in res:
        assert np.abs(res[key] - res[(True, 21)]).max() < 2e-5, key
    for i in (0, 100, 256):
        def f(t, y, kk=k[i]):
            r1, r2, r3 = kk[0] * y[0], kk[1] * y[1] * y[1], kk[2] * y[2]
            return [-r1, r1 - r2, r2 - r3, r3]
        ref = solve_ivp(f, (0.0, 100.0), y0, method="Radau", t_eval=touts, rtol=1e-10, atol=1e-14)
        assert ref.success
        err = np.abs(res[(False, 21)][:, i, :] - ref.y.T)
        assert err.max() < 2e-5, (i, err.max())
