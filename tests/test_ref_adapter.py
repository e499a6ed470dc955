"""User callbacks in the reference's own form -- f(neq, t, y, ydot), jac(neq, t, y, ml, mu, pd, nrowpd), M:316-334 --
on the one-system-per-warp kernels (csrc/bvode_ref_problem.cuh): examples/advection_ref.cu and examples/vdp_ref.cu
are dvdemo's f2 / jac2 and f1 / jac1 argument for argument.  They must give the bits of the oracle (strict build),
hence the bits of the lane-wise built-in forms."""
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _register(name):
    import dvode_b200 as B
    out = os.path.join(ROOT, "examples", "_build", "lib%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    so = B.plugins.build_problem(os.path.join(ROOT, "examples", name + ".cu"), out)
    return B.plugins.register_problem(so)


def test_reference_form_problems_build_and_register():
    import dvode_b200 as B
    for name, neq, npar in (("advection_ref", 25, 2), ("vdp_ref", 2, 1)):
        pid, got = _register(name)
        assert got == name and B.problem_info(name) == (pid, neq, npar)


def _run(problem, par, y0, touts, rtol, atol, mf, strict, iw0=None):
    import dvode_b200 as B
    nsys = par.shape[0]
    s = B.dvode_t().initialize(problem, nsys, par, strict=strict)
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, s.neq))).copy()
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20))
    iw = np.zeros((nsys, 30), dtype=np.int32)
    if iw0 is not None:
        iw[0, :len(iw0)] = iw0
    yo = np.zeros((len(touts), nsys, s.neq))
    s.solve_schedule(s.neq, y, t, np.asarray(touts), 1, rtol, atol, 1, ist, 0, rw, 20, iw, 30, mf, yo)
    s.close()
    return yo.transpose(1, 0, 2), ist, iw[:, 10:22].copy()


@pytest.mark.gpu
@pytest.mark.parametrize("mf", [24, 25, -24, 20, 23, 14])
def test_advection_reference_form_equals_oracle_and_lanewise_form(mf):
    import dvode_b200 as B
    _register("advection_ref")
    nsys = 48
    par = B.batches.advection_batch(nsys)
    touts = B.batches.advection_touts()
    iw0 = [5, 0]
    ref = O.batch_solve("advection", par, B.batches.advection_y0(), 0.0, touts, [0.0], [1e-6], mf, itol=1, ml=5, mu=0,
                        pow_mode=1)
    a = _run("advection_ref", par, B.batches.advection_y0(), touts, [0.0], [1e-6], mf, True, iw0)
    b = _run("advection", par, B.batches.advection_y0(), touts, [0.0], [1e-6], mf, True, iw0)
    assert (a[1] == 2).all()
    assert np.array_equal(a[0], ref["y"]) and np.array_equal(a[2], ref["istats"][:, -1])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    # fast build: the same integration to tolerance
    c = _run("advection_ref", par, B.batches.advection_y0(), touts, [0.0], [1e-6], mf, False, iw0)
    # (the diagonal approximation, miter 3, is loose by itself: the reference's own error overrun there is > 100)
    assert (c[1] == 2).all() and np.abs(c[0] - ref["y"]).max() < (2e-3 if abs(mf) % 10 == 3 else 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("mf", [21, 22, -21, 10, 23])
def test_vdp_reference_form_on_the_dense_warp_kernel_equals_oracle(mf):
    import dvode_b200 as B
    _register("vdp_ref")
    nsys = 40
    par = B.batches.perturbed((3.0,), nsys)
    touts = [1.39283880203, 3.6076126770299997, 5.82238655203, 8.03716042703]
    ref = O.batch_solve("vdp", par, [2.0, 0.0], 0.0, touts, [0.0], [1e-6], mf, itol=1, pow_mode=1)
    a = _run("vdp_ref", par, [2.0, 0.0], touts, [0.0], [1e-6], mf, True)
    assert (a[1] == 2).all()
    assert np.array_equal(a[0], ref["y"]) and np.array_equal(a[2], ref["istats"][:, -1])
    c = _run("vdp_ref", par, [2.0, 0.0], touts, [0.0], [1e-6], mf, False)
    assert (c[1] == 2).all()
    w = np.abs(c[0] - ref["y"]) / 1e-6
    assert np.median(w) < 10.0 and abs(int(c[2][:, 0].sum()) - int(ref["istats"][:, -1, 0].sum())) <= 0.03 * ref["istats"][:, -1, 0].sum()
