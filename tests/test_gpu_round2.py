"""GPU parity tests added in round 2 (all through the C ABI, strict build bit for bit vs the oracle):

  * the device LINPACK routines in isolation (bvode_lu_test): random, singular (info = k, LP:87-88) and
    exact-pivot-tie matrices on every mapping -- registers, row per lane, two rows per lane, band
  * every failure exit of the driver: istate -1, -2, -4, -5, -6 (M:1496-1579, 1625-1645) and a Newton
    matrix that is exactly singular at a step (LP:87-88 -> ierpj, M:2999) on the thread, dense-warp and
    band kernels
  * per-system tolerances and optional inputs (BVODE_FLAG_PER_SYSTEM, M:729-751, 928-965)
  * the multi-GPU entry points of the C ABI on whatever GPUs are visible
  * dvsrco into a fresh handle; a tout schedule longer than one launch with failing systems
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from test_gpu_parity import gpu_schedule

pytestmark = pytest.mark.gpu


def _B():
    import dvode_b200 as B
    return B


# ------------------------------------------------------------------ LINPACK known answers ----
def _oracle_dense(a, b):
    """dgefa + dgesl of the oracle on a[m, i, j]; returns (factors[m, i, j], ipvt, x, info)."""
    L = O.lib()
    nmat, n, _ = a.shape
    fac = np.zeros_like(a)
    ipv = np.zeros((nmat, n), dtype=np.int32)
    xs = np.zeros((nmat, n))
    info = np.zeros(nmat, dtype=np.int32)
    for m in range(nmat):
        cm = np.asfortranarray(a[m]).copy(order="F")
        ip = np.zeros(n, dtype=np.int32)
        inf = C.c_int()
        L.dvo_dgefa(cm.ctypes.data_as(C.POINTER(C.c_double)), n, n, ip.ctypes.data_as(C.POINTER(C.c_int)), C.byref(inf))
        x = b[m].copy()
        if inf.value == 0:
            L.dvo_dgesl(cm.ctypes.data_as(C.POINTER(C.c_double)), n, n, ip.ctypes.data_as(C.POINTER(C.c_int)),
                        x.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            x[:] = 0.0
        fac[m], ipv[m], xs[m], info[m] = cm, ip, x, inf.value
    return fac, ipv, xs, info


def _linpack_layout(rows, perm):
    """The LINPACK array and ipvt that the implicit-pivoting row image (row r never moves; perm[k] =
    pivot row of step k) stands for: LINPACK exchanges the active parts of two rows, so the multiplier
    of column k sits where its row was at step k, and row k of U is the active part of perm[k]."""
    n = len(perm)
    pos = list(range(n))
    A = np.zeros((n, n))
    ipvt = np.zeros(n, dtype=np.int32)
    for k in range(n - 1):
        l = int(perm[k])
        pl = pos[l]
        ipvt[k] = pl + 1
        for r in range(n):
            if pos[r] == k:
                pos[r] = pl
        pos[l] = k
        for r in range(n):
            if pos[r] > k:
                A[pos[r], k] = rows[r, k]
        A[k, k:] = rows[l, k:]
    l = int(perm[n - 1])
    A[n - 1, n - 1] = rows[l, n - 1]
    ipvt[n - 1] = n
    return A, ipvt


def _test_matrices(n, rng, nrand=24):
    mats = [rng.standard_normal((n, n)) for _ in range(nrand)]
    # diagonally dominant (no interchanges), like P = I - h*rl1*J at small h
    mats += [np.eye(n) + 0.01 * rng.standard_normal((n, n)) for _ in range(4)]
    # exact pivot ties: two (or all) rows of equal magnitude in a column -> first maximum (BL:334)
    t = rng.standard_normal((n, n))
    t[:, 0] = 1.0
    t[1::2, 0] = -1.0
    mats.append(t)
    t = rng.integers(-2, 3, size=(n, n)).astype(np.float64) + np.eye(n) * 0.0
    mats.append(t + 4.0 * np.eye(n)[::-1])  # small integers: exact arithmetic for a while, many ties
    t = np.ones((n, n)) + np.diag(np.arange(1.0, n + 1))
    mats.append(t)
    # singular: a zero column (info = that column), a repeated row (zero pivot late), all zeros
    s = rng.standard_normal((n, n))
    s[:, n // 2] = 0.0
    mats.append(s)
    s = rng.integers(-3, 4, size=(n, n)).astype(np.float64)
    s[n - 1] = s[0]
    mats.append(s)
    mats.append(np.zeros((n, n)))
    s = np.eye(n)
    s[n - 1, n - 1] = 0.0
    mats.append(s)  # a(n, n) == 0 -> info = n (LP:119)
    return np.array(mats)


@pytest.mark.parametrize("kind,n", [(0, 2), (0, 3), (0, 4), (1, 5), (1, 25), (1, 32), (2, 48), (2, 64)])
def test_dense_lu_known_answers_strict(kind, n):
    B = _B()
    rng = np.random.default_rng(100 * kind + n)
    a = _test_matrices(n, rng)
    b = rng.standard_normal((a.shape[0], n))
    fac, ipv, x, info = _oracle_dense(a, b)
    gfac, gipv, gx, ginfo = B.api.lu_test(kind, a, b, strict=True)
    assert np.array_equal(ginfo, info), (ginfo, info)
    assert (info != 0).sum() >= 3 and (info == 0).sum() >= 24  # (the repeated-row matrix is not exactly singular in floating point for large n)
    assert np.array_equal(gx, x), "solution differs from LINPACK's"
    for m in range(a.shape[0]):
        if kind == 0:
            A, ip = gfac[m], gipv[m]
        else:
            A, ip = _linpack_layout(gfac[m], gipv[m])
        assert np.array_equal(ip, ipv[m]), (m, ip, ipv[m])
        assert np.array_equal(A, fac[m]), m
    # the tie matrices really pivot on ties somewhere, and some matrix interchanges rows
    assert any((ipv[m] != np.arange(1, n + 1)).any() for m in range(a.shape[0]))


@pytest.mark.parametrize("kind,n", [(0, 3), (1, 32), (2, 48)])
def test_dense_lu_fast_build_solves(kind, n):
    """The fast build reorders / fuses operations: the solution agrees to rounding, info is the same."""
    B = _B()
    rng = np.random.default_rng(7 + n)
    a = _test_matrices(n, rng)
    b = rng.standard_normal((a.shape[0], n))
    _, _, x, info = _oracle_dense(a, b)
    _, _, gx, ginfo = B.api.lu_test(kind, a, b, strict=False)
    ok = info == 0
    # (matrix 32, the repeated row in small integers, is singular only if the elimination is exact: the fast
    # build's fused operations may leave a tiny pivot instead of a zero one)
    keep = np.arange(a.shape[0]) != 32
    assert np.array_equal((ginfo != 0)[keep], (info != 0)[keep])
    for m in np.nonzero(ok & keep)[0]:
        cond = np.linalg.cond(a[m])
        assert np.abs(gx[m] - x[m]).max() <= 1e-13 * cond * max(1.0, np.abs(x[m]).max()), (m, cond)


def _band_storage(full, ml, mu):
    n = full.shape[0]
    lda = 2 * ml + mu + 1
    abd = np.zeros((n, lda))  # abd[j, i] = LINPACK abd(i+1, j+1)
    m = ml + mu
    for j in range(n):
        for i in range(max(0, j - mu), min(n, j + ml + 1)):
            abd[j, i - j + m] = full[i, j]
    return abd


@pytest.mark.parametrize("ml,mu", [(5, 0), (5, 5), (1, 1), (0, 3)])
def test_band_lu_known_answers_strict(ml, mu):
    B = _B()
    n = 25
    rng = np.random.default_rng(10 * ml + mu)
    L = O.lib()
    mats, singular = [], 0
    for k in range(24):
        f = rng.standard_normal((n, n))
        if k % 3 == 0:
            f = f + 10.0 * np.eye(n)      # no interchanges
        if k % 3 == 1:
            f[np.arange(n), np.arange(n)] *= 0.01  # sub-diagonals outgrow the diagonal: interchanges
        if k == 20:
            f[:, 3] = 0.0                 # singular
        if k == 21:
            f = np.sign(f)                # ties everywhere
        mats.append(_band_storage(f, ml, mu))
    a = np.array(mats)
    b = rng.standard_normal((a.shape[0], n))
    lda = 2 * ml + mu + 1
    gfac, gipv, gx, ginfo = B.api.lu_test(3, a, b, ml=ml, mu=mu, strict=True)
    swaps = 0
    for m in range(a.shape[0]):
        abd = np.ascontiguousarray(a[m]).copy()  # [j][i] == column-major abd(lda, n)
        ip = np.zeros(n, dtype=np.int32)
        inf = C.c_int()
        L.dvo_dgbfa(abd.ctypes.data_as(C.POINTER(C.c_double)), lda, n, ml, mu, ip.ctypes.data_as(C.POINTER(C.c_int)), C.byref(inf))
        x = b[m].copy()
        if inf.value == 0:
            L.dvo_dgbsl(abd.ctypes.data_as(C.POINTER(C.c_double)), lda, n, ml, mu, ip.ctypes.data_as(C.POINTER(C.c_int)),
                        x.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            x[:] = 0.0
            singular += 1
        assert ginfo[m] == inf.value, (m, ginfo[m], inf.value)
        assert np.array_equal(gipv[m], ip), m
        assert np.array_equal(gfac[m], abd), m
        assert np.array_equal(gx[m], x), m
        swaps += int((ip != np.arange(1, n + 1)).sum())
    assert singular >= 1
    if ml > 0:
        assert swaps > 0


# ------------------------------------------------------------------ failure exits of the driver ----
@pytest.mark.parametrize("ml,mu", [(5, 0), (5, 5), (1, 1), (0, 3)])
def test_band_lu_fast_build_solves(ml, mu):
    """Fast build of the band policy: the band matrix is expanded to one dense row per lane and inverted with dgbfa's
    pivot sequence, the solve is a matrix-vector product (policy_warp.cuh, WarpBandPolicy::invert).  The solution
    agrees with LINPACK's to rounding (1e-13 * cond), info reports the same singular matrix, and the band array is
    left unfactored (it is what the solver saves with a system's state)."""
    B = _B()
    n = 25
    rng = np.random.default_rng(100 + 10 * ml + mu)
    L = O.lib()
    fulls, mats = [], []
    for k in range(24):
        f = rng.standard_normal((n, n))
        if k % 3 == 0:
            f = f + 10.0 * np.eye(n)
        if k % 3 == 1:
            f[np.arange(n), np.arange(n)] *= 0.01  # interchanges
        if k == 20:
            f[:, 3] = 0.0  # singular
        band = np.zeros((n, n))
        for j in range(n):
            lo, hi = max(0, j - mu), min(n, j + ml + 1)
            band[lo:hi, j] = f[lo:hi, j]
        fulls.append(band)
        mats.append(_band_storage(f, ml, mu))
    a = np.array(mats)
    b = rng.standard_normal((a.shape[0], n))
    lda = 2 * ml + mu + 1
    gfac, _, gx, ginfo = B.api.lu_test(3, a, b, ml=ml, mu=mu, strict=False)
    for m in range(a.shape[0]):
        abd = np.ascontiguousarray(a[m]).copy()
        ip = np.zeros(n, dtype=np.int32)
        inf = C.c_int()
        L.dvo_dgbfa(abd.ctypes.data_as(C.POINTER(C.c_double)), lda, n, ml, mu, ip.ctypes.data_as(C.POINTER(C.c_int)), C.byref(inf))
        assert (ginfo[m] != 0) == (inf.value != 0), (m, ginfo[m], inf.value)
        assert np.array_equal(gfac[m], a[m]), m  # unfactored
        if inf.value != 0:
            continue
        x = b[m].copy()
        L.dvo_dgbsl(abd.ctypes.data_as(C.POINTER(C.c_double)), lda, n, ml, mu, ip.ctypes.data_as(C.POINTER(C.c_int)),
                    x.ctypes.data_as(C.POINTER(C.c_double)))
        cond = np.linalg.cond(fulls[m])
        assert np.abs(gx[m] - x).max() <= 1e-13 * cond * max(1.0, np.abs(x).max()), (m, cond)


def _strict_vs_oracle(problem, par, y0, touts, rtol, atol, mf, itol, want, iopt=0, rw=None, iw=None, okw=None):
    """Run the strict build and the oracle (pow_mode = 1); everything must be identical; `want` is the set of
    istate values that must occur."""
    okw = okw or {}
    ref = O.batch_solve(problem, par, y0, 0.0, touts, rtol, atol, mf, itol=itol, pow_mode=1, iopt=iopt, **okw)
    got = gpu_schedule(problem, par, y0, 0.0, touts, rtol, atol, mf, itol, strict=True, iopt=iopt,
                       rwork_opts=rw, iwork_opts=iw)
    seen = set(ref["istate"][:, -1].tolist())
    assert set(want) <= seen, (want, seen)
    assert np.array_equal(got["istate"][:, -1], ref["istate"][:, -1])
    assert np.array_equal(got["y"], ref["y"])
    assert np.array_equal(got["t"][:, -1], ref["t"][:, -1])
    assert np.array_equal(got["istats"][:, -1], ref["istats"][:, -1])
    assert np.array_equal(got["rstats"][:, -1], ref["rstats"][:, -1])
    return got, ref


def _opts(h0=0.0, hmax=0.0, hmin=0.0, maxord=0, mxstep=0, ml=0, mu=0):
    rw = np.zeros(20)
    iw = np.zeros(30, dtype=np.int32)
    rw[4], rw[5], rw[6] = h0, hmax, hmin
    iw[0], iw[1], iw[4], iw[5] = ml, mu, maxord, mxstep
    return rw, iw


def test_istate_minus2_too_much_accuracy_mid_run():
    """tolsf > 1 after the start (M:1525-1543, 1636-1645): pure absolute tolerances near round-off."""
    B = _B()
    par = B.batches.robertson_batch(96)
    rw, iw = _opts(mxstep=20000)
    got, ref = _strict_vs_oracle("robertson", par, [1.0, 0.0, 0.0], B.batches.robertson_touts(), [0.0],
                                 [1e-15, 1e-15, 1e-17], 21, 2, {-2}, iopt=1, rw=rw, iw=iw, okw=dict(mxstep=20000))
    assert (ref["rstats"][:, -1, 3] > 1.0).all()  # tolsf is reported
    msg = got["solver"].message(0)
    assert "too much accuracy requested" in msg


def test_istate_minus4_error_test_at_hmin():
    B = _B()
    par = B.batches.robertson_batch(96)
    tol = B.batches.ROBERTSON_TOL_HEADLINE
    rw, iw = _opts(hmin=1e-3)
    got, ref = _strict_vs_oracle("robertson", par, [1.0, 0.0, 0.0], B.batches.robertson_touts(), tol["rtol"],
                                 tol["atol"], 21, 2, {-4}, iopt=1, rw=rw, iw=iw, okw=dict(hmin=1e-3))
    assert (ref["istats"][:, -1, 5] >= 1).all()  # imxer
    assert "error\n      test failed repeatedly" in got["solver"].message(0)


def test_istate_minus5_and_minus4_with_functional_iteration_at_hmin():
    B = _B()
    par = B.batches.robertson_batch(128)
    tol = B.batches.ROBERTSON_TOL_HEADLINE
    rw, iw = _opts(hmin=1e-3)
    _strict_vs_oracle("robertson", par, [1.0, 0.0, 0.0], B.batches.robertson_touts(), tol["rtol"], tol["atol"], 20, 2,
                      {-4, -5}, iopt=1, rw=rw, iw=iw, okw=dict(hmin=1e-3))


def test_repeated_failures_without_hmin_diagonal_approximation():
    """mf = 23 on the stiff batch: kfh = -7 consecutive error-test failures (M:2213-2216) and mxncf = 10
    convergence failures (M:2354-2366) end systems with istate -4 / -5 without any hmin."""
    B = _B()
    par = B.batches.robertson_batch(128)
    tol = B.batches.ROBERTSON_TOL_HEADLINE
    rw, iw = _opts(mxstep=100000)
    _strict_vs_oracle("robertson", par, [1.0, 0.0, 0.0], B.batches.robertson_touts(), tol["rtol"], tol["atol"], 23, 2,
                      {2, -5}, iopt=1, rw=rw, iw=iw, okw=dict(mxstep=100000))


def test_istate_minus6_weight_becomes_nonpositive_mid_run():
    B = _B()
    par = B.batches.perturbed((3.0,), 64)
    got, ref = _strict_vs_oracle("vdp_ewtdrop", par, [2.0, 0.0], [1.0, 2.0, 4.0, 8.0, 16.0], [1e-6], [1e-8], 21, 1, {-6})
    assert (ref["t"][:, -1] > 2.0).all() and (ref["t"][:, -1] < 8.0).all()  # mid-run, not at the start
    assert "ewt(i1) has become r2 <= 0" in got["solver"].message(0)


def test_singular_newton_matrix_thread_kernel():
    """Robertson at its steady state (0, 0, 1) with k1 = k2 = 1/4, integrated backwards with h0 = -2: the first
    Newton matrix I + 2J is [[.5, .5, 0], [.5, .5, 0], [0, 0, 1]] exactly -- a pivot tie, then a zero pivot
    (LP:87-88) -> ierpj -> the step is repeated with h/4 (M:2999, 2791, 2354-2366)."""
    B = _B()
    par = B.batches.robertson_batch(64)
    par[::2, 0] = 0.25
    par[::2, 1] = 0.25
    O.singular_events()
    rw, iw = _opts(h0=-2.0)
    tol = B.batches.ROBERTSON_TOL_REF
    for mf in (21, -21):
        got, ref = _strict_vs_oracle("robertson", par, [0.0, 0.0, 1.0], [-1.0, -4.0], tol["rtol"], tol["atol"], mf, 2,
                                     {2}, iopt=1, rw=rw, iw=iw, okw=dict(h0=-2.0))
        assert O.singular_events() == 32
        assert (ref["istats"][::2, -1, 10] >= 1).all() and (ref["istats"][1::2, -1, 10] == 0).all()  # ncfn


def test_singular_newton_matrix_band_kernel():
    """dvdemo problem 2 backwards with h0 = -1/2: P = I - h J has the diagonal 1 + 2h = 0 exactly."""
    B = _B()
    par = B.batches.advection_batch(48)
    O.singular_events()
    rw, iw = _opts(h0=-0.5, ml=5, mu=0)
    _strict_vs_oracle("advection", par, B.batches.advection_y0(), [-0.1, -0.2], [0.0], [1e-6], 24, 1, {2}, iopt=1,
                      rw=rw, iw=iw, okw=dict(h0=-0.5, ml=5, mu=0))
    assert O.singular_events() == 48


def test_singular_newton_matrix_dense_warp_kernel():
    """mech32 from the all-zero state (f = 0; J = its unimolecular part) with k_0 = 1/2 and h0 = -2."""
    B = _B()
    par = B.batches.mech32_batch(32)
    par[::2, 0] = 0.5
    O.singular_events()
    rw, iw = _opts(h0=-2.0, mxstep=5000)
    for mf in (21, -21):
        got, ref = _strict_vs_oracle("mech32", par, np.zeros(32), [-1.0, -3.0], [1e-6], [1e-12], mf, 1, {2}, iopt=1,
                                     rw=rw, iw=iw, okw=dict(h0=-2.0, mxstep=5000))
        assert O.singular_events() == 16
        assert (ref["istats"][::2, -1, 10] >= 1).all()


def test_schedule_longer_than_one_launch_keeps_failed_systems_state():
    """nout > BVODE_MAX_NOUT splits a schedule over several launches: systems that failed in the first launch keep
    their last state in the output slots of the later ones (they used to be left unwritten)."""
    B = _B()
    nsys = 64
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = 0.4 * 10.0 ** np.linspace(0.0, 5.0, 40)
    rw, iw = _opts(mxstep=60)  # some systems run out of steps between two close output points, some do not
    ref = O.batch_solve("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, tol["rtol"], tol["atol"], 21, itol=2,
                        pow_mode=1, iopt=1, mxstep=60)
    got = gpu_schedule("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, tol["rtol"], tol["atol"], 21, 2, strict=True,
                       iopt=1, rwork_opts=rw, iwork_opts=iw)
    first_fail = (ref["istate"] < 0).argmax(axis=1)
    assert ((ref["istate"][:, -1] == -1) & (first_fail < 16)).any(), "no system fails inside the first launch"
    assert np.array_equal(got["istate"][:, -1], ref["istate"][:, -1])
    assert np.array_equal(got["y"], ref["y"])


# ------------------------------------------------------------------ per-system tolerances / optional inputs ----
def _run_handle(B, problem, par, y0, touts, itol, rtol, atol, mf, iopt, rwork, iwork, per_system, per_call=False):
    nsys = par.shape[0]
    s = B.dvode_t().initialize(problem, nsys, par, strict=True, per_system=per_system)
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, s.neq))).copy()
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    yo = np.zeros((len(touts), nsys, s.neq))
    rw, iw = rwork.copy(), iwork.copy()
    if per_call:
        for k, to in enumerate(touts):
            s.solve(s.neq, y, t, to, itol, rtol, atol, 1, ist, iopt, rw, 20, iw, 30, mf)
            yo[k] = y
    else:
        s.solve_schedule(s.neq, y, t, np.asarray(touts), itol, rtol, atol, 1, ist, iopt, rw, 20, iw, 30, mf, yo)
    s.close()
    return yo, t, ist, rw[:, 10:14].copy(), iw[:, 10:22].copy()


@pytest.mark.parametrize("problem,mf", [("robertson", 21), ("mech32", 22), ("advection", 24)])
def test_per_system_tolerances_equal_separate_handles(problem, mf):
    """Two halves of a batch with different rtol / atol / maxord / mxstep / hmax in ONE handle
    (BVODE_FLAG_PER_SYSTEM) give exactly what two handles with batch-uniform inputs give."""
    B = _B()
    nsys = 64
    if problem == "robertson":
        par, y0, touts, itol = B.batches.robertson_batch(nsys), [1.0, 0.0, 0.0], B.batches.robertson_touts()[:8], 2
        tolA, tolB = ([1e-4], [1e-8, 1e-14, 1e-6]), ([1e-6], [1e-10, 1e-16, 1e-8])
        ml = mu = 0
    elif problem == "mech32":
        par, y0, touts, itol = B.batches.mech32_batch(nsys), B.batches.mech32_y0(), B.batches.mech32_touts()[:3], 1
        tolA, tolB = ([1e-6], [1e-12]), ([1e-4], [1e-10])
        ml = mu = 0
    else:
        par, y0, touts, itol = B.batches.advection_batch(nsys), B.batches.advection_y0(), B.batches.advection_touts(), 1
        tolA, tolB = ([0.0], [1e-6]), ([1e-5], [1e-8])
        ml, mu = 5, 0
    half = nsys // 2
    optA = dict(maxord=0, mxstep=100000, hmax=0.0)
    optB = dict(maxord=3, mxstep=50000, hmax=0.5 * float(touts[-1]))

    def headers(n, o):
        rw = np.zeros((n, 20))
        iw = np.zeros((n, 30), dtype=np.int32)
        rw[:, 5] = o["hmax"]
        iw[:, 0], iw[:, 1], iw[:, 4], iw[:, 5] = ml, mu, o["maxord"], o["mxstep"]
        return rw, iw

    rwA, iwA = headers(half, optA)
    rwB, iwB = headers(nsys - half, optB)
    a = _run_handle(B, problem, par[:half], y0, touts, itol, tolA[0], tolA[1], mf, 1, rwA, iwA, False)
    b = _run_handle(B, problem, par[half:], y0, touts, itol, tolB[0], tolB[1], mf, 1, rwB, iwB, False)
    nr, na = len(tolA[0]), len(tolA[1])
    rtol = np.concatenate([np.tile(tolA[0], (half, 1)), np.tile(tolB[0], (nsys - half, 1))]).reshape(nsys, nr)
    atol = np.concatenate([np.tile(tolA[1], (half, 1)), np.tile(tolB[1], (nsys - half, 1))]).reshape(nsys, na)
    both = _run_handle(B, problem, par, y0, touts, itol, rtol, atol, mf, 1, np.concatenate([rwA, rwB]),
                       np.concatenate([iwA, iwB]), True)
    for k in range(5):
        joined = np.concatenate([a[k], b[k]], axis=1 if k == 0 else 0)
        assert np.array_equal(both[k], joined), k
    assert (both[2] == 2).all()
    assert not np.array_equal(a[0][-1][: nsys - half], b[0][-1][:half])  # the halves really differ


def test_per_system_illegal_row_is_refused_alone():
    B = _B()
    nsys = 8
    par = B.batches.robertson_batch(nsys)
    s = B.dvode_t().initialize("robertson", nsys, par, strict=True, per_system=True)
    y = np.tile(np.array([1.0, 0.0, 0.0]), (nsys, 1))
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    rtol = np.full((nsys, 1), 1e-4)
    rtol[3, 0] = -1.0  # M:1283: rtol < 0 is illegal input for THAT instance
    atol = np.tile(np.array([1e-8, 1e-14, 1e-6]), (nsys, 1))
    s.solve(3, y, t, 0.4, 2, rtol, atol, 1, ist, 0, None, 0, None, 0, 21)
    assert ist[3] == -3 and (np.delete(ist, 3) == 2).all()
    assert t[3] == 0.0 and np.array_equal(y[3], [1.0, 0.0, 0.0])
    s.close()


# ------------------------------------------------------------------ multi-GPU entry points ----
def _gpus():
    import torch
    return torch.cuda.device_count()


def test_multi_handle_host_path_equals_single_handle():
    """bvode_multi_*: the batch sharded over device slots (the same GPU twice when only one is visible: the host
    path needs no inter-GPU traffic) gives the single-handle results, rows in place."""
    B = _B()
    nsys = 1001  # ragged blocks
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = B.batches.robertson_touts()
    ref = gpu_schedule("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, tol["rtol"], tol["atol"], 21, 2, strict=True)
    devs = list(range(_gpus())) if _gpus() > 1 else [0, 0]
    devs = (devs * 3)[:3] if len(devs) < 3 else devs[:3]
    m = B.dvode_multi_t().initialize("robertson", nsys, par, devices=devs, strict=True)
    y = np.tile(np.array([1.0, 0.0, 0.0]), (nsys, 1))
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20))
    iw = np.zeros((nsys, 30), dtype=np.int32)
    yo = np.zeros((touts.size, nsys, 3))
    m.solve_schedule(3, y, t, touts, 2, tol["rtol"], tol["atol"], 1, ist, 0, rw, 20, iw, 30, 21, yo)
    assert (ist == 2).all()
    assert np.array_equal(yo.transpose(1, 0, 2), ref["y"])
    assert np.array_equal(iw[:, 10:22], ref["istats"][:, -1])
    assert np.array_equal(rw[:, 10:14], ref["rstats"][:, -1])
    m.close()


def test_gather_device_single_rank_and_shard_ranges():
    import torch
    B = _B()
    for n, w in ((1000003, 8), (7, 3), (64, 2)):
        edges = [B.api.shard_range(n, r, w) for r in range(w)]
        assert edges == [B.sharding.shard_range(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
    nsys = 300
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    got = gpu_schedule("robertson", par, [1.0, 0.0, 0.0], 0.0, B.batches.robertson_touts(), tol["rtol"], tol["atol"],
                       21, 2, strict=True)
    s = got["solver"]
    comm = B.Comm(1, 0, 0)
    dev = torch.device("cuda", 0)
    y = torch.zeros(nsys, 3, dtype=torch.float64, device=dev)
    t = torch.zeros(nsys, dtype=torch.float64, device=dev)
    ist = torch.zeros(nsys, dtype=torch.int32, device=dev)
    ro = torch.zeros(nsys, 4, dtype=torch.float64, device=dev)
    io = torch.zeros(nsys, 12, dtype=torch.int32, device=dev)
    s.gather_device(comm, 0, nsys, y.data_ptr(), t.data_ptr(), ist.data_ptr(), ro.data_ptr(), io.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), got["y"][:, -1])
    assert np.array_equal(io.cpu().numpy(), got["istats"][:, -1])
    assert np.array_equal(ro.cpu().numpy(), got["rstats"][:, -1])
    assert (ist.cpu().numpy() == 2).all() and np.array_equal(t.cpu().numpy(), got["t"][:, -1])
    host = s.gather(comm, 0, nsys)
    assert np.array_equal(host["y"], got["y"][:, -1]) and np.array_equal(host["iout"], got["istats"][:, -1])
    comm.close()


@pytest.mark.skipif("_gpus() < 2")
def test_multi_gather_device_over_nccl():
    """One process, every visible GPU: solve sharded, then the final gather into GPU 0's memory over NCCL."""
    import torch
    B = _B()
    ng = min(_gpus(), 8)
    nsys = 4099
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = B.batches.robertson_touts()
    ref = gpu_schedule("robertson", par, [1.0, 0.0, 0.0], 0.0, touts, tol["rtol"], tol["atol"], 21, 2, strict=True)
    m = B.dvode_multi_t().initialize("robertson", nsys, par, ndev=ng, strict=True)
    y = np.tile(np.array([1.0, 0.0, 0.0]), (nsys, 1))
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    m.solve_schedule(3, y, t, touts, 2, tol["rtol"], tol["atol"], 1, ist, 0, None, 0, None, 0, 21, None)
    dev = torch.device("cuda", 0)
    ya = torch.zeros(nsys, 3, dtype=torch.float64, device=dev)
    ia = torch.zeros(nsys, 12, dtype=torch.int32, device=dev)
    sa = torch.zeros(nsys, dtype=torch.int32, device=dev)
    m.gather_device(0, y_all=ya.data_ptr(), istate_all=sa.data_ptr(), iout_all=ia.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(ya.cpu().numpy(), ref["y"][:, -1])
    assert np.array_equal(ia.cpu().numpy(), ref["istats"][:, -1])
    assert (sa.cpu().numpy() == 2).all()
    m.close()


# ------------------------------------------------------------------ dvsrco into a fresh handle ----
def test_dvsrco_restores_everything_a_fresh_handle_needs():
    B = _B()
    nsys = 128
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = B.batches.robertson_touts()
    rw0, iw0 = _opts(mxstep=80)  # some systems stop with istate = -1
    s = B.dvode_t().initialize("robertson", nsys, par, strict=True)
    y = np.tile(np.array([1.0, 0.0, 0.0]), (nsys, 1))
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20))
    iw = np.zeros((nsys, 30), dtype=np.int32)
    rw[0], iw[0] = rw0, iw0
    s.solve_schedule(3, y, t, touts[:6], 2, tol["rtol"], tol["atol"], 1, ist, 1, rw, 20, iw, 30, 21, None)
    assert (ist == -1).any() and (ist == 2).any()
    sav = s.dvsrco(None, 1)
    summ, msg = s.istate_summary(), s.message(int(np.argmax(ist == -1)))
    with pytest.raises(B.BvodeError):  # a short buffer is refused, not read past its end
        B.api._check(B.api._lib.bvode_dvsrco_n(s._h, B.api._ptr(sav), int(sav.nbytes) - 8, 1))
    s2 = B.dvode_t().initialize("robertson", nsys, par, strict=True)
    s2.dvsrco(sav, 2)
    assert s2.istate_summary() == summ and s2.message(int(np.argmax(ist == -1))) == msg
    assert "mxstep (=i1) steps" in msg and "        80" in msg
    # both handles continue the systems that are still fine, identically
    ok = ist == 2
    for h in (s, s2):
        y1, t1, i1 = y.copy(), t.copy(), np.where(ok, 2, 1).astype(np.int32)
        # (failed systems are restarted from their last state as new problems)
        rw1, iw1 = rw.copy(), iw.copy()
        iw1[0, 5] = 100000
        h.solve_schedule(3, y1, t1, touts[6:9], 2, tol["rtol"], tol["atol"], 1, i1, 1, rw1, 20, iw1, 30, 21, None)
        if h is s:
            keep = (y1.copy(), t1.copy(), i1.copy(), iw1[:, 10:22].copy())
        else:
            assert np.array_equal(y1, keep[0]) and np.array_equal(t1, keep[1]) and np.array_equal(i1, keep[2])
            assert np.array_equal(iw1[:, 10:22], keep[3])
    s.close()
    s2.close()


# ------------------------------------------------------------------ istate = 3 across method families ----
@pytest.mark.parametrize("problem,mf,mf2", [("vdp", 10, 21), ("vdp", 21, 10), ("vdp", 12, -22), ("vdp", 22, -12),
                                            ("vdp", 10, 20), ("vdp", 23, 13), ("advection", 14, -24),
                                            ("advection", 25, 10), ("mech32", 12, -22), ("mech32", 22, -12),
                                            ("diffusion48", 12, -22)])
def test_istate3_changes_the_method_family(problem, mf, mf2):
    """meth is re-read from mf on every call (M:1149-1153): an istate = 3 call may switch Adams <-> BDF.  The
    Nordsieck array carries over (13 <-> 6 columns); from Adams at order > 5 the order is cut to the BDF maximum
    through M:1385 / 2044-2079 -- orders 4..8 occur at the switch of the Van der Pol batch, so both the
    `dvjust(-1)` case (order 6) and the plain truncation (order >= 7) are on the path.  Strict build, bit for bit.
    (Where the old flag kept a Jacobian copy the new one is taken with jsv = -1: with jsv = +1 the reference's dvjac
    may reuse a "saved copy" from work-array words that hold the old Nordsieck columns after lwm moved, M:1246-1268 --
    not mirrored here, the Jacobian is re-evaluated at the restart instead; DESIGN.md.)"""
    B = _B()
    nsys = 48
    kw, iw = {}, None
    if problem == "vdp":
        par, y0 = B.batches.perturbed((3.0,), nsys), [2.0, 0.0]
        touts = [0.5, 1.0, 1.39283880203, 3.6076126770299997, 5.82238655203, 8.03716042703]
        rtol, atol, nsplit = [1e-8], [1e-10], 3
    elif problem == "advection":
        par, y0, touts = B.batches.advection_batch(nsys), B.batches.advection_y0(), B.batches.advection_touts()
        rtol, atol, nsplit = [0.0], [1e-6], 2
        kw = dict(ml=5, mu=0)
        iw = np.zeros(30, dtype=np.int32)
        iw[0] = 5
    elif problem == "mech32":
        par, y0, touts = B.batches.mech32_batch(nsys), B.batches.mech32_y0(), [1e-6, 1e-5, 1e-4, 1e-3]
        rtol, atol, nsplit = [1e-6], [1e-12], 2
        kw = dict(iopt=1, mxstep=50000, iopt2=1, mxstep2=50000)  # (Adams on the stiff mechanism needs its steps)
        iw = np.zeros(30, dtype=np.int32)
        iw[5] = 50000
    else:
        par, y0, touts = B.batches.diffusion48_batch(nsys), B.batches.diffusion48_y0(), [1e-4, 1e-3, 1e-2, 1e-1]
        rtol, atol, nsplit = B.batches.DIFFUSION48_TOL["rtol"], B.batches.DIFFUSION48_TOL["atol"], 2
    ref = O.batch_solve2(problem, par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1, mf2=mf2, **kw)
    iopt = kw.get("iopt", 0)
    got = gpu_schedule(problem, par, y0, 0.0, touts, rtol, atol, mf, 1, strict=True, per_call=True, iopt=iopt,
                       iwork_opts=iw, rwork_opts=np.zeros(20),
                       phase2=dict(nsplit=nsplit, rtol=rtol, atol=atol, itol=1, itask=1, iopt=iopt, mf=mf2,
                                   iwork_opts=iw, rwork_opts=np.zeros(20)))
    assert (ref["istate"] == 2).all()
    for k in ("y", "t", "istate", "istats", "rstats"):
        assert np.array_equal(got[k], ref[k]), k
    if problem == "vdp" and mf in (10, 12):
        nqu = ref["istats"][:, nsplit - 1, 3]
        assert (nqu == 6).any() and (nqu >= 7).any() and (nqu <= 5).any()


# ------------------------------------------------------------------ istate = 3 with a smaller neq ----
@pytest.mark.parametrize("problem,mf,neq2", [("advection", -24, 20), ("advection", -25, 20), ("advection", 10, 13),
                                             ("advection", 23, 20), ("advection", 13, 7), ("mech32", -22, 24),
                                             ("mech32", -21, 24), ("mech32", 13, 17),
                                             # two components per lane (32 < n <= 64): the cut inside the second slot,
                                             # at the slot boundary and inside the first slot.  System 17 of the
                                             # diagonal-approximation runs restarts with an order DECREASE pending: the
                                             # reference then clears the top Nordsieck column (M:2045-2053).
                                             ("diffusion48", -22, 40), ("diffusion48", -22, 32), ("diffusion48", 10, 20),
                                             ("diffusion48", 13, 33), ("diffusion48", 13, 47), ("diffusion48", 23, 40)])
def test_istate3_reduces_neq(problem, mf, neq2):
    """`neq` may be reduced (never raised) by an istate = 3 call (M:1127-1135): the solver goes on with the first
    neq equations; the dropped components stay what the caller's y array held for them (for mech32 the kept
    equations read them).  y, the output points and the work-array lengths follow the new neq.  Strict, bit for bit
    (jsv = -1 where the flag has a Jacobian: the reference's saved copy would change its leading dimension)."""
    B = _B()
    nsys = 40
    kw, iw = {}, np.zeros(30, dtype=np.int32)
    if problem == "advection":
        par, y0, touts = B.batches.advection_batch(nsys), B.batches.advection_y0(), B.batches.advection_touts()
        rtol, atol, nsplit = [0.0], [1e-6], 2
        kw = dict(ml=5, mu=0, iopt=1, mxstep=5000, iopt2=1, mxstep2=5000)
        iw[0], iw[5] = 5, 5000
    elif problem == "diffusion48":
        par, y0, touts = B.batches.diffusion48_batch(nsys), B.batches.diffusion48_y0(), list(B.batches.diffusion48_touts())
        rtol, atol, nsplit = [1e-6], [1e-9], 2
        if mf in (10, 13, 23):
            touts = touts[:3]
        kw = dict(iopt=1, mxstep=50000, iopt2=1, mxstep2=50000)
        iw[5] = 50000
    else:
        par, y0, touts = B.batches.mech32_batch(nsys), B.batches.mech32_y0(), [1e-6, 1e-5, 1e-4, 1e-3]
        rtol, atol, nsplit = [1e-6], [1e-12], 2
        kw = dict(iopt=1, mxstep=50000, iopt2=1, mxstep2=50000)
        iw[5] = 50000
    ref = O.batch_solve2(problem, par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1, neq2=neq2, **kw)
    s = B.dvode_t().initialize(problem, nsys, par, strict=True)
    n0 = s.neq
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, n0))).copy()
    t = np.zeros(nsys)
    ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20))
    iwk = np.zeros((nsys, 30), dtype=np.int32)
    iwk[0] = iw
    for k, to in enumerate(touts):
        neq = n0 if k < nsplit else neq2
        if k == nsplit:
            ist[:] = 3
            y = np.ascontiguousarray(y[:, :neq2])  # the caller's y now has neq2 entries per system
        s.solve(neq, y, t, to, 1, rtol, atol, 1, ist, 1, rw, 20, iwk, 30, mf)
        assert (ist == 2).all(), (k, ist)
        assert np.array_equal(y, ref["y"][:, k, :neq]), k
        assert np.array_equal(t, ref["t"][:, k]) and np.array_equal(iwk[:, 10:22], ref["istats"][:, k]), k
        assert np.array_equal(rw[:, 10:14], ref["rstats"][:, k]), k
    assert ref["istats"][0, -1, 6] < ref["istats"][0, 0, 6]  # lenrw follows neq (M:1246-1268)
    # a larger neq afterwards is illegal input; dvindy has the new width
    dky, iflag = s.dvindy(t, 0)
    assert dky.shape == (nsys, neq2) and (iflag == 0).all() and np.array_equal(dky, ref["y"][:, -1, :neq2])
    with pytest.raises(B.BvodeError):
        ist[:] = 3
        s.solve(n0, np.zeros((nsys, n0)), t, touts[-1] * 2, 1, rtol, atol, 1, ist, 1, rw, 20, iwk, 30, mf)
    s.close()


# ------------------------------------------------------------------ istate = 3 with another ml / mu ----
@pytest.mark.parametrize("mf,band,band2", [(-25, (5, 0), (5, 5)), (-25, (5, 5), (5, 0)), (-25, (5, 0), (7, 2)),
                                           (-25, (5, 0), (3, 0)), (-24, (5, 0), (5, 5)), (-24, (5, 5), (5, 0)),
                                           (25, (5, 0), (5, 5)), (24, (5, 5), (5, 0))])
def test_istate3_changes_the_bandwidth(mf, band, band2):
    """ml / mu are re-read from iwork(1:2) on every istate = 3 call (M:1136-1148 inside block B): the band image of
    the state takes the new shape, the Newton matrix is rebuilt at the restart, lenrw follows.  Widening, narrowing
    back, a band wider than the Jacobian's, and one that cuts entries off (ml = 3 < 5: a chord iteration with an
    incomplete matrix).  Strict build, bit for bit.  With jsv = +1 the reference's dvjac would read its "saved copy"
    with the new leading dimension from words written with the old one (M:1246-1268) -- not mirrored: the Jacobian is
    re-evaluated at the restart, so those runs are compared with the oracle's jsv = -1 run of the second phase."""
    B = _B()
    nsys = 40
    par, y0, touts = B.batches.advection_batch(nsys), B.batches.advection_y0(), B.batches.advection_touts()
    rtol, atol, nsplit = [0.0], [1e-6], 2
    (ml, mu), (ml2, mu2) = band, band2
    iw, iw2 = np.zeros(30, dtype=np.int32), np.zeros(30, dtype=np.int32)
    iw[0], iw[1], iw2[0], iw2[1] = ml, mu, ml2, mu2
    ref = O.batch_solve2("advection", par, y0, 0.0, touts, nsplit, rtol, atol, mf, rtol, atol, itol=1, pow_mode=1,
                         ml=ml, mu=mu, ml2=ml2, mu2=mu2, mf2=-abs(mf))
    assert (ref["istate"] == 2).all()
    got = gpu_schedule("advection", par, y0, 0.0, touts, rtol, atol, mf, 1, strict=True, per_call=True,
                       iwork_opts=iw, rwork_opts=np.zeros(20),
                       phase2=dict(nsplit=nsplit, rtol=rtol, atol=atol, itol=1, itask=1, iopt=0, mf=mf,
                                   iwork_opts=iw2, rwork_opts=np.zeros(20)))
    if mf < 0:
        for k in ("y", "t", "istate", "istats", "rstats"):
            assert np.array_equal(got[k], ref[k]), k
    else:  # a saved copy against none: not comparable call for call (see test_istate3_maxord_change_with_saved_jacobian)
        assert (got["istate"] == 2).all()
        assert np.array_equal(got["y"][:, :nsplit], ref["y"][:, :nsplit])
        err = np.abs(got["y"][:, nsplit:] - ref["y"][:, nsplit:]) / 1e-6
        assert np.percentile(err, 99) <= 10.0, float(np.percentile(err, 99))
        a = got["istats"][:, -1, O.ISTAT["nst"]].astype(np.int64).sum()
        b = ref["istats"][:, -1, O.ISTAT["nst"]].astype(np.int64).sum()
        assert abs(a - b) <= 0.03 * b, (a, b)
        assert (got["istats"][:, nsplit:, O.ISTAT["lenrw"]] == O.work_sizes(25, mf, ml2, mu2)[0]).all()
    assert ref["istats"][0, -1, O.ISTAT["lenrw"]] != ref["istats"][0, 0, O.ISTAT["lenrw"]]
