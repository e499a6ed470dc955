"""GPU tests at BASELINE.json's full batch sizes through size-independent properties, and the
edge cases small batches do not reach: row interchanges in the banded LU, ragged / tiny /
empty batches, determinism, independence of a system's result from its neighbours.

The oracle cannot run 2^20 systems in seconds, so these tests check what the domain offers:
  * Robertson (test/example.f90:79-109): y1 + y2 + y3 = 1 for all t (the three rates sum to 0);
  * dvdemo problem 2 (test/dvdemo.f90:230-301): the analytic solution `edit2` evaluates, which
    holds for ANY alph1 / alph2, so every perturbed system of the batch has its own exact answer;
  * the 32-species mechanism: every reaction conserves sum(y);
  * a system's trajectory depends on its own inputs only: permuting the batch permutes the
    results bit for bit, and a sub-batch reproduces the big batch's rows bit for bit.
"""
import math

import numpy as np
import pytest

import oracle_lib as O
from test_gpu_parity import _B, _iw, gpu_schedule

pytestmark = pytest.mark.gpu


def _schedule_device_resident(problem, par, y0, touts, tol, mf, iopt=0, iw=None, want_y_out=True):
    """Run one schedule launch with host buffers; returns (y_out[nout, nsys, neq], istate, iwork)."""
    B = _B()
    nsys = par.shape[0]
    s = B.dvode_t().initialize(problem, nsys, par)
    neq = s.neq
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, neq))).copy()
    t = np.zeros(nsys)
    istate = np.ones(nsys, dtype=np.int32)
    rwork = np.zeros((nsys, 20))
    iwork = np.zeros((nsys, 30), dtype=np.int32)
    if iw is not None:
        iwork[0, :] = iw
    yo = np.zeros((len(touts), nsys, neq)) if want_y_out else None
    s.solve_schedule(neq, y, t, np.asarray(touts, dtype=np.float64), tol["itol"], tol["rtol"], tol["atol"],
                     1, istate, iopt, rwork, 20, iwork, 30, mf, yo)
    s.close()
    return yo, y, t, istate, iwork


# ------------------------------------------------------------------ full batch sizes ------
def test_robertson_full_size_conservation_and_counters():
    """BASELINE.json configs[1]: 2^20 perturbed Robertson systems, rtol 1e-6, 12 decades."""
    B = _B()
    nsys = 1 << 20
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_HEADLINE
    touts = B.batches.robertson_touts()
    yo, y, t, istate, iwork = _schedule_device_resident("robertson", par, B.batches.ROBERTSON_Y0, touts, tol, 21)
    assert (istate == 2).all()
    assert np.array_equal(t, np.full(nsys, touts[-1]))
    assert np.array_equal(yo[-1], y)
    # mass conservation: the error of the sum is bounded by the local tolerances accumulated over
    # ~1100 steps; the reference's own run (example.f90, rtol 1e-4) conserves to ~1e-6
    assert np.abs(yo.sum(axis=2) - 1.0).max() < 2e-6
    assert yo.min() > -1e-9
    # y1 decays monotonically decade to decade, y3 grows
    assert (np.diff(yo[:, :, 0], axis=0) < 0).all() and (np.diff(yo[:, :, 2], axis=0) > 0).all()
    # counters: the batch means must be those of the oracle on a sample of the same batch (+-2 %)
    ns = 2048
    ref = O.batch_solve("robertson", par[:ns], B.batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"], tol["atol"],
                        21, itol=2)
    for name in ("nst", "nfe", "nje", "nlu", "nni"):
        a = iwork[:, 10 + O.ISTAT[name]].astype(np.float64).mean()
        b = ref["istats"][:, -1, O.ISTAT[name]].astype(np.float64).mean()
        assert abs(a - b) <= 0.02 * b, (name, a, b)


def _advection_exact(par, t):
    """edit2 of test/dvdemo.f90:276-301 for per-system alph1, alph2: y[nsys, 25]."""
    out = np.zeros((par.shape[0], 25))
    ex = math.exp(-2.0 * t) if t <= 30.0 else 0.0
    for j in range(1, 6):
        for i in range(1, 6):
            k = (i - 1) + (j - 1) * 5
            out[:, k] = (t ** (i + j - 2) * ex * par[:, 0] ** (i - 1) / math.factorial(i - 1)
                         * par[:, 1] ** (j - 1) / math.factorial(j - 1))
    return out


@pytest.mark.parametrize("mf", [24, 25])
def test_advection_full_size_against_analytic_solution(mf):
    """BASELINE.json configs[2]: 2^18 banded systems; every system against its own exact solution.
    The reference's run has max errors 0.56e-6 .. 5.3e-6 at atol 1e-6 (test/dvdemo.txt:481-485)."""
    B = _B()
    nsys = 1 << 18
    par = B.batches.advection_batch(nsys)
    touts = B.batches.advection_touts()
    yo, _, _, istate, iwork = _schedule_device_resident("advection", par, B.batches.advection_y0(), touts,
                                                        B.batches.ADVECTION_TOL, mf, iw=_iw(5, 0))
    assert (istate == 2).all()
    for k, t in enumerate(touts):
        err = np.abs(yo[k] - _advection_exact(par, float(t))).max()
        assert err < 2e-5, (t, err)
    nst = iwork[:, 10].astype(np.float64)
    assert 100 < nst.mean() < 140 and nst.max() < 200  # the reference takes 119 / 120 steps


def test_mech32_full_size_mass_conservation():
    """BASELINE.json configs[3]: 2^18 systems of the 32-species mechanism, dense FD Jacobian."""
    B = _B()
    nsys = 1 << 18
    par = B.batches.mech32_batch(nsys)
    yo, _, _, istate, iwork = _schedule_device_resident("mech32", par, B.batches.mech32_y0(), B.batches.mech32_touts(),
                                                        B.batches.MECH32_TOL, 22, iopt=1, iw=_iw(0, 0, 100000))
    assert (istate == 2).all()
    assert np.abs(yo.sum(axis=2) - 1.0).max() < 1e-5
    assert yo.min() > -1e-7
    ns = 256
    ref = O.batch_solve("mech32", par[:ns], B.batches.mech32_y0(), 0.0, B.batches.mech32_touts(), [1e-6], [1e-12],
                        22, itol=1, iopt=1, mxstep=100000)
    for name in ("nst", "nfe", "nje", "nlu", "nni"):
        a = iwork[:, 10 + O.ISTAT[name]].astype(np.float64).mean()
        b = ref["istats"][:, -1, O.ISTAT[name]].astype(np.float64).mean()
        assert abs(a - b) <= 0.03 * b, (name, a, b)


# -------------------------------------------------- independence of the systems of a batch ------
@pytest.mark.parametrize("problem", ["robertson", "advection", "mech32"])
def test_permutation_and_subbatch_invariance(problem):
    """Results are a function of a system's own inputs: a permuted batch gives the permuted
    results, a sub-batch (different block / warp assignment, ragged tail) the same rows, and a
    second run the same bits."""
    B = _B()
    if problem == "robertson":
        nsys, mf, iopt, iw = 4096 + 37, 21, 0, None
        par, y0, touts, tol = B.batches.robertson_batch(nsys), B.batches.ROBERTSON_Y0, B.batches.robertson_touts(), B.batches.ROBERTSON_TOL_HEADLINE
    elif problem == "advection":
        nsys, mf, iopt, iw = 1024 + 5, 25, 0, _iw(5, 0)
        par, y0, touts, tol = B.batches.advection_batch(nsys), B.batches.advection_y0(), B.batches.advection_touts(), B.batches.ADVECTION_TOL
    else:
        nsys, mf, iopt, iw = 256 + 3, 22, 1, _iw(0, 0, 100000)
        par, y0, touts, tol = B.batches.mech32_batch(nsys), B.batches.mech32_y0(), B.batches.mech32_touts(), B.batches.MECH32_TOL
    run = lambda p: _schedule_device_resident(problem, p, y0, touts, tol, mf, iopt=iopt, iw=iw)
    yo, y, t, ist, iwk = run(par)
    assert (ist == 2).all()
    yo2, _, _, _, iwk2 = run(par)
    assert np.array_equal(yo, yo2) and np.array_equal(iwk[:, 10:22], iwk2[:, 10:22])
    perm = np.random.default_rng(7).permutation(nsys)
    yop, _, _, _, iwkp = run(np.ascontiguousarray(par[perm]))
    assert np.array_equal(yop, yo[:, perm])
    assert np.array_equal(iwkp[:, 10:22], iwk[perm, 10:22])
    sub = slice(33, 33 + 71)
    yos, _, _, _, iwks = run(np.ascontiguousarray(par[sub]))
    assert np.array_equal(yos, yo[:, sub])
    assert np.array_equal(iwks[:, 10:22], iwk[sub, 10:22])


@pytest.mark.parametrize("nsys", [1, 2, 31, 33, 129])
def test_tiny_and_ragged_batches_match_the_oracle(nsys):
    B = _B()
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = B.batches.robertson_touts()
    ref = O.batch_solve("robertson", par, B.batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"], tol["atol"], 21,
                        itol=2, pow_mode=1)
    got = gpu_schedule("robertson", par, B.batches.ROBERTSON_Y0, 0.0, touts, tol["rtol"], tol["atol"], 21, 2,
                       strict=True)
    assert np.array_equal(got["y"], ref["y"])
    assert np.array_equal(got["istats"][:, -1], ref["istats"][:, -1])
    par = B.batches.mech32_batch(nsys)
    ref = O.batch_solve("mech32", par, B.batches.mech32_y0(), 0.0, B.batches.mech32_touts()[:2], [1e-6], [1e-12], 22,
                        itol=1, iopt=1, mxstep=100000, pow_mode=1)
    got = gpu_schedule("mech32", par, B.batches.mech32_y0(), 0.0, B.batches.mech32_touts()[:2], [1e-6], [1e-12], 22,
                       1, strict=True, iopt=1, iwork_opts=_iw(0, 0, 100000))
    assert np.array_equal(got["y"], ref["y"])
    assert np.array_equal(got["istats"][:, -1], ref["istats"][:, -1])


# ------------------------------------------------------- row interchanges in the banded LU ------
def _strong_coupling_batch(nsys):
    """alph1, alph2 in [3, 5): the sub-diagonals of P = I - h*rl1*J outgrow its diagonal
    (|h*rl1*alph| > 1 + 2*h*rl1 once h*rl1 > 1), so dgbfa interchanges rows and fills the ml extra
    rows of the band array (LP:318-361) -- the advection batch with alph ~ 1 never does."""
    B = _B()
    return 4.0 + B.batches.uniform_pm1(nsys, 2, B.batches.SEED, 0)


@pytest.mark.parametrize("mf,mu", [(24, 0), (25, 0), (-24, 0), (25, 5)])
def test_band_pivoting_strict_bit_exact(mf, mu):
    B = _B()
    nsys = 64
    par = _strong_coupling_batch(nsys)
    tol = B.batches.ADVECTION_TOL
    touts = B.batches.advection_touts()
    ref = O.batch_solve("advection", par, B.batches.advection_y0(), 0.0, touts, tol["rtol"], tol["atol"], mf,
                        itol=1, ml=5, mu=mu, pow_mode=1)
    got = gpu_schedule("advection", par, B.batches.advection_y0(), 0.0, touts, tol["rtol"], tol["atol"], mf, 1,
                       strict=True, iwork_opts=_iw(5, mu))
    assert (ref["istate"][:, -1] == 2).all() and (got["istate"][:, -1] == 2).all()
    assert np.array_equal(got["y"], ref["y"])
    assert np.array_equal(got["istats"][:, -1], ref["istats"][:, -1])
    assert np.array_equal(got["rstats"][:, -1], ref["rstats"][:, -1])


def test_band_pivoting_fast_build_against_analytic_solution():
    B = _B()
    nsys = 4096
    par = _strong_coupling_batch(nsys)
    touts = B.batches.advection_touts()[:4]  # the exact solution reaches 1e3 at alph ~ 4: relative bar
    for mf in (24, 25):
        yo, _, _, istate, iwork = _schedule_device_resident("advection", par, B.batches.advection_y0(), touts,
                                                            dict(itol=1, rtol=[1e-6], atol=[1e-8]), mf, iw=_iw(5, 0))
        assert (istate == 2).all()
        ref = O.batch_solve("advection", par[:128], B.batches.advection_y0(), 0.0, touts, [1e-6], [1e-8], mf, itol=1,
                            ml=5, mu=0)
        for k, t in enumerate(touts):
            ex = _advection_exact(par, float(t))
            rel = np.abs(yo[k] - ex) / (np.abs(ex) + 1e-3)
            eref = (np.abs(ref["y"][:, k] - ex[:128]) / (np.abs(ex[:128]) + 1e-3)).max()
            # the same 128 systems: the accuracy of the reference's arithmetic; all 4096: global
            # error a few hundred rtol at most
            assert rel[:128].max() < 10.0 * eref + 1e-5, (mf, t, rel[:128].max(), eref)
            assert rel.max() < 1e-3, (mf, t, rel.max())
        for name in ("nst", "nfe", "nlu", "nni"):
            a = iwork[:128, 10 + O.ISTAT[name]].astype(np.int64).sum()
            b = ref["istats"][:, -1, O.ISTAT[name]].astype(np.int64).sum()
            assert abs(a - b) <= 0.03 * b, (mf, name, a, b)


def test_empty_batch_is_refused_at_create():
    """nsys = 0 has no reference counterpart (one dvode_t is one system); bvode_create answers
    BVODE_EINVAL and the Python mirror raises instead of returning a handle."""
    B = _B()
    with pytest.raises(Exception):
        B.dvode_t().initialize("robertson", 0, np.zeros((0, 3)))


# ------------------------------------------------------- compaction of finished systems ------
@pytest.mark.parametrize("strict", [False, True])
def test_compaction_gives_identical_results(strict):
    """BVODE_FLAG_COMPACT: the grid holds only the threads resident at once and a thread whose
    system is finished takes the next unclaimed one.  Which thread integrates which system must
    not change a single bit (strict build): same batch (larger than the resident capacity, ragged,
    with a heavy tail through per-system tout) with and without the flag."""
    B = _B()
    nsys = (1 << 17) + 77
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_HEADLINE
    tps = np.full(nsys, 40.0)
    tps[::61] = 4.0e6
    out = []
    for compact in (False, True):
        s = B.dvode_t().initialize("robertson", nsys, par, strict=strict, compact=compact)
        y = np.tile(np.array(B.batches.ROBERTSON_Y0), (nsys, 1))
        t = np.zeros(nsys)
        ist = np.ones(nsys, dtype=np.int32)
        rw = np.zeros((nsys, 20))
        iw = np.zeros((nsys, 30), dtype=np.int32)
        iw[0, 5] = 100000
        s.solve(3, y, t, tps, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 1, rw, 20, iw, 30, 21)
        assert (ist == 2).all()
        # a continuation call (istate = 2: the saved state is reloaded) one decade on
        s.solve(3, y, t, tps * 10.0, tol["itol"], tol["rtol"], tol["atol"], 1, ist, 1, rw, 20, iw, 30, 21)
        assert (ist == 2).all()
        out.append((y.copy(), t.copy(), rw[:, 10:14].copy(), iw[:, 10:22].copy()))
        s.close()
    if strict:
        for a, b in zip(out[0], out[1]):
            assert np.array_equal(a, b)
    else:
        # The compacting kernel is a separate instantiation of the same source; in the fast build the
        # compiler is free to contract a*b+c differently in it (round 2: one DADD of the default kernel is a
        # DFMA there), so the two runs are two round-off-different VODE runs: same statistics, values
        # inside the tolerance-weighted spread.  (The strict build, above, is the bit-for-bit statement.)
        ya, yb = out[0][0], out[1][0]
        w = np.abs(ya - yb) / (tol["rtol"][0] * np.abs(ya) + np.array(tol["atol"]))
        assert np.median(w.max(axis=1)) < 5.0 and (w.max(axis=1) <= 10.0).mean() > 0.8, (np.median(w), w.max())
        assert np.array_equal(out[0][1], out[1][1])  # t
        ca, cb = out[0][3].astype(np.int64).sum(axis=0), out[1][3].astype(np.int64).sum(axis=0)
        for k in (0, 1, 2, 8, 9):  # nst, nfe, nje, nlu, nni
            assert abs(ca[k] - cb[k]) <= 0.002 * ca[k], (k, ca[k], cb[k])
    # and against the oracle on a sample (strict build: bit for bit)
    if strict:
        idx = np.arange(0, nsys, 997)
        ref = O.batch_solve("robertson", par[idx], B.batches.ROBERTSON_Y0, 0.0, [40.0, 400.0], tol["rtol"],
                            tol["atol"], 21, itol=2, iopt=1, mxstep=100000, pow_mode=1)
        short = tps[idx] == 40.0
        assert np.array_equal(out[1][0][idx][short], ref["y"][short, 1])


def test_device_pointer_call_with_per_system_tout():
    """bvode_solve_device (inputs in HBM, tout per system) = bvode_solve with tout_per_system = 1."""
    import torch
    B = _B()
    nsys = 1000
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    tps = 0.4 * 10.0 ** (np.arange(nsys) % 5)
    s = B.dvode_t().initialize("robertson", nsys, par)
    y = np.tile(np.array(B.batches.ROBERTSON_Y0), (nsys, 1)); t = np.zeros(nsys); ist = np.ones(nsys, dtype=np.int32)
    rw = np.zeros((nsys, 20)); iw = np.zeros((nsys, 30), dtype=np.int32)
    s.solve(3, y, t, tps, 2, tol["rtol"], tol["atol"], 1, ist, 0, rw, 20, iw, 30, 21)
    s.close()
    dev = torch.device("cuda", 0)
    s = B.dvode_t().initialize("robertson", nsys, par)
    yd = torch.tensor(np.tile(np.array(B.batches.ROBERTSON_Y0)[:, None], (1, nsys)), dtype=torch.float64, device=dev)
    td = torch.zeros(nsys, dtype=torch.float64, device=dev)
    isd = torch.ones(nsys, dtype=torch.int32, device=dev)
    tod = torch.tensor(tps, dtype=torch.float64, device=dev)
    s.solve_device(3, yd.data_ptr(), td.data_ptr(), tod.data_ptr(), 2, tol["rtol"], tol["atol"], 1, isd.data_ptr(), 0,
                   np.zeros(20), np.zeros(30, dtype=np.int32), 21)
    torch.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy().T, y)
    assert np.array_equal(td.cpu().numpy(), t) and np.array_equal(t, tps)
    assert np.array_equal(isd.cpu().numpy(), ist) and (ist == 2).all()
    s.close()


def test_pipelined_host_path_ragged_chunks():
    """Batches of 2^18 systems and more go through the host-buffer entry points in chunks of 2^17 on
    two streams (copies under kernels).  A batch that does not divide evenly: every system's rows are
    those of a small batch run on its own, for the schedule call and for per-system tout."""
    B = _B()
    nsys = (1 << 18) + 333
    par = B.batches.robertson_batch(nsys)
    tol = B.batches.ROBERTSON_TOL_REF
    touts = B.batches.robertson_touts()[:4]
    yo, y, t, ist, iw = _schedule_device_resident("robertson", par, B.batches.ROBERTSON_Y0, touts, tol, 21)
    assert (ist == 2).all() and np.array_equal(t, np.full(nsys, touts[-1]))
    assert np.abs(yo.sum(axis=2) - 1.0).max() < 1e-5
    for lo in (0, (1 << 17) - 7, nsys - 300):  # inside the first chunk, across a boundary, the ragged tail
        sl = slice(lo, lo + 300)
        yo_s, y_s, _, _, iw_s = _schedule_device_resident("robertson", np.ascontiguousarray(par[sl]),
                                                          B.batches.ROBERTSON_Y0, touts, tol, 21)
        assert np.array_equal(yo_s, yo[:, sl]) and np.array_equal(y_s, y[sl])
        assert np.array_equal(iw_s[:, 10:22], iw[sl, 10:22])
    # one call with a tout per system
    s = B.dvode_t().initialize("robertson", nsys, par)
    y2 = np.tile(np.array(B.batches.ROBERTSON_Y0), (nsys, 1)); t2 = np.zeros(nsys); ist2 = np.ones(nsys, dtype=np.int32)
    tps = np.where(np.arange(nsys) % 3 == 0, touts[1], touts[3])
    s.solve(3, y2, t2, tps, tol["itol"], tol["rtol"], tol["atol"], 1, ist2, 0, None, 0, None, 0, 21)
    s.close()
    assert (ist2 == 2).all() and np.array_equal(t2, tps)
    for lo in ((1 << 17) - 7, nsys - 300):
        sl = slice(lo, lo + 300)
        s = B.dvode_t().initialize("robertson", 300, np.ascontiguousarray(par[sl]))
        y3 = np.tile(np.array(B.batches.ROBERTSON_Y0), (300, 1)); t3 = np.zeros(300); ist3 = np.ones(300, dtype=np.int32)
        s.solve(3, y3, t3, tps[sl], tol["itol"], tol["rtol"], tol["atol"], 1, ist3, 0, None, 0, None, 0, 21)
        s.close()
        assert np.array_equal(y3, y2[sl]) and np.array_equal(t3, t2[sl])
