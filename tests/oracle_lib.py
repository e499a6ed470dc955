"""ctypes wrapper around the CPU oracle (oracle/_build/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(dvode_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

PROBLEMS = {"robertson": 0, "vdp": 1, "advection": 2, "mech32": 3, "vdp_maxnorm": 4,
            "advection_maxnorm": 5, "diffusion48": 6, "vdp_ewtdrop": 7}


class BatchCfg(C.Structure):
    _fields_ = [("problem", C.c_int), ("mf", C.c_int), ("itol", C.c_int), ("itask", C.c_int),
                ("iopt", C.c_int), ("ml", C.c_int), ("mu", C.c_int), ("maxord", C.c_int),
                ("mxstep", C.c_int), ("mxhnil", C.c_int), ("h0", C.c_double),
                ("hmax", C.c_double), ("hmin", C.c_double), ("tcrit", C.c_double),
                ("pow_mode", C.c_int)]


class BatchCfg32(C.Structure):  # liboracle32.so: every `real` a float
    _fields_ = [("problem", C.c_int), ("mf", C.c_int), ("itol", C.c_int), ("itask", C.c_int),
                ("iopt", C.c_int), ("ml", C.c_int), ("mu", C.c_int), ("maxord", C.c_int),
                ("mxstep", C.c_int), ("mxhnil", C.c_int), ("h0", C.c_float),
                ("hmax", C.c_float), ("hmin", C.c_float), ("tcrit", C.c_float),
                ("pow_mode", C.c_int)]


def build(force=False):
    srcs = [os.path.join(_ROOT, "oracle", f) for f in
            ("dvode_oracle.c", "problems.c", "dvode_oracle.h", "oracle_batch.h", "Makefile")]
    srcs.append(os.path.join(_ROOT, "include", "bvode_mech32.h"))
    fma = _SO.replace("liboracle.so", "liboracle_fma.so")
    r32 = _SO.replace("liboracle.so", "liboracle32.so")
    stale = force or not all(os.path.exists(x) for x in (_SO, fma, r32)) or any(
        os.path.getmtime(s) > min(os.path.getmtime(_SO), os.path.getmtime(fma), os.path.getmtime(r32))
        for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")],
                              stdout=subprocess.DEVNULL)
    return _SO


_libs = {}


def lib(variant=""):
    """variant "" = the parity oracle (no FMA); "fma" = the FMA-contracted sensitivity yardstick."""
    if variant not in _libs:
        path = build()
        if variant:
            path = path.replace("liboracle.so", "liboracle32.so" if variant == "32" else "liboracle_%s.so" % variant)
        _lib = C.CDLL(path)
        if variant == "32":  # only the batch driver is used on the REAL32 twin
            fp = C.POINTER(C.c_float)
            ip = C.POINTER(C.c_int)
            _lib.dvo_batch_solve.restype = C.c_int
            _lib.dvo_batch_solve.argtypes = [C.POINTER(BatchCfg32), C.c_int, fp, fp, C.c_float,
                                             C.c_int, fp, fp, fp, fp, fp, ip, fp, ip, C.c_int]
            _lib.dvo_problem_info.argtypes = [C.c_int, ip, ip]
            _libs[variant] = _lib
            return _lib
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        _lib.dvo_batch_solve.restype = C.c_int
        _lib.dvo_batch_solve.argtypes = [C.POINTER(BatchCfg), C.c_int, dp, dp, C.c_double,
                                         C.c_int, dp, dp, dp, dp, dp, ip, dp, ip, C.c_int]
        _lib.dvo_single_dense.restype = C.c_int
        _lib.dvo_single_dense.argtypes = [C.POINTER(BatchCfg), dp, dp, C.c_double, C.c_double,
                                          dp, dp, C.c_int, C.c_int, dp, dp, ip, dp, dp]
        _lib.dvo_problem_info.argtypes = [C.c_int, ip, ip]
        _lib.dvo_problem_f.argtypes = [C.c_int, dp, C.c_double, dp, dp]
        _lib.dvo_work_sizes.argtypes = [C.c_int] * 5 + [ip, ip]
        _lib.dvo_powi.restype = C.c_double
        _lib.dvo_powi.argtypes = [C.c_double, C.c_int]
        _lib.dvo_ppow.restype = C.c_double
        _lib.dvo_ppow.argtypes = [C.c_double, C.c_double]
        for name in ("dvo_dgefa", "dvo_dgesl", "dvo_dgbfa", "dvo_dgbsl"):
            getattr(_lib, name).restype = None
        _lib.dvo_dgefa.argtypes = [dp, C.c_int, C.c_int, ip, ip]
        _lib.dvo_dgesl.argtypes = [dp, C.c_int, C.c_int, ip, dp]
        _lib.dvo_dgbfa.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        _lib.dvo_dgbsl.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, dp]
        _libs[variant] = _lib
    return _libs[variant]


def singular_events(reset=True, variant=""):
    """Singular Newton matrices (dvjac's ierpj != 0, M:2999) seen since the last reset."""
    L = lib(variant)
    L.dvo_singular_events.restype = C.c_int
    L.dvo_singular_events.argtypes = [C.c_int]
    return int(L.dvo_singular_events(1 if reset else 0))


def threads_used(variant=""):
    """Threads the last batch_solve call of this library actually ran on (omp_get_num_threads)."""
    L = lib(variant)
    L.dvo_last_threads.restype = C.c_int
    return int(L.dvo_last_threads())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def problem_info(problem):
    neq, npar = C.c_int(), C.c_int()
    assert lib().dvo_problem_info(PROBLEMS[problem], C.byref(neq), C.byref(npar)) == 0
    return neq.value, npar.value


def work_sizes(neq, mf, ml=0, mu=0, maxord=0):
    lrw, liw = C.c_int(), C.c_int()
    lib().dvo_work_sizes(neq, mf, ml, mu, maxord, C.byref(lrw), C.byref(liw))
    return lrw.value, liw.value


def batch_solve(problem, params, y0, t0, touts, rtol, atol, mf, itol=1, itask=1, iopt=0,
                ml=0, mu=0, maxord=0, mxstep=0, mxhnil=0, h0=0.0, hmax=0.0, hmin=0.0,
                tcrit=0.0, pow_mode=0, nthreads=0, variant=""):
    """Run the oracle on a batch.  Returns dict(y[nsys,nout,neq], t, istate, rstats, istats).
    variant "32": the REAL32 twin (float arrays in and out)."""
    ft = np.float32 if variant == "32" else np.float64
    fptr = (lambda a: a.ctypes.data_as(C.POINTER(C.c_float))) if variant == "32" else _dp
    neq, npar = problem_info(problem)
    params = np.ascontiguousarray(np.asarray(params, dtype=ft).reshape(-1, max(npar, 1)))
    nsys = params.shape[0]
    y0 = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=ft), (nsys, neq)))
    touts = np.ascontiguousarray(np.asarray(touts, dtype=ft).reshape(-1))
    nout = touts.size
    rtol = np.ascontiguousarray(np.atleast_1d(np.asarray(rtol, dtype=ft)))
    atol = np.ascontiguousarray(np.atleast_1d(np.asarray(atol, dtype=ft)))
    Cfg = BatchCfg32 if variant == "32" else BatchCfg
    cfg = Cfg(PROBLEMS[problem], mf, itol, itask, iopt, ml, mu, maxord, mxstep, mxhnil,
              h0, hmax, hmin, tcrit, pow_mode)
    y = np.zeros((nsys, nout, neq), dtype=ft)
    t = np.zeros((nsys, nout), dtype=ft)
    ist = np.zeros((nsys, nout), dtype=np.int32)
    rst = np.zeros((nsys, nout, 4), dtype=ft)
    ista = np.zeros((nsys, nout, 12), dtype=np.int32)
    nfail = lib(variant).dvo_batch_solve(C.byref(cfg), nsys, fptr(params), fptr(y0), float(t0), nout,
                                         fptr(touts), fptr(rtol), fptr(atol), fptr(y), fptr(t), _ip(ist),
                                         fptr(rst), _ip(ista), nthreads)
    assert nfail >= 0
    return {"y": y, "t": t, "istate": ist, "rstats": rst, "istats": ista, "nfail": nfail}


def batch_solve2(problem, params, y0, t0, touts, nsplit, rtol, atol, mf, rtol2, atol2, itol=1,
                 itol2=None, itask=1, itask2=None, iopt=0, iopt2=None, ml=0, mu=0, maxord=0,
                 maxord2=0, mxstep=0, mxstep2=0, hmax=0.0, hmax2=0.0, hmin=0.0, hmin2=0.0,
                 tcrit=0.0, tcrit2=0.0, pow_mode=0, nthreads=0, variant="", mf2=None, neq2=0, ml2=None, mu2=None):
    """Outputs [0, nsplit) with the first parameter set, then an istate = 3 call with the second
    set (and method flag mf2, default the same) and continuation calls: the reference's
    parameter-change restart."""
    neq, npar = problem_info(problem)
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1, max(npar, 1)))
    nsys = params.shape[0]
    y0 = np.ascontiguousarray(np.broadcast_to(np.asarray(y0, dtype=np.float64), (nsys, neq)))
    touts = np.ascontiguousarray(np.asarray(touts, dtype=np.float64).reshape(-1))
    nout = touts.size
    arr = lambda a: np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
    rtol, atol, rtol2, atol2 = arr(rtol), arr(atol), arr(rtol2), arr(atol2)
    pick = lambda a, b: b if b is not None else a
    cfg = BatchCfg(PROBLEMS[problem], mf, itol, itask, iopt, ml, mu, maxord, mxstep, 0, 0.0, hmax,
                   hmin, tcrit, pow_mode)
    cfg2 = BatchCfg(PROBLEMS[problem], mf if mf2 is None else mf2, pick(itol, itol2), pick(itask, itask2), pick(iopt, iopt2),
                    pick(ml, ml2), pick(mu, mu2), maxord2, mxstep2, 0, 0.0, hmax2, hmin2, tcrit2, pow_mode)
    y = np.zeros((nsys, nout, neq))
    t = np.zeros((nsys, nout))
    ist = np.zeros((nsys, nout), dtype=np.int32)
    rst = np.zeros((nsys, nout, 4))
    ista = np.zeros((nsys, nout, 12), dtype=np.int32)
    L = lib(variant)
    L.dvo_batch_solve2.restype = C.c_int
    L.dvo_set_neq2.argtypes = [C.c_int]
    L.dvo_set_neq2.restype = None
    L.dvo_set_neq2(int(neq2))  # neq of the calls from the istate = 3 call on (0 = unchanged)
    nfail = L.dvo_batch_solve2(C.byref(cfg), C.byref(cfg2), int(nsplit), nsys, _dp(params), _dp(y0),
                               C.c_double(float(t0)), nout, _dp(touts), _dp(rtol), _dp(atol),
                               _dp(rtol2), _dp(atol2), _dp(y), _dp(t), _ip(ist), _dp(rst), _ip(ista),
                               nthreads)
    L.dvo_set_neq2(0)
    assert nfail >= 0
    return {"y": y, "t": t, "istate": ist, "rstats": rst, "istats": ista, "nfail": nfail}


# iwork(11:22) slot names in istats[..., k]
ISTAT = {"nst": 0, "nfe": 1, "nje": 2, "nqu": 3, "nqcur": 4, "imxer": 5, "lenrw": 6, "leniw": 7,
         "nlu": 8, "nni": 9, "ncfn": 10, "netf": 11}
RSTAT = {"hu": 0, "hcur": 1, "tcur": 2, "tolsf": 3}
