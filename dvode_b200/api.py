"""Host-side mirror of the reference's `dvode_t` type over the C ABI (include/bvode.h).

Reference interface mirrored here (file:line into /root/reference/src/dvode_module.F90):
    call solver%initialize(f, jac)                                   :377
    call solver%solve(neq,y,t,tout,itol,rtol,atol,itask,istate,
                      iopt,rwork,lrw,iwork,liw,mf)                    :604
    call solver%dvindy(t,k,yh,ldyh,dky,iflag)                        :1878
    call solver%dvsrco(sav,job)                                      :3198
Same argument names, meaning and error behaviour (per-system `istate`), for a batch of
`nsys` independent systems.  Arrays are numpy arrays modified in place, like the Fortran
intent(inout) arguments.  The f / jac callbacks are __device__ functors compiled into the
library; `initialize` selects them by problem name.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# The working precision of this module: REAL64 over libbvode.so, or -- when this file is loaded as
# dvode_b200.api32 -- REAL32 over libbvode32.so (the reference's kind switch, dvode_kinds_module.F90:29-31).
_REAL32 = __name__.endswith("api32")
REAL = np.float32 if _REAL32 else np.float64
# BVODE_LIB lets experiments point at another build of the same library (never a fallback)
_LIB_PATH = (os.path.join(_HERE, "libbvode32.so") if _REAL32 else
             (os.environ.get("BVODE_LIB") or os.path.join(_HERE, "libbvode.so")))


class BvodeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bvode error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return _LIB_PATH


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            "dvode_b200: %s is missing. Build it with `python -m dvode_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    lib.bvode_last_error.restype = C.c_char_p
    lib.bvode_problem_lookup.argtypes = [C.c_char_p]
    lib.bvode_problem_info.argtypes = [C.c_int, ip, ip, C.POINTER(C.c_char_p)]
    lib.bvode_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_uint]
    lib.bvode_destroy.argtypes = [vp]
    lib.bvode_destroy.restype = None
    lib.bvode_set_params.argtypes = [vp, vp]
    lib.bvode_set_params_device.argtypes = [vp, vp]
    lib.bvode_solve.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp,
                                C.c_int, vp, C.c_int, vp, C.c_int, C.c_int]
    lib.bvode_solve_schedule.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp,
                                         C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int,
                                         vp]
    lib.bvode_solve_schedule_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp,
                                                C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp]
    lib.bvode_solve_device.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int,
                                       vp, vp, C.c_int, vp]
    lib.bvode_dvindy.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    lib.bvode_get_stats.argtypes = [vp, vp, vp]
    lib.bvode_state_bytes.argtypes = [vp]
    lib.bvode_state_bytes.restype = C.c_longlong
    lib.bvode_dvsrco.argtypes = [vp, vp, C.c_int]
    lib.bvode_dvsrco_n.argtypes = [vp, vp, C.c_longlong, C.c_int]
    lib.bvode_reset.argtypes = [vp]
    # multi-GPU
    ll, llp = C.c_longlong, C.POINTER(C.c_longlong)
    lib.bvode_shard_range.argtypes = [ll, C.c_int, C.c_int, llp, llp]
    lib.bvode_shard_range.restype = None
    lib.bvode_comm_unique_id.argtypes = [C.c_char_p]
    lib.bvode_comm_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.bvode_comm_adopt.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int]
    lib.bvode_comm_destroy.argtypes = [vp]
    lib.bvode_comm_destroy.restype = None
    lib.bvode_gather_device.argtypes = [vp, vp, C.c_int, ll, vp, vp, vp, vp, vp, vp]
    lib.bvode_gather.argtypes = [vp, vp, C.c_int, ll, vp, vp, vp, vp, vp]
    lib.bvode_multi_create.argtypes = [C.POINTER(vp), C.c_int, ll, C.c_int, ip, C.c_uint]
    lib.bvode_multi_destroy.argtypes = [vp]
    lib.bvode_multi_destroy.restype = None
    lib.bvode_multi_device_count.argtypes = [vp]
    lib.bvode_multi_handle.argtypes = [vp, C.c_int]
    lib.bvode_multi_handle.restype = vp
    lib.bvode_multi_set_params.argtypes = [vp, vp]
    lib.bvode_multi_solve.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp,
                                      C.c_int, vp, C.c_int, vp, C.c_int, C.c_int]
    lib.bvode_multi_solve_schedule.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp,
                                               C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp]
    lib.bvode_multi_gather_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.bvode_last_launch_count.argtypes = [vp]
    lib.bvode_istate_summary.argtypes = [vp, vp]
    lib.bvode_format_message.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    lib.bvode_measure_fp64_peak.argtypes = [C.c_int, dp]
    lib.bvode_lu_test.argtypes = [C.c_int] * 5 + [vp] * 6 + [C.c_int, C.c_uint]
    return lib


_lib = _load()

FLAG_STRICT = 1
FLAG_COMPACT = 2  # hand finished threads the next system of the batch (bvode.h)
FLAG_PER_SYSTEM = 4  # rtol / atol / optional inputs per system, as the reference has them per instance
ISTAT = {"nst": 0, "nfe": 1, "nje": 2, "nqu": 3, "nqcur": 4, "imxer": 5, "lenrw": 6, "leniw": 7,
         "nlu": 8, "nni": 9, "ncfn": 10, "netf": 11}
RSTAT = {"hu": 0, "hcur": 1, "tcur": 2, "tolsf": 3}


def _check(rc):
    if rc != 0:
        raise BvodeError(rc, (_lib.bvode_last_error() or b"").decode())


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=REAL)
    if shape is not None:
        a = a.reshape(shape)
    return a


def problem_names():
    out = []
    for i in range(_lib.bvode_problem_count()):
        name = C.c_char_p()
        _lib.bvode_problem_info(i, None, None, C.byref(name))
        out.append(name.value.decode())
    return out


def problem_info(name):
    pid = _lib.bvode_problem_lookup(name.encode())
    if pid < 0:
        raise BvodeError(-1, "unknown problem %r (have %s)" % (name, problem_names()))
    neq, npar = C.c_int(), C.c_int()
    _check(_lib.bvode_problem_info(pid, C.byref(neq), C.byref(npar), None))
    return pid, neq.value, npar.value


def fp64_peak_tflops(device=0):
    v = C.c_double()
    _check(_lib.bvode_measure_fp64_peak(device, C.byref(v)))
    return v.value


class dvode_t:
    """Batched counterpart of the reference's `type(dvode_t)` (M:251-314)."""

    def __init__(self):
        self._h = None
        self.nsys = self.neq = self.npar = 0

    # -- initialize(f, jac) -> initialize(problem, nsys, params) ------------------------
    def initialize(self, problem, nsys, params=None, device=0, strict=False, compact=False,
                   per_system=False):
        """per_system=True: rtol / atol are [nsys, 1 or neq] and the optional inputs are read from
        every row of rwork / iwork (each system its own solver instance, M:729-751)."""
        self.close()
        pid, self.neq, self.npar = problem_info(problem)
        h = C.c_void_p()
        _check(_lib.bvode_create(C.byref(h), pid, int(nsys), int(device),
                                 (FLAG_STRICT if strict else 0) | (FLAG_COMPACT if compact else 0) |
                                 (FLAG_PER_SYSTEM if per_system else 0)))
        self._h = h
        self.nsys = int(nsys)
        self.device = device
        if params is not None:
            self.set_params(params)
        return self

    def set_params(self, params):
        p = _f64(params, (self.nsys, max(self.npar, 1)))
        _check(_lib.bvode_set_params(self._h, _ptr(p)))

    def set_params_device(self, ptr):
        _check(_lib.bvode_set_params_device(self._h, C.c_void_p(ptr)))

    def close(self):
        if self._h is not None:
            _lib.bvode_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- solve(neq,y,t,tout,itol,rtol,atol,itask,istate,iopt,rwork,lrw,iwork,liw,mf) ---
    def solve(self, neq, y, t, tout, itol, rtol, atol, itask, istate, iopt, rwork, lrw, iwork,
              liw, mf):
        """y[nsys,neq], t[nsys], istate[nsys] (int32) are updated in place.  tout: scalar or
        array[nsys].  rwork[nsys,lrw] / iwork[nsys,liw] (or None) carry the optional
        inputs (row 0) and receive the optional outputs (every row)."""
        self._neq_of_call(neq, istate)
        self._chk_arrays(y, t, istate, rwork, iwork, lrw, liw)
        tout = _f64(tout).reshape(-1)
        per = 1 if tout.size > 1 else 0
        if per and tout.size != self.nsys:
            raise BvodeError(-1, "tout must be a scalar or have nsys entries")
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        _check(_lib.bvode_solve(self._h, int(neq), _ptr(y), _ptr(t), _ptr(tout), per, int(itol),
                                _ptr(rtol), _ptr(atol), int(itask), _ptr(istate), int(iopt),
                                _ptr(rwork), int(lrw), _ptr(iwork), int(liw), int(mf)))

    def solve_schedule(self, neq, y, t, touts, itol, rtol, atol, itask, istate, iopt, rwork, lrw,
                       iwork, liw, mf, y_out=None):
        """The reference's `do iout = 1, nout; call solve(...); tout = ...` loop as one call.
        y_out[nout,nsys,neq] (optional) receives every output point."""
        self._neq_of_call(neq, istate)
        self._chk_arrays(y, t, istate, rwork, iwork, lrw, liw)
        touts = _f64(touts).reshape(-1)
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        if y_out is not None:
            assert y_out.dtype == REAL and y_out.flags.c_contiguous
            assert y_out.shape == (touts.size, self.nsys, self.neq)
        _check(_lib.bvode_solve_schedule(self._h, int(neq), _ptr(y), _ptr(t), touts.size,
                                         _ptr(touts), int(itol), _ptr(rtol), _ptr(atol),
                                         int(itask), _ptr(istate), int(iopt), _ptr(rwork),
                                         int(lrw), _ptr(iwork), int(liw), int(mf), _ptr(y_out)))

    def solve_schedule_device(self, neq, y_ptr, t_ptr, touts, itol, rtol, atol, itask, istate_ptr,
                              iopt, rwork_opts, iwork_opts, mf, y_out_ptr=0, stream=0):
        """Device-resident variant: *_ptr are raw device addresses, SoA layout (bvode.h)."""
        touts = _f64(touts).reshape(-1)
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        _check(_lib.bvode_solve_schedule_device(
            self._h, int(neq), C.c_void_p(y_ptr), C.c_void_p(t_ptr), touts.size, _ptr(touts),
            int(itol), _ptr(rtol), _ptr(atol), int(itask), C.c_void_p(istate_ptr), int(iopt),
            _ptr(rwork_opts), _ptr(iwork_opts), int(mf),
            C.c_void_p(y_out_ptr) if y_out_ptr else None, C.c_void_p(stream) if stream else None))

    def solve_device(self, neq, y_ptr, t_ptr, tout_ptr, itol, rtol, atol, itask, istate_ptr, iopt,
                     rwork_opts, iwork_opts, mf, stream=0):
        """One call with device-resident SoA buffers and a tout per system (tout_ptr: double[nsys] in HBM)."""
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        _check(_lib.bvode_solve_device(
            self._h, int(neq), C.c_void_p(y_ptr), C.c_void_p(t_ptr), C.c_void_p(tout_ptr), int(itol),
            _ptr(rtol), _ptr(atol), int(itask), C.c_void_p(istate_ptr), int(iopt), _ptr(rwork_opts),
            _ptr(iwork_opts), int(mf), C.c_void_p(stream) if stream else None))

    # -- dvindy(t, k, yh, ldyh, dky, iflag): yh/ldyh are the handle's device state -------
    def dvindy(self, t, k):
        t = _f64(t).reshape(-1)
        per = 1 if t.size > 1 else 0
        dky = np.zeros((self.nsys, self.neq), dtype=REAL)
        iflag = np.zeros(self.nsys, dtype=np.int32)
        _check(_lib.bvode_dvindy(self._h, _ptr(t), per, int(k), _ptr(dky), _ptr(iflag)))
        return dky, iflag

    # -- dvsrco(sav, job) -------------------------------------------------------------------
    def dvsrco(self, sav, job):
        if job == 1:
            n = _lib.bvode_state_bytes(self._h)
            sav = np.zeros(n, dtype=np.uint8)
        _check(_lib.bvode_dvsrco_n(self._h, _ptr(sav), int(sav.nbytes), int(job)))
        return sav

    def reset(self):
        """Drop the integration state (the next call must have istate = 1)."""
        _check(_lib.bvode_reset(self._h))

    # -- multi-GPU: the final gather of solutions and statistics (bvode.h) -------------------
    def gather_device(self, comm, root, nsys_total, y_all=0, t_all=0, istate_all=0, rout_all=0,
                      iout_all=0, stream=0):
        """Device pointers (ints) on the root; on other ranks any non-zero value = wanted."""
        vp = lambda x: C.c_void_p(x) if x else None
        _check(_lib.bvode_gather_device(self._h, comm._c, int(root), int(nsys_total), vp(y_all),
                                        vp(t_all), vp(istate_all), vp(rout_all), vp(iout_all),
                                        vp(stream)))

    def gather(self, comm, root, nsys_total, want=("y", "t", "istate", "rout", "iout")):
        """Host arrays on the root (dict of numpy arrays); None on the other ranks."""
        is_root = comm.rank == root
        out = {}
        shapes = {"y": ((nsys_total, self.neq), REAL), "t": ((nsys_total,), REAL),
                  "istate": ((nsys_total,), np.int32), "rout": ((nsys_total, 4), REAL),
                  "iout": ((nsys_total, 12), np.int32)}
        args = []
        for k in ("y", "t", "istate", "rout", "iout"):
            if k in want:
                a = np.zeros(*shapes[k]) if is_root else np.zeros(1, dtype=shapes[k][1])
                out[k] = a
                args.append(_ptr(a))
            else:
                args.append(None)
        _check(_lib.bvode_gather(self._h, comm._c, int(root), int(nsys_total), *args))
        return out if is_root else None

    def stats(self):
        """Optional outputs: (rwork(11:14)[nsys,4], iwork(11:22)[nsys,12])."""
        r = np.zeros((self.nsys, 4), dtype=REAL)
        i = np.zeros((self.nsys, 12), dtype=np.int32)
        _check(_lib.bvode_get_stats(self._h, _ptr(r), _ptr(i)))
        return r, i

    def last_launch_count(self):
        return _lib.bvode_last_launch_count(self._h)

    # -- xerrwd (M:3351-3394) as a host-side message stream --------------------------------
    def istate_summary(self):
        """{istate: number of systems} of the last host-buffer solve (0 = no error)."""
        c = np.zeros(8, dtype=np.int32)
        _check(_lib.bvode_istate_summary(self._h, _ptr(c)))
        out = {0: int(c[0])}
        out.update({-k: int(c[k]) for k in range(1, 7)})
        out["other"] = int(c[7])
        return out

    def message(self, sys):
        """The reference's xerrwd text for system `sys` ('' when it ended without error)."""
        buf = C.create_string_buffer(1024)
        n = _lib.bvode_format_message(self._h, int(sys), buf, 1024)
        if n < 0:
            _check(n)
        return buf.value.decode()

    def _neq_of_call(self, neq, istate):
        """An istate = 3 call may pass a smaller neq (M:1127-1135): y, y_out and dky then have that many columns."""
        if self._h is not None and isinstance(istate, np.ndarray) and istate.size and int(neq) < self.neq and \
                (istate == 3).all() and int(neq) >= 1:
            self.neq = int(neq)

    def _chk_arrays(self, y, t, istate, rwork, iwork, lrw, liw):
        if self._h is None:
            raise BvodeError(-1, "initialize() has not been called")
        ok = (isinstance(y, np.ndarray) and y.dtype == REAL and y.flags.c_contiguous and
              y.size == self.nsys * self.neq and isinstance(t, np.ndarray) and
              t.dtype == REAL and t.size == self.nsys and isinstance(istate, np.ndarray) and
              istate.dtype == np.int32 and istate.size == self.nsys)
        if rwork is not None:
            ok = ok and rwork.dtype == REAL and rwork.flags.c_contiguous and \
                rwork.size == self.nsys * lrw
        if iwork is not None:
            ok = ok and iwork.dtype == np.int32 and iwork.flags.c_contiguous and \
                iwork.size == self.nsys * liw
        if not ok:
            raise BvodeError(-1, "y/t/istate/rwork/iwork must be contiguous numpy arrays of "
                                 "real/real/int32/real/int32 (real = %s) with nsys rows" % REAL.__name__)


def lu_test(kind, a, b, ml=0, mu=0, strict=False, device=0):
    """The device LINPACK routines in isolation (bvode_lu_test): a[nmat, n, n] with a[m, i, j] = A(i, j)
    (kind 0-2) or a[nmat, n, lda] LINPACK band columns (kind 3), b[nmat, n].  Returns
    (afac, ipvt, x, info) as described in include/bvode.h; afac for kind 0 comes back as [m, i, j]."""
    a = np.asarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nmat, n = b.shape
    if kind == 3:
        ain = np.ascontiguousarray(a)
    else:
        ain = np.ascontiguousarray(a.transpose(0, 2, 1))  # column-major per matrix
    afac = np.zeros_like(ain)
    ipvt = np.zeros((nmat, n), dtype=np.int32)
    x = np.zeros((nmat, n))
    info = np.zeros(nmat, dtype=np.int32)
    _check(_lib.bvode_lu_test(int(kind), int(n), int(ml), int(mu), int(nmat), _ptr(ain), _ptr(b), _ptr(afac),
                              _ptr(ipvt), _ptr(x), _ptr(info), int(device), FLAG_STRICT if strict else 0))
    if kind == 0:
        afac = afac.transpose(0, 2, 1)
    return afac, ipvt, x, info


def shard_range(nsys_total, rank, world):
    """[first, last) of the contiguous block of systems owned by `rank` (bvode_shard_range)."""
    f, l = C.c_longlong(), C.c_longlong()
    _lib.bvode_shard_range(int(nsys_total), int(rank), int(world), C.byref(f), C.byref(l))
    return f.value, l.value


def nccl_version():
    return _lib.bvode_nccl_version()


class Comm:
    """An NCCL communicator behind the C ABI (bvode_comm_*): one process per GPU."""

    def __init__(self, world, rank, device, unique_id=None):
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        c = C.c_void_p()
        _check(_lib.bvode_comm_create(C.byref(c), unique_id, self.world, self.rank, self.device))
        self._c = c

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _check(_lib.bvode_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if self._c is not None:
            _lib.bvode_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class dvode_multi_t:
    """One process, several GPUs (bvode_multi_*): the batch is sharded by system index over the
    devices; solve / solve_schedule take arrays of nsys_total rows with the arguments of dvode_t."""

    def __init__(self):
        self._m = None

    def initialize(self, problem, nsys_total, params=None, devices=None, ndev=None, strict=False,
                   per_system=False):
        self.close()
        pid, self.neq, self.npar = problem_info(problem)
        if devices is None:
            devices = list(range(int(ndev or 1)))
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        m = C.c_void_p()
        _check(_lib.bvode_multi_create(C.byref(m), pid, int(nsys_total), dv.size,
                                       dv.ctypes.data_as(C.POINTER(C.c_int)),
                                       (FLAG_STRICT if strict else 0) | (FLAG_PER_SYSTEM if per_system else 0)))
        self._m = m
        self.nsys = int(nsys_total)
        self.ndev = dv.size
        if params is not None:
            p = _f64(params, (self.nsys, max(self.npar, 1)))
            _check(_lib.bvode_multi_set_params(self._m, _ptr(p)))
        return self

    def solve(self, neq, y, t, tout, itol, rtol, atol, itask, istate, iopt, rwork, lrw, iwork, liw, mf):
        tout = _f64(tout).reshape(-1)
        per = 1 if tout.size > 1 else 0
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        _check(_lib.bvode_multi_solve(self._m, int(neq), _ptr(y), _ptr(t), _ptr(tout), per, int(itol),
                                      _ptr(rtol), _ptr(atol), int(itask), _ptr(istate), int(iopt),
                                      _ptr(rwork), int(lrw), _ptr(iwork), int(liw), int(mf)))

    def solve_schedule(self, neq, y, t, touts, itol, rtol, atol, itask, istate, iopt, rwork, lrw,
                       iwork, liw, mf, y_out=None):
        touts = _f64(touts).reshape(-1)
        rtol, atol = _f64(rtol).reshape(-1), _f64(atol).reshape(-1)
        if y_out is not None:
            assert y_out.dtype == REAL and y_out.flags.c_contiguous
            assert y_out.shape == (touts.size, self.nsys, self.neq)
        _check(_lib.bvode_multi_solve_schedule(self._m, int(neq), _ptr(y), _ptr(t), touts.size, _ptr(touts),
                                               int(itol), _ptr(rtol), _ptr(atol), int(itask), _ptr(istate),
                                               int(iopt), _ptr(rwork), int(lrw), _ptr(iwork), int(liw),
                                               int(mf), _ptr(y_out)))

    def gather_device(self, root_slot, y_all=0, t_all=0, istate_all=0, rout_all=0, iout_all=0):
        vp = lambda x: C.c_void_p(x) if x else None
        _check(_lib.bvode_multi_gather_device(self._m, int(root_slot), vp(y_all), vp(t_all),
                                              vp(istate_all), vp(rout_all), vp(iout_all)))

    def close(self):
        if self._m is not None:
            _lib.bvode_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
