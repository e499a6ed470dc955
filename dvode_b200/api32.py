"""REAL32 twin of dvode_b200.api: the same ctypes mirror over dvode_b200/libbvode32.so, the library compiled with
-DBVODE_REAL32 (the reference's kind switch, src/dvode_kinds_module.F90:29-31).  numpy arrays are float32.
Strict arithmetic only, thread-per-system problems (robertson, vdp, vdp_maxnorm, vdp_ewtdrop)."""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "api.py")
with open(_src) as _fh:
    exec(compile(_fh.read(), _src, "exec"))
