"""User problems at run time: compile a functor (csrc/bvode_user_problem.cuh) into a problem library
with nvcc for sm_100a and register it with libbvode.so (bvode_register_problem)."""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
_loaded = {}  # path -> (CDLL, problem id): keeps the library mapped for as long as its kernels may run


def build_problem(src, out=None, strict=True, extra=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a on `src` (fast build, and the strict one unless
    strict=False), linked into the shared library `out` (default: next to the source).  Returns `out`."""
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found")
    src = os.path.abspath(src)
    out = os.path.abspath(out or os.path.splitext(src)[0] + ".so")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    base = [nv, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-I", CSRC] + list(extra)
    stem = os.path.splitext(out)[0]
    jobs = [(stem + "_fast.o", ["-DBVODE_STRICT=0"])]
    if strict:
        jobs.append((stem + "_strict.o", ["-DBVODE_STRICT=1", "-fmad=false"]))
    procs = [subprocess.Popen(base + flags + ["-c", src, "-o", obj], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for obj, flags in jobs]
    for p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log.decode())
    subprocess.check_call([nv, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] +
                          [obj for obj, _ in jobs])
    for obj, _ in jobs:
        os.remove(obj)
    return out


def register_problem(path):
    """Load a problem library and register its problem; returns (problem id, name)."""
    from . import api
    path = os.path.abspath(path)
    if path in _loaded:
        pid = _loaded[path][1]
    else:
        lib = C.CDLL(path)
        lib.bvode_user_entry_fast.restype = C.c_void_p
        lib.bvode_user_abi_fast.restype = C.c_int
        fast = lib.bvode_user_entry_fast()
        strict = None
        if hasattr(lib, "bvode_user_entry_strict"):
            lib.bvode_user_entry_strict.restype = C.c_void_p
            strict = lib.bvode_user_entry_strict()
        core = api._lib
        core.bvode_register_problem.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        core.bvode_register_problem.restype = C.c_int
        pid = core.bvode_register_problem(fast, strict, lib.bvode_user_abi_fast())
        if pid < 0:
            raise RuntimeError("bvode_register_problem: %s" % core.bvode_last_error().decode())
        _loaded[path] = (lib, pid)
    name = C.c_char_p()
    neq, npar = C.c_int(), C.c_int()
    api._lib.bvode_problem_info(pid, C.byref(neq), C.byref(npar), C.byref(name))
    return pid, name.value.decode()
