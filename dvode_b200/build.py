"""Build libbvode.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two builds of the kernels are linked into one library:
  fast   : default code generation (FMA contraction), CUDA pow()  -- the product
  strict : -fmad=false + portable pow                             -- bit-for-bit parity runs
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbvode.so")
LIB32 = os.path.join(HERE, "libbvode32.so")  # the REAL32 build (-DBVODE_REAL32): strict arithmetic, thread problems

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found; the bvode kernels cannot be built")
    return nv


def _sources():
    out = [os.path.join(HERE, "..", "include", "bvode.h"), os.path.join(HERE, "..", "include", "bvode_mech32.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return [p for p in out if os.path.exists(p)]


def is_stale(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(p) > t for p in _sources())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nv = _nvcc()
    fast_extra = os.environ.get("BVODE_EXTRA_NVCC", "").split()
    units = [("capi.o", "bvode_capi.cu", []), ("multi.o", "bvode_multi.cu", [])]
    units.append(("lutest_fast.o", "bvode_lutest.cu", ["-DBVODE_STRICT=0"] + fast_extra))
    units.append(("lutest_strict.o", "bvode_lutest.cu", ["-DBVODE_STRICT=1", "-fmad=false"]))
    for u in range(8):  # one translation unit per problem and build, compiled in parallel
        units.append(("reg_fast_%d.o" % u, "bvode_registry.cu", ["-DBVODE_STRICT=0", "-DBVODE_UNIT=%d" % u] + fast_extra))
        units.append(("reg_strict_%d.o" % u, "bvode_registry.cu", ["-DBVODE_STRICT=1", "-fmad=false", "-DBVODE_UNIT=%d" % u]))
    procs = []
    for obj, src, extra in units:
        cmd = [nv] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nv] + ARCH + ["-shared", "-o", LIB] + [os.path.join(OBJ, u[0]) for u in units] + ["-ldl"]
    subprocess.check_call(cmd)
    return LIB


def _compile_and_link(nv, units, lib, verbose):
    procs = []
    for obj, src, extra in units:
        cmd = [nv] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nv] + ARCH + ["-shared", "-o", lib] + [os.path.join(OBJ, u[0]) for u in units] + ["-ldl"])
    return lib


def build32(force=False, verbose=False):
    """libbvode32.so: the same sources with -DBVODE_REAL32 (the reference's REAL32 kind, K:29-31): every
    `bvode_real` of the C ABI is a float.  Strict arithmetic only (each operation individually rounded, so
    that the float oracle reproduces it bit for bit) and the thread-per-system problems (units 0, 1, 4, 7)."""
    if not force and not is_stale(LIB32):
        return LIB32
    os.makedirs(OBJ, exist_ok=True)
    r32 = ["-DBVODE_REAL32"]
    units = [("capi32.o", "bvode_capi.cu", r32), ("multi32.o", "bvode_multi.cu", r32)]
    for u in (0, 1, 4, 7):
        units.append(("reg32_strict_%d.o" % u, "bvode_registry.cu",
                      r32 + ["-DBVODE_STRICT=1", "-fmad=false", "-DBVODE_UNIT=%d" % u]))
    return _compile_and_link(_nvcc(), units, LIB32, verbose)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build32(force="--force" in sys.argv, verbose="-v" in sys.argv))
