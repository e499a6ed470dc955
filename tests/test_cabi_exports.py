"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/bvode.h declares (no compute without a GPU), and the host-side helpers behave."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dvode_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "bvode.h")).read()
    names = re.findall(r"BVODE_API[^;(]*?\b(bvode_\w+)\s*\(", hdr)
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n


def test_registry_and_python_mirror_import():
    import dvode_b200 as B
    names = B.problem_names()
    assert "robertson" in names and "vdp" in names and "vdp_maxnorm" in names
    assert B.problem_info("robertson")[1:] == (3, 3)
    assert B.problem_info("vdp")[1:] == (2, 1)


def test_batch_generator_is_deterministic_and_bounded():
    import dvode_b200 as B
    a = B.batches.robertson_batch(1000)
    b = B.batches.robertson_batch(500, first=500)
    assert np.array_equal(a[500:], b)
    base = np.array(B.batches.ROBERTSON_K)
    rel = a / base - 1.0
    assert rel.min() >= -0.1 and rel.max() < 0.1 and abs(rel.mean()) < 0.01
    # scalar restatement of the hash for one element (seed + npar*i + j)
    M = (1 << 64) - 1
    z = (B.batches.SEED + 3 * 7 + 1 + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z = z ^ (z >> 31)
    u = (z >> 11) / float(1 << 53) * 2.0 - 1.0
    assert a[7, 1] == base[1] * (1.0 + 0.1 * u)


def test_fortran_binding_names_are_exported():
    """fortran/bvode_module.f90 cannot be compiled here (no Fortran compiler in the image): at
    least every bind(C, name=...) in it must exist in the library."""
    from dvode_b200 import build
    lib = ctypes.CDLL(build.build())
    src = open(os.path.join(ROOT, "fortran", "bvode_module.f90")).read()
    names = set(re.findall(r"bind\(C,\s*name='(\w+)'\)", src))
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_host_mirror_rejects_bad_arrays_without_a_gpu():
    import dvode_b200 as B
    s = B.dvode_t()
    import pytest
    with pytest.raises(B.BvodeError):
        s.solve(3, np.zeros((4, 3)), np.zeros(4), 1.0, 1, [1e-4], [1e-8], 1,
                np.ones(4, dtype=np.int32), 0, None, 0, None, 0, 21)
