"""The reference's example program (test/example.f90) as a plain C program on the C ABI
(examples/robertson.c): no Python, no ctypes between the caller and libbvode.so.
CPU: it compiles and links against include/bvode.h + libbvode.so.  GPU: its printed table is the
CPU oracle's, bit for bit (strict build), for the reference's rate constants and a scaled system."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "_build", "robertson")


def _compile():
    import dvode_b200.build as b
    b.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "dvode_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "robertson.c"), "-L", libdir, "-lbvode",
                           "-Wl,-rpath," + libdir, "-o", EXE])
    return EXE


def test_c_example_compiles_against_the_header():
    assert os.path.exists(_compile())


@pytest.mark.gpu
def test_c_example_prints_the_oracles_table():
    import oracle_lib as O
    import dvode_b200 as B
    exe = _compile()
    nsys = 5
    out = subprocess.run([exe, str(nsys), "strict"], check=True, capture_output=True, text=True).stdout
    rows = [l for l in out.splitlines() if l.startswith(" at t =")]
    assert len(rows) == 12
    k = np.array([[0.04, 1.0e4, 3.0e7]]) * (1.0 + 0.01 * np.arange(nsys))[:, None]
    tol = B.batches.ROBERTSON_TOL_REF
    ref = O.batch_solve("robertson", k, B.batches.ROBERTSON_Y0, 0.0, B.batches.robertson_touts(), tol["rtol"],
                        tol["atol"], 21, itol=2, pow_mode=1)
    for i, row in enumerate(rows):
        v = [float(x) for x in re.findall(r"[-+]?\d\.?\d*(?:[eE][-+]?\d+)?", row.split("t =")[1])]
        assert v[0] == B.batches.robertson_touts()[i]
        assert v[1:4] == list(ref["y"][0, i]) and v[4:7] == list(ref["y"][nsys - 1, i]), (i, v)
    st = [int(x) for x in re.findall(r"= +(\d+)", out.split("no. steps")[1])]
    r = ref["istats"][0, -1]
    assert st == [r[0], r[1], r[2], r[8], r[9], r[10], r[11]]
