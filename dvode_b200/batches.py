"""Synthetic parameter-perturbed batches.

System i of a batch gets parameters p_j = p0_j * (1 + amp * u_ij) with
u_ij in [-1, 1) taken from splitmix64(seed + npar*i + j): a pure integer hash, so the
same batch can be regenerated bit for bit anywhere (host, device, oracle side).
"""
import numpy as np

SEED = 20261019
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser of the uint64 array x (state increment included)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_pm1(nsys, npar, seed=SEED, first=0):
    """u[nsys, npar] in [-1, 1): top 53 bits of the hash / 2^53 * 2 - 1."""
    idx = (np.arange(first, first + nsys, dtype=np.uint64)[:, None] * np.uint64(npar) +
           np.arange(npar, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        z = splitmix64(idx + np.uint64(seed))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0


def perturbed(base, nsys, amp=0.1, seed=SEED, first=0):
    base = np.asarray(base, dtype=np.float64)
    return base[None, :] * (1.0 + amp * uniform_pm1(nsys, base.size, seed, first))


ROBERTSON_K = (0.04, 1.0e4, 3.0e7)           # example.f90:87-89
ROBERTSON_Y0 = (1.0, 0.0, 0.0)               # example.f90:38-40
ROBERTSON_TOUTS = tuple(0.4 * 10.0 ** k for k in range(12))


def robertson_touts():
    """tout = 0.4, then tout = tout*10 eleven times, accumulated like example.f90:66."""
    t, out = 0.4, []
    for _ in range(12):
        out.append(t)
        t = t * 10.0
    return np.array(out)


def robertson_batch(nsys, amp=0.1, seed=SEED, first=0):
    return perturbed(ROBERTSON_K, nsys, amp, seed, first)


# tolerances: the reference's own (example.f90:44-47) and the tightened headline set
ROBERTSON_TOL_REF = dict(itol=2, rtol=[1e-4], atol=[1e-8, 1e-14, 1e-6])
ROBERTSON_TOL_HEADLINE = dict(itol=2, rtol=[1e-6], atol=[1e-10, 1e-16, 1e-8])


# ---- dvdemo problem 2 (banded advection, neq = 25, ml = 5, mu = 0), dvdemo.f90:116-190 ----
ADVECTION_PAR = (1.0, 1.0)


def advection_y0():
    y0 = np.zeros(25)
    y0[0] = 1.0
    return y0


def advection_touts():
    t, out = 0.01, []
    for _ in range(5):
        out.append(t)
        t = t * 10.0
    return np.array(out)


def advection_batch(nsys, amp=0.1, seed=SEED, first=0):
    return perturbed(ADVECTION_PAR, nsys, amp, seed, first)


ADVECTION_TOL = dict(itol=1, rtol=[0.0], atol=[1e-6])


# ---- synthetic 32-species mechanism (BASELINE.json config 4; builder-defined) ------------
def _mech_k0():
    import os
    import re
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                       "bvode_mech32.h")
    txt = open(hdr).read()
    m = re.search(r"#define BVODE_MECH_K0_INIT \{([^}]*)\}", txt)
    return np.array([float(x) for x in m.group(1).split(",")])


def mech32_y0():
    y0 = np.zeros(32)
    y0[:4] = 0.25
    return y0


def mech32_touts():
    return np.array([1.0, 10.0, 100.0, 1000.0])


def mech32_batch(nsys, amp=0.1, seed=SEED, first=0):
    return perturbed(_mech_k0(), nsys, amp, seed, first)


MECH32_TOL = dict(itol=1, rtol=[1e-6], atol=[1e-12])


# ---- reaction-diffusion chain, n = 48 (builder-defined; two components per lane) -----------
DIFFUSION48_PAR = (50.0, 1.0)
DIFFUSION48_TOL = dict(itol=1, rtol=[1e-6], atol=[1e-9])


def diffusion48_batch(nsys, amp=0.1, seed=SEED, first=0):
    return perturbed(DIFFUSION48_PAR, nsys, amp, seed, first)


def diffusion48_y0():
    return np.zeros(48)


def diffusion48_touts():
    return np.array([0.01, 0.1, 1.0, 10.0])
