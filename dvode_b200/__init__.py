"""dvode_b200 -- B200-native batched VODE integrator behind dvode's operator API.

The product is the CUDA library `libbvode.so` (sources in dvode_b200/csrc, C ABI in
include/bvode.h).  This package is the thin host-side mirror of the reference's
`dvode_t` interface (initialize / solve / dvindy / dvsrco) over that C ABI.  There is
no CPU fallback: importing `dvode_b200.api` fails loudly if the library is missing.
"""
from .api import (BvodeError, dvode_t, fp64_peak_tflops, library_path, problem_info,  # noqa: F401
                  problem_names)
from . import batches  # noqa: F401

__all__ = ["dvode_t", "BvodeError", "problem_info", "problem_names", "fp64_peak_tflops",
           "library_path", "batches"]
