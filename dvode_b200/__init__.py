"""dvode_b200 -- B200-native batched VODE integrator behind dvode's operator API.

The product is the CUDA library `libbvode.so` (sources in dvode_b200/csrc, C ABI in
include/bvode.h).  This package is the thin host-side mirror of the reference's
`dvode_t` interface (initialize / solve / dvindy / dvsrco) over that C ABI.  There is
no CPU fallback: touching any solver name (`dvode_b200.dvode_t`, `dvode_b200.api`, ...)
fails loudly if the library is missing.  Only `dvode_b200.build` (the nvcc recipe) and
`dvode_b200.batches` / `dvode_b200.sharding` (pure host helpers) import without it.
"""
import importlib

_API_NAMES = ("BvodeError", "Comm", "dvode_multi_t", "dvode_t", "fp64_peak_tflops", "library_path",
              "nccl_version", "problem_info", "problem_names")
__all__ = list(_API_NAMES) + ["api", "api32", "batches", "build", "plugins", "sharding"]


def __getattr__(name):
    if name in _API_NAMES:
        return getattr(importlib.import_module(".api", __name__), name)
    if name in ("api", "api32", "batches", "build", "plugins", "sharding"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
