"""orbit_b200 — B200-native (sm_100a CUDA) implementation of Orbit's GPU-driven visibility pipeline behind the
reference's culling-pass interface (src/passes/draw_gen.rs, src/passes/cluster.rs).

    orbit_b200.passes   host mirror of the reference pass API over the C ABI (include/orbit_cuda.h)
    orbit_b200.frame    the caller protocol (early -> Hi-Z -> late -> main; shadow cascades)
    orbit_b200.scenes   procedural inputs in the reference buffer layouts
    orbit_b200.layouts  byte layouts (numpy dtypes + ctypes structs)
    orbit_b200.build    nvcc build of orbit_b200/lib/liborbit_b200.so

Nothing here computes on the CPU: every stage call goes to the CUDA library and raises if it is not built or
no GPU is present.
"""
from . import layouts, scenes  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: passes / frame import torch
    if name in ("passes", "frame", "multi_gpu"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
