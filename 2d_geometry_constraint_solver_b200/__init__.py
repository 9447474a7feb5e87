"""B200-native batched Newton-Raphson sub-system solver (one hot path of
SolyomBalint/2D_geometry_constraint_solver).

The product is `libgcs_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/gcs_b200.h) and the C++ host mirror in host/.  The Python here is plumbing for the
benchmark and the tests.  Import as
    gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
(the directory name starts with a digit, so the `import` statement cannot spell it).
"""
from . import capi, sketch_io, synth  # noqa: F401

__version__ = "0.1.0"
