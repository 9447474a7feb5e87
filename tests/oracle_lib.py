"""ctypes binding of the CPU oracle (oracle/libgcs_oracle.so).  Test infrastructure only."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgcs_oracle.so")

gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi = gcs.capi

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        lib = C.CDLL(ORACLE_SO)
        lib.gcs_oracle_solve.argtypes = [C.POINTER(capi.CBatch), C.c_int]
        lib.gcs_oracle_newton2d.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.gcs_oracle_decision_slack.argtypes = [C.POINTER(capi.CBatch), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double), C.c_int]
        lib.gcs_oracle_qr_solve_2x2.argtypes = [C.POINTER(C.c_double)] * 3
        lib.gcs_oracle_qr_solve_2x2.restype = None
        _lib = lib
    return _lib


def solve(batch, threads=0):
    """Run the oracle on a capi.HostBatch (allocates outputs if needed)."""
    if not batch.out:
        batch.alloc_outputs()
    cb = batch.cbatch(dense=True)
    rc = load().gcs_oracle_solve(C.byref(cb), threads)
    if rc != 0:
        raise RuntimeError(f"gcs_oracle_solve -> {rc}")
    return batch


def decision_slack(batch, sdr, band, threads=0):
    """gcs_oracle_solve + slack[2][n_seeds][n]: how far the literal trajectory's convergence decisions
    stayed from the threshold beyond half the margins the contracted kernels claim (gcs_oracle.c)."""
    if not batch.out:
        batch.alloc_outputs()
    cb = batch.cbatch(dense=True)
    sdr = np.ascontiguousarray(sdr, dtype=np.float64)
    band = np.ascontiguousarray(band, dtype=np.float64)
    slack = np.empty((2, batch.n_seeds, batch.n))
    dp = C.POINTER(C.c_double)
    rc = load().gcs_oracle_decision_slack(C.byref(cb), sdr.ctypes.data_as(dp), band.ctypes.data_as(dp), slack.ctypes.data_as(dp), threads)
    if rc != 0:
        raise RuntimeError(f"gcs_oracle_decision_slack -> {rc}")
    return slack


def newton2d(kind, consts, gx, gy):
    k = (C.c_double * 12)(*(list(consts) + [0.0] * (12 - len(consts))))
    x, y, it, cv = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    rc = load().gcs_oracle_newton2d(kind, k, gx, gy, C.byref(x), C.byref(y), C.byref(it), C.byref(cv))
    assert rc == 0
    return x.value, y.value, it.value, cv.value


def qr_solve(J, rhs):
    j = (C.c_double * 4)(*J)
    r = (C.c_double * 2)(*rhs)
    s = (C.c_double * 2)()
    load().gcs_oracle_qr_solve_2x2(j, r, s)
    return np.array([s[0], s[1]])


def max_threads():
    return load().gcs_oracle_max_threads()
