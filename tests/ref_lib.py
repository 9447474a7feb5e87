"""ctypes binding of oracle/_ref/libgcs_ref.so: the reference's own solve2D / primitives /
heuristics / solver translation units compiled against stand-in third-party headers
(oracle/build_ref.sh).  Exists only where it was built (the container with /root/reference, and
the GPU box, to which the built .so travels).  Test infrastructure only."""
import ctypes as C
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgcs_ref.so")
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi = gcs.capi

_lib = None


class RefElement(C.Structure):
    _fields_ = [("type", C.c_int32), ("is_set", C.c_int32), ("canvas", C.c_double * 4), ("pos", C.c_double * 4)]


class RefEdge(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("type", C.c_int32), ("flip", C.c_int32), ("value", C.c_double)]


def available():
    return os.path.exists(REF_SO)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(REF_SO)
        lib.gcs_ref_solve_batch.argtypes = [C.POINTER(capi.CBatch), C.c_int, C.c_int]
        lib.gcs_ref_component_solve.argtypes = [C.c_int, C.POINTER(RefElement), C.c_int, C.POINTER(RefEdge)]
        if hasattr(lib, "gcs_ref_leaves_solve"):
            lib.gcs_ref_leaves_solve.argtypes = [C.c_int, C.POINTER(RefElement), C.c_int, C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_int32), C.POINTER(RefEdge), C.POINTER(C.c_int32)]
        if hasattr(lib, "gcs_ref_m3_free_line"):
            dp, bp = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
            lib.gcs_ref_m3_point_pp.argtypes = [C.c_int64, dp, dp]
            for f in (lib.gcs_ref_m3_free_line, lib.gcs_ref_m3_point_pl, lib.gcs_ref_m3_point_ll):
                f.argtypes = [C.c_int64, dp, dp, bp]
            lib.gcs_ref_m3_rigid_transform.argtypes = [C.c_int, dp, dp, dp]
            lib.gcs_ref_m3_score.argtypes = [C.c_int, C.POINTER(C.c_int32), dp, dp, bp]
            lib.gcs_ref_m3_score.restype = C.c_double
        if hasattr(lib, "gcs_ref_m3_merge"):
            ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
            lib.gcs_ref_m3_merge.restype = C.c_int
            lib.gcs_ref_m3_merge.argtypes = [C.c_int, C.c_int, ip, dp, ip, ip, dp, ip, dp, ip]
        if hasattr(lib, "gcs_ref_m3_ppp_merge"):
            ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
            lib.gcs_ref_m3_ppp_merge.argtypes = [C.c_int, ip, dp, ip, ip, dp, ip, dp]
        if hasattr(lib, "gcs_ref_model_solve_transform"):
            dp, bp, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
            lib.gcs_ref_model_solve_transform.argtypes = [C.c_int, ip, dp, dp, bp, C.c_int, ip, dp, dp, dp]
        _lib = lib
    return _lib


def solve_batch(batch, count_iters=True, threads=0):
    if not batch.out:
        batch.alloc_outputs()
    cb = batch.cbatch(dense=True)
    rc = load().gcs_ref_solve_batch(C.byref(cb), 1 if count_iters else 0, threads)
    if rc != 0:
        raise RuntimeError(f"gcs_ref_solve_batch -> {rc}")
    return batch


def component_solve(elements, edges):
    """elements: list of dicts {type, canvas, pos?, is_set?}; edges: list of dicts {a, b, type, value, flip}.
    Returns (status, elements-with-positions).  stderr of the reference (its per-solve prints) is
    left alone."""
    els = (RefElement * len(elements))()
    for i, e in enumerate(elements):
        els[i].type = e["type"]
        els[i].is_set = 1 if e.get("is_set") else 0
        for j, v in enumerate(e["canvas"]):
            els[i].canvas[j] = v
        for j, v in enumerate(e.get("pos", [])):
            els[i].pos[j] = v
    eds = (RefEdge * len(edges))()
    for i, e in enumerate(edges):
        eds[i].a, eds[i].b, eds[i].type = e["a"], e["b"], e["type"]
        eds[i].flip = 1 if e.get("flip") else 0
        eds[i].value = e.get("value", 0.0)
    status = load().gcs_ref_component_solve(len(elements), els, len(edges), eds)
    out = []
    for i, e in enumerate(elements):
        k = 2 if e["type"] == 0 else 4
        out.append({"type": e["type"], "is_set": bool(els[i].is_set), "pos": [els[i].pos[j] for j in range(k)]})
    return status, out


def leaves_solve(elements, leaves):
    """The reference's sequential for_each(leaves, classifyAndSolve) over shared elements.
    Returns (rc, per-leaf status, elements-with-positions)."""
    import host_lib as H
    els, _ = H.to_c(elements, [], RefElement, RefEdge)
    flat, offs = [], [0]
    for lf in leaves:
        flat += lf["edges"]
        offs.append(len(flat))
    _, eds = H.to_c([], flat, RefElement, RefEdge)
    n = len(leaves)
    le = (C.c_int32 * max(3 * n, 1))(*[i for lf in leaves for i in lf["elems"]])
    eo = (C.c_int32 * (n + 1))(*offs)
    status = (C.c_int32 * max(n, 1))()
    rc = load().gcs_ref_leaves_solve(len(elements), els, n, le, eo, eds, status)
    return rc, list(status)[:n], H.from_c(elements, els)


# ---- the reference's bottom-up Merge3 numeric helpers (oracle/ref_merge3_driver.cpp) ----
def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def m3_solve(kase, rows):
    """kase 1 point from two points, 2 line from two points, 3 point from point+line, 4 point from two
    lines (row layouts: oracle/ref_merge3_driver.cpp).  Returns (out, ok)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    nout = {1: 2, 2: 4, 3: 2, 4: 2}[kase]
    out = np.zeros((n, nout))
    ok = np.ones(n, dtype=np.uint8)
    okp = ok.ctypes.data_as(C.POINTER(C.c_uint8))
    lib = load()
    if kase == 1:
        lib.gcs_ref_m3_point_pp(n, _dp(rows), _dp(out))
    else:
        {2: lib.gcs_ref_m3_free_line, 3: lib.gcs_ref_m3_point_pl, 4: lib.gcs_ref_m3_point_ll}[kase](n, _dp(rows), _dp(out), okp)
    return out, ok


def m3_ppp_merge(types, canvas4, clusters):
    """The reference's own Merge3PppSolver::solve.  Returns (n merged or 0, ids, pose4).  Its progress
    lines go to stderr."""
    import host_lib as H
    types, canvas4, counts, ids, pose4, out_ids, out_pose, ip = H._m3_ppp_args(types, canvas4, clusters)
    n = load().gcs_ref_m3_ppp_merge(len(types), types.ctypes.data_as(ip), _dp(canvas4), counts.ctypes.data_as(ip), ids.ctypes.data_as(ip),
                                    _dp(pose4), out_ids.ctypes.data_as(ip), _dp(out_pose))
    return n, out_ids[:max(n, 0)].copy(), out_pose[:max(n, 0)].copy()


def m3_merge(which, types, canvas4, clusters):
    """The reference's own Merge3{Ppp,Pll,Lpp,Llp,Fallback}Solver::solve, or ("node") the case order of a merge
    node.  Returns (n merged or 0, ids, pose4, case that produced the pose).  Progress lines go to stderr."""
    import host_lib as H
    types, canvas4, counts, ids, pose4, out_ids, out_pose, ip = H._m3_ppp_args(types, canvas4, clusters)
    by = C.c_int32(-1)
    n = load().gcs_ref_m3_merge(H.M3_CASES[which], len(types), types.ctypes.data_as(ip), _dp(canvas4), counts.ctypes.data_as(ip),
                                ids.ctypes.data_as(ip), _dp(pose4), out_ids.ctypes.data_as(ip), _dp(out_pose), C.byref(by))
    return n, out_ids[:max(n, 0)].copy(), out_pose[:max(n, 0)].copy(), by.value


def m3_rigid_transform(src, dst):
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    out = np.zeros(6)
    rc = load().gcs_ref_m3_rigid_transform(src.shape[0], _dp(src), _dp(dst), _dp(out))
    return rc, out


def m3_score(types, canvas4, pose4, in_pose):
    types = np.ascontiguousarray(types, dtype=np.int32)
    canvas4 = np.ascontiguousarray(canvas4, dtype=np.float64)
    pose4 = np.ascontiguousarray(pose4, dtype=np.float64)
    in_pose = np.ascontiguousarray(in_pose, dtype=np.uint8)
    return load().gcs_ref_m3_score(len(types), types.ctypes.data_as(C.POINTER(C.c_int32)), _dp(canvas4), _dp(pose4),
                                   in_pose.ctypes.data_as(C.POINTER(C.c_uint8)))


def model_solve_transform(types, canvas4, pos4, solved, con4, value):
    """The reference GUI model: build it (add* calls), install the given solver positions in place
    of the solve, run solveConstraintSystem() -> applySolverToCanvasTransform.
    Returns (accepted constraints, canvas4 afterwards, stored constraint values)."""
    types = np.ascontiguousarray(types, dtype=np.int32)
    canvas4 = np.ascontiguousarray(canvas4, dtype=np.float64)
    pos4 = np.ascontiguousarray(pos4, dtype=np.float64)
    solved = np.ascontiguousarray(solved, dtype=np.uint8)
    con4 = np.ascontiguousarray(con4, dtype=np.int32).reshape(-1, 4)
    value = np.ascontiguousarray(value, dtype=np.float64)
    out = np.zeros_like(canvas4)
    stored = np.zeros(max(len(value), 1))
    ip = C.POINTER(C.c_int32)
    n = load().gcs_ref_model_solve_transform(len(types), types.ctypes.data_as(ip), _dp(canvas4), _dp(pos4),
                                             solved.ctypes.data_as(C.POINTER(C.c_uint8)), len(value),
                                             con4.ctypes.data_as(ip), _dp(value), _dp(out), _dp(stored))
    return n, out, stored[:len(value)]
