"""ctypes binding of the C entry points of the host mirror
(2d_geometry_constraint_solver_b200/host/libgcs_host.so, src/capi_host.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "2d_geometry_constraint_solver_b200", "host")
HOST_SO = os.path.join(HOST_DIR, "libgcs_host.so")

_lib = None


class Element(C.Structure):
    _fields_ = [("type", C.c_int32), ("is_set", C.c_int32), ("canvas", C.c_double * 4), ("pos", C.c_double * 4)]


class Edge(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("type", C.c_int32), ("flip", C.c_int32), ("value", C.c_double)]


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_SO):
            subprocess.check_call(["make", "-C", HOST_DIR, "-s"])
        lib = C.CDLL(HOST_SO)
        lib.gcs_host_last_error.restype = C.c_char_p
        lib.gcs_host_component_solve.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(Edge)]
        lib.gcs_host_system_solve.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(Edge)]
        lib.gcs_host_component_pack.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(Edge), C.POINTER(C.c_int32),
                                                C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        lib.gcs_host_leaves_solve.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                              C.POINTER(Edge), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        lib.gcs_host_system_solve_ex.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(Edge), C.POINTER(C.c_int64)]
        lib.gcs_host_decompose.argtypes = [C.c_int, C.POINTER(Element), C.c_int, C.POINTER(Edge), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
        lib.gcs_host_solve2d.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        if hasattr(lib, "gcs_host_canvas_transform"):
            lib.gcs_host_canvas_transform.argtypes = [C.c_int, C.POINTER(Element)]
        if hasattr(lib, "gcs_host_m3_solve"):
            lib.gcs_host_m3_solve.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                              C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
            lib.gcs_host_m3_rigid_transform.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
            lib.gcs_host_m3_score.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                              C.POINTER(C.c_uint8)]
            lib.gcs_host_m3_score.restype = C.c_double
        if hasattr(lib, "gcs_host_m3_merge"):
            ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
            lib.gcs_host_m3_merge.restype = C.c_int
            lib.gcs_host_m3_merge.argtypes = [C.c_int, C.c_int, ip, dp, ip, ip, dp, ip, dp, dp, C.POINTER(C.c_int64)]
        if hasattr(lib, "gcs_host_m3_ppp_merge"):
            ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
            lib.gcs_host_m3_ppp_merge.argtypes = [C.c_int, ip, dp, ip, ip, dp, ip, dp, dp, C.POINTER(C.c_int64)]
        _lib = lib
    return _lib


def last_error():
    return load().gcs_host_last_error().decode()


def to_c(elements, edges, el_type=Element, ed_type=Edge):
    els = (el_type * max(len(elements), 1))()
    for i, e in enumerate(elements):
        els[i].type = e["type"]
        els[i].is_set = 1 if e.get("is_set") else 0
        for j, v in enumerate(e["canvas"]):
            els[i].canvas[j] = v
        for j, v in enumerate(e.get("pos", [])):
            els[i].pos[j] = v
    eds = (ed_type * max(len(edges), 1))()
    for i, e in enumerate(edges):
        eds[i].a, eds[i].b, eds[i].type = e["a"], e["b"], e["type"]
        eds[i].flip = 1 if e.get("flip") else 0
        eds[i].value = e.get("value", 0.0)
    return els, eds


def from_c(elements, els):
    out = []
    for i, e in enumerate(elements):
        k = 2 if e["type"] == 0 else 4
        out.append({"type": e["type"], "is_set": bool(els[i].is_set), "pos": [els[i].pos[j] for j in range(k)]})
    return out


def component_solve(elements, edges):
    """classifyAndSolve on one leaf (CUDA path).  Returns (status, elements)."""
    els, eds = to_c(elements, edges)
    status = load().gcs_host_component_solve(len(elements), els, len(edges), eds)
    return status, from_c(elements, els)


def system_solve(elements, edges):
    els, eds = to_c(elements, edges)
    rc = load().gcs_host_system_solve(len(elements), els, len(edges), eds)
    return rc, from_c(elements, els)


def system_solve_ex(elements, edges):
    """Whole sketch through check -> decompose -> batched solveGcs.  Returns (rc, elements, stats dict)."""
    els, eds = to_c(elements, edges)
    stats = (C.c_int64 * 12)()
    rc = load().gcs_host_system_solve_ex(len(elements), els, len(edges), eds, stats)
    keys = ("leaves", "waves", "launches", "solved", "decompose_us", "solve_us", "plan_us", "pack_us", "device_us", "apply_us",
            "sharded_launches")
    return rc, from_c(elements, els), dict(zip(keys, list(stats)))


def decompose(elements, edges):
    """Decomposition only.  Returns (n_leaves, [(i, j, k)], virtual-edge counts, real-edge counts)."""
    els, eds = to_c(elements, edges)
    cap = max(len(elements), 1)
    le = (C.c_int32 * (3 * cap))()
    lv = (C.c_int32 * cap)()
    lr = (C.c_int32 * cap)()
    n = load().gcs_host_decompose(len(elements), els, len(edges), eds, le, lv, lr, cap)
    m = max(min(n, cap), 0)
    return n, [tuple(le[3 * i:3 * i + 3]) for i in range(m)], list(lv)[:m], list(lr)[:m]


def component_pack(elements, edges):
    """Packer only.  Returns (solver_id, kind, in[13], code, target_index, elements-with-anchors)."""
    els, eds = to_c(elements, edges)
    kind, target, code = C.c_int32(), C.c_int32(), C.c_uint8()
    row = (C.c_double * 13)()
    sid = load().gcs_host_component_pack(len(elements), els, len(edges), eds, C.byref(kind), row, C.byref(code), C.byref(target))
    return sid, kind.value, np.array(list(row)), code.value, target.value, from_c(elements, els)


def leaves_solve(elements, leaves, mode):
    """leaves: list of {"elems": [i, j, k], "edges": [edge dicts with element indices]}.
    mode 0 sequential classifyAndSolve, 1 batched solveGcs, 2 plan only.
    Returns dict(rc, status, level, solver, waves, launches, solved, elements)."""
    els, _ = to_c(elements, [])
    flat = []
    offs = [0]
    for lf in leaves:
        flat += lf["edges"]
        offs.append(len(flat))
    _, eds = to_c([], flat)
    n = len(leaves)
    le = (C.c_int32 * max(3 * n, 1))(*[i for lf in leaves for i in lf["elems"]])
    eo = (C.c_int32 * (n + 1))(*offs)
    status = (C.c_int32 * max(n, 1))()
    level = (C.c_int32 * max(n, 1))()
    solver = (C.c_int32 * max(n, 1))()
    stats = (C.c_int64 * 4)()
    rc = load().gcs_host_leaves_solve(len(elements), els, n, le, eo, eds, mode, status, level, solver, stats)
    return {"rc": rc, "status": list(status)[:n], "level": list(level)[:n], "solver": list(solver)[:n],
            "waves": stats[0], "launches": stats[1], "solved": stats[2], "plan_us": stats[3], "elements": from_c(elements, els)}


def solve2d(pair, params, guesses=None):
    p = (C.c_double * len(params))(*params)
    g = None if guesses is None else (C.c_double * 4)(*guesses)
    cand = (C.c_double * 4)()
    it = (C.c_int32 * 2)()
    cv = (C.c_int32 * 2)()
    rc = load().gcs_host_solve2d(pair, p, g, cand, it, cv)
    return rc, np.array(list(cand)).reshape(2, 2), list(it), list(cv)


# ---- bottom-up Merge3 numeric helpers ----
M3_WIDTH = {1: 12, 2: 14, 3: 16, 4: 20}
M3_OUT = {1: 2, 2: 4, 3: 2, 4: 2}
M3_KIND = {1: 1, 2: 2, 3: 3, 4: 4}  # equation-pair kind each case packs into


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def m3_solve(kase, rows, mode):
    """mode 1: one Merge3Batch; 2: single-call functions.  Returns (rc, out[n][M3_OUT], ok[n], launches)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    out = np.zeros((n, M3_OUT[kase]))
    ok = np.zeros(n, dtype=np.uint8)
    stats = (C.c_int64 * 1)()
    rc = load().gcs_host_m3_solve(kase, n, _dp(rows), _dp(out), ok.ctypes.data_as(C.POINTER(C.c_uint8)), mode, None, None, stats)
    return rc, out, ok, stats[0]


def m3_pack(kase, rows):
    """Pack only (no device).  Returns (rc, packed[n][13], code[n], needs_numerics[n])."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    out = np.zeros((n, M3_OUT[kase]))
    ok = np.zeros(n, dtype=np.uint8)
    packed = np.zeros((n, 13))
    code = np.zeros(n, dtype=np.uint8)
    rc = load().gcs_host_m3_solve(kase, n, _dp(rows), _dp(out), ok.ctypes.data_as(C.POINTER(C.c_uint8)), 0, _dp(packed),
                                  code.ctypes.data_as(C.POINTER(C.c_uint8)), None)
    return rc, packed, code, ok


def m3_rigid_transform(src, dst):
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    out = np.zeros(6)
    rc = load().gcs_host_m3_rigid_transform(src.shape[0], _dp(src), _dp(dst), _dp(out))
    return rc, out


def m3_score(types, canvas4, pose4, in_pose):
    types = np.ascontiguousarray(types, dtype=np.int32)
    canvas4 = np.ascontiguousarray(canvas4, dtype=np.float64)
    pose4 = np.ascontiguousarray(pose4, dtype=np.float64)
    in_pose = np.ascontiguousarray(in_pose, dtype=np.uint8)
    return load().gcs_host_m3_score(len(types), types.ctypes.data_as(C.POINTER(C.c_int32)), _dp(canvas4), _dp(pose4),
                                    in_pose.ctypes.data_as(C.POINTER(C.c_uint8)))


def _m3_ppp_args(types, canvas4, clusters):
    """clusters: three lists of (element id, pose4) in insertion order."""
    types = np.ascontiguousarray(types, dtype=np.int32)
    canvas4 = np.ascontiguousarray(canvas4, dtype=np.float64)
    counts = np.array([len(c) for c in clusters], dtype=np.int32)
    ids = np.array([i for c in clusters for i, _ in c], dtype=np.int32)
    pose4 = np.ascontiguousarray([p for c in clusters for _, p in c], dtype=np.float64).reshape(-1, 4)
    out_ids = np.zeros(len(types), dtype=np.int32)
    out_pose = np.zeros((len(types), 4))
    ip = C.POINTER(C.c_int32)
    return types, canvas4, counts, ids, pose4, out_ids, out_pose, ip


def m3_ppp_merge(types, canvas4, clusters):
    """Gcs::B200::solveMerge3Ppp.  Returns (n merged or 0 / -1, ids, pose4, score, (candidates, scored, launches))."""
    types, canvas4, counts, ids, pose4, out_ids, out_pose, ip = _m3_ppp_args(types, canvas4, clusters)
    score = C.c_double()
    stats = (C.c_int64 * 3)()
    n = load().gcs_host_m3_ppp_merge(len(types), types.ctypes.data_as(ip), _dp(canvas4), counts.ctypes.data_as(ip), ids.ctypes.data_as(ip),
                                     _dp(pose4), out_ids.ctypes.data_as(ip), _dp(out_pose), C.byref(score), stats)
    return n, out_ids[:max(n, 0)].copy(), out_pose[:max(n, 0)].copy(), score.value, tuple(stats)


M3_CASES = {"ppp": 0, "pll": 1, "lpp": 2, "llp": 3, "fallback": 4, "node": 5}


def m3_merge(which, types, canvas4, clusters):
    """Gcs::B200::solveMerge3{Ppp,Pll,Lpp,Llp,Fallback,Node}.  Returns (n merged or 0 / -1, ids, pose4, score,
    (candidates, scored, launches, case that produced the pose))."""
    types, canvas4, counts, ids, pose4, out_ids, out_pose, ip = _m3_ppp_args(types, canvas4, clusters)
    score = C.c_double()
    stats = (C.c_int64 * 4)()
    n = load().gcs_host_m3_merge(M3_CASES[which], len(types), types.ctypes.data_as(ip), _dp(canvas4), counts.ctypes.data_as(ip),
                                 ids.ctypes.data_as(ip), _dp(pose4), out_ids.ctypes.data_as(ip), _dp(out_pose), C.byref(score), stats)
    return n, out_ids[:max(n, 0)].copy(), out_pose[:max(n, 0)].copy(), score.value, tuple(stats)


def m3_level(nodes):
    """Gcs::B200::solveMerge3Level over `nodes` = [(types, canvas4, clusters), ...].  Returns (rc, [(n, ids, pose4, case)],
    (nodes, candidates, launches))."""
    ip = C.POINTER(C.c_int32)
    n_el = np.array([len(t) for t, _, _ in nodes], dtype=np.int32)
    types = np.ascontiguousarray(np.concatenate([np.zeros(0, dtype=np.int32)] + [np.asarray(t, dtype=np.int32) for t, _, _ in nodes]))
    canvas4 = np.ascontiguousarray(np.concatenate([np.zeros((0, 4))] + [np.asarray(c, dtype=np.float64).reshape(-1, 4) for _, c, _ in nodes]))
    counts = np.array([len(c) for _, _, cl in nodes for c in cl], dtype=np.int32)
    ids = np.array([i for _, _, cl in nodes for c in cl for i, _ in c], dtype=np.int32)
    pose4 = np.ascontiguousarray([p for _, _, cl in nodes for c in cl for _, p in c], dtype=np.float64).reshape(-1, 4)
    counts, ids = counts.astype(np.int32), ids.astype(np.int32)  # empty lists come out as float64
    out_n = np.zeros(len(nodes), dtype=np.int32)
    out_ids = np.zeros(len(types), dtype=np.int32)
    out_pose = np.zeros((len(types), 4))
    by = np.zeros(len(nodes), dtype=np.int32)
    stats = (C.c_int64 * 3)()
    lib = load()
    lib.gcs_host_m3_level.restype = C.c_int
    lib.gcs_host_m3_level.argtypes = [C.c_int, ip, ip, C.POINTER(C.c_double), ip, ip, C.POINTER(C.c_double), ip, ip, C.POINTER(C.c_double), ip,
                                      C.POINTER(C.c_int64)]
    rc = lib.gcs_host_m3_level(len(nodes), n_el.ctypes.data_as(ip), types.ctypes.data_as(ip), _dp(canvas4), counts.ctypes.data_as(ip),
                               ids.ctypes.data_as(ip), _dp(pose4), out_n.ctypes.data_as(ip), out_ids.ctypes.data_as(ip), _dp(out_pose),
                               by.ctypes.data_as(ip), stats)
    res, at = [], 0
    for k, n in enumerate(out_n):
        res.append((int(n), out_ids[at:at + n].copy(), out_pose[at:at + n].copy(), int(by[k])))
        at += int(n_el[k])
    return rc, res, tuple(stats)


def canvas_transform(elements):
    """Solver -> canvas rigid motion over elements carrying is_set/pos.  Returns (rc, canvas list)."""
    els, _ = to_c(elements, [])
    rc = load().gcs_host_canvas_transform(len(elements), els)
    return rc, [[els[i].canvas[j] for j in range(2 if e["type"] == 0 else 4)] for i, e in enumerate(elements)]
