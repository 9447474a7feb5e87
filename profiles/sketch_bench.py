"""BASELINE config 4: one large sketch (a rigidly well-constrained linkage of N points) through
GeometricConstraintSystem -> peel decomposition -> batched GPU solve, next to the reference
build's sequential leaf loop (oracle/_ref) on the same leaves.
Usage: python profiles/sketch_bench.py [n_points]"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
importlib.import_module("2d_geometry_constraint_solver_b200").capi.init([0])
import host_lib as H  # noqa: E402
import sketch_gen as S  # noqa: E402
from test_gpu_host import _leaf_dicts  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
el, edges = S.make_linkage(n, seed=4)
H.system_solve_ex(el[:2000], [e for e in edges if max(e["a"], e["b"]) < 2000])  # warm-up (context, arena)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    rc, got, stats = H.system_solve_ex(el, edges)
    wall = time.perf_counter() - t0
    assert rc == 0, H.last_error()
    if best is None or stats["solve_us"] < best["solve_us"]:
        best = dict(stats, wall_s=wall)
out = {"workload": f"configs[3]: linkage of {n} points, {len(edges)} distance constraints", **best,
       "leaves_per_s_solve": best["leaves"] / (best["solve_us"] * 1e-6),
       "leaves_per_s_incl_decomposition": best["leaves"] / ((best["solve_us"] + best["decompose_us"]) * 1e-6)}
try:
    import ref_lib as R
    if R.available():
        nl, leaves, _, _ = H.decompose(el, edges)
        lv = _leaf_dicts(el, edges, leaves)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)   # the reference prints two lines per leaf to stderr
        t0 = time.perf_counter()
        rc, status, exp = R.leaves_solve(el, lv)
        ref_s = time.perf_counter() - t0
        os.dup2(saved, 2)
        same = all(a["pos"] == b["pos"] for a, b in zip(got, exp))
        out["reference_loop_s"] = ref_s
        out["reference_leaves_per_s"] = len(lv) / ref_s
        out["identical_to_reference_loop"] = bool(same)
except Exception as ex:  # pragma: no cover
    out["reference_loop_error"] = str(ex)
print(json.dumps(out))
