"""The Merge3 enumeration loops as batch consumers: one merge of three clusters with
k shared elements per group, through Gcs::B200::solveMerge3{Ppp,Pll,Lpp,Llp} / solveMerge3Node (pass 1 packs every
candidate, ONE launch per kind, pass 2 places / merges / scores) next to the reference's own Merge3*Solver::solve
(oracle/_ref, its per-candidate progress lines to /dev/null).  Wall clock around the C entry points, best of 3; both
sides build the same graph and poses from the same arrays first.  Usage: python profiles/merge3_node_bench.py [k]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
importlib.import_module("2d_geometry_constraint_solver_b200").capi.init([0])
import host_lib as H  # noqa: E402
import ref_lib as R  # noqa: E402
import test_merge3 as T  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
specs = {"ppp": dict(ra=(k, 0), rb=(k, 0), f=(k, 0), r=(2, 2), a=(2, 1), b=(1, 2)),
         "pll": dict(ra=(k, 0), rb=(k, 0), f=(0, k), r=(2, 2), a=(2, 1), b=(1, 2)),
         "lpp": dict(ra=(k, 0), rb=(0, k), f=(k, 0), r=(2, 2), a=(2, 1), b=(1, 2)),
         "llp": dict(ra=(0, k), rb=(0, k), f=(k, 0), r=(2, 2), a=(2, 1), b=(1, 2))}
rng = np.random.default_rng(12)
H.m3_merge("ppp", *T._m3_scenario(np.random.default_rng(1), dict(ra=(1, 0), rb=(1, 0), f=(1, 0))))  # warm-up (context, arena)
out = {"workload": f"one Merge3 node, {k} shared elements per group", "cases": {}}
for case, spec in specs.items():
    types, canvas4, clusters = T._m3_scenario(rng, spec)
    for which in (case, "node"):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            n, ids, pose, score, (cands, scored, launches, by) = H.m3_merge(which, types, canvas4, clusters)
            dt = time.perf_counter() - t0
            assert n >= 0, H.last_error()
            best = dt if best is None else min(best, dt)
        rec = {"elements": int(len(types)), "candidates": int(cands), "scored": int(scored), "launches": int(launches), "batched_ms": best * 1e3}
        if R.available() and hasattr(R.load(), "gcs_ref_m3_merge"):
            with T._quiet_stderr():
                rbest = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    n_ref, ids_ref, pose_ref, by_ref = R.m3_merge(which, types, canvas4, clusters)
                    dt = time.perf_counter() - t0
                    rbest = dt if rbest is None else min(rbest, dt)
            rec.update(reference_ms=rbest * 1e3, speed_up=rbest / best,
                       identical_to_reference=bool(n == n_ref and np.array_equal(ids, ids_ref) and T.same(pose, pose_ref).all()))
        out["cases"][f"{case} shape through {'solveMerge3Node' if which == 'node' else 'solveMerge3' + case.capitalize()}"] = rec
print(json.dumps(out))
