"""Bottom-up Merge3 numeric helpers: N candidate sub-problems per case
through Gcs::B200::Merge3Batch (packing on the host + ONE launch per kind through the host-buffer
C ABI) next to the reference's own helper functions called in a loop (oracle/_ref, one thread -
the Merge3 solvers are single threaded).  Usage: python profiles/merge3_bench.py [n_candidates]"""
import importlib
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
importlib.import_module("2d_geometry_constraint_solver_b200").capi.init([0])
import host_lib as H  # noqa: E402

spec = importlib.util.spec_from_file_location("mg3", os.path.join(ROOT, "oracle", "make_golden_merge3.py"))
mg = importlib.util.module_from_spec(spec)
try:
    spec.loader.exec_module(mg)   # imports ref_lib: fine where oracle/_ref travelled
    import ref_lib as R
    have_ref = R.available()
except Exception:
    have_ref = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mg.N = n
rng = np.random.default_rng(9)
names = {1: "point from two points (K1)", 2: "line from two points (K2)", 3: "point from point+line (K3)", 4: "point from two lines (K4)"}
out = {"workload": f"{n} Merge3 candidates per case", "cases": {}}
H.m3_solve(1, mg.rows_pp(np.random.default_rng(1))[:64], 1)  # warm-up (context, arena)
for kase, gen in ((1, mg.rows_pp), (2, mg.rows_line), (3, mg.rows_pl), (4, mg.rows_ll)):
    rows = gen(rng)
    if kase == 4:
        rows = rows[np.hypot(rows[:, 2] - rows[:, 0], rows[:, 3] - rows[:, 1]) >= 1e-9]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        rc, got, ok, launches = H.m3_solve(kase, rows, 1)
        dt = time.perf_counter() - t0
        assert rc == 0, H.last_error()
        best = dt if best is None else min(best, dt)
    rec = {"candidates": len(rows), "launches": int(launches), "batch_s": best, "candidates_per_s": len(rows) / best}
    if have_ref:
        devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(2)
        os.dup2(devnull, 2)
        t0 = time.perf_counter()
        exp, exp_ok = R.m3_solve(kase, rows)
        ref_s = time.perf_counter() - t0
        os.dup2(saved, 2)
        same = np.array_equal(ok, exp_ok) and bool(np.all((got[exp_ok == 1].view(np.uint64) == exp[exp_ok == 1].view(np.uint64))))
        rec.update(reference_loop_s=ref_s, reference_candidates_per_s=len(rows) / ref_s, identical_to_reference=same)
    out["cases"][names[kase]] = rec
print(json.dumps(out))
