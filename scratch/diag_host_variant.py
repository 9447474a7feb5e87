"""Leaf by leaf: pack each leaf of a sketch on the current element state, solve the row at the C ABI with
variant 0 and variant 5, report the first leaf where the two disagree."""
import importlib, os, sys
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
import host_lib as H, sketch_gen as S
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi = gcs.capi
capi.init([0])
el, lv = S.make_sketch(300, seed=11, first_shape=2)
state = [dict(e) for e in el]
nbad = 0
for li, lf in enumerate(lv):
    ids = lf["elems"]
    sub = [dict(state[i]) for i in ids]
    loc = {g: l for l, g in enumerate(ids)}
    edges = [dict(e, a=loc[e["a"]], b=loc[e["b"]]) for e in lf["edges"]]
    sid, kind, row, code, target, _ = H.component_pack(sub, edges)
    if sid <= 0:
        continue
    res = {}
    for v in (0, 5):
        hb = capi.HostBatch(kind, 2, [np.array([row[c]]) for c in range(capi.IN_COLS[kind])], np.array([code], np.uint8), None, v, want_cand=True)
        capi.solve_host(hb.alloc_outputs(), 0)
        res[v] = hb
    a, b = res[0], res[5]
    same = np.array_equal(a.iters, b.iters) and np.array_equal(a.converged, b.converged) and np.array_equal(a.root_index, b.root_index)
    err = max(float(np.max(np.abs(x - y))) for x, y in zip(a.out, b.out))
    if not same or not err <= 1e-6:
        nbad += 1
        if nbad <= 4:
            print("leaf", li, "solver", sid, "kind", kind, "code", code, "row", list(row[:capi.IN_COLS[kind]]))
            print("   v0 iters", a.iters.ravel(), "conv", a.converged.ravel(), "root", a.root_index, "cand", a.cand.ravel(), "out", [float(o[0]) for o in a.out])
            print("   v5 iters", b.iters.ravel(), "conv", b.converged.ravel(), "root", b.root_index, "cand", b.cand.ravel(), "out", [float(o[0]) for o in b.out])
    if not a.converged.all():
        print("leaf", li, "kind", kind, "conv", a.converged.ravel(), "iters", a.iters.ravel(), "root", a.root_index, "-> selected candidate converged:", bool(a.converged.ravel()[int(a.root_index[0])]), "target element", ids[target])
    st, new = H.component_solve(sub, edges)
    for l, g in enumerate(ids):
        state[g].update(new[l])
print("leaves", len(lv), "disagreeing", nbad)
