#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the reference's own sources
(oracle/_ref/libgcs_ref.so, built by oracle/build_ref.sh from /root/reference).  Run in the
build container (where /root/reference exists):

    python oracle/make_golden.py

  tests/golden/numeric_k{1..5}.npz  K-input batches (synth generators + hand-made edge cases)
      with what the reference's solve2D + heuristics return: candidates, recovered iteration
      counts, chosen root, chosen point / (normal, offset).
  tests/golden/components.json      3-element leaf components for all eight sub-problem solvers
      (plus unsupported shapes) with the element positions the reference's classifyAndSolve
      leaves behind.

The vectors pin everything the reference owns (loop semantics, formulas, heuristics, role
assignment, anchoring, sign conventions, line reconstruction, dispatch order).  The arithmetic
inside Eigen / autodiff is the stand-ins' (oracle/ref_shim), i.e. restated - see DESIGN.md.
"""
import importlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
import ref_lib as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_NUMERIC = 768


def numeric():
    capi, synth = gcs.capi, gcs.synth
    for kind in (1, 2, 3, 4, 5):
        hb = synth.make(kind, N_NUMERIC, seed=0x601D0000 + kind)
        if kind == 1:
            # hand-made rows: 3-4-5, equilateral 100, flat and collinear-canvas cases
            c = hb.cols
            rows = [(0, 0, 4, 3, 0, 5, 1), (0, 0, 4, 3, 0, 5, -1), (0, 0, 100, 100, 0, 100, 1),
                    (0, 0, 4, 3, 0, 5, 0), (0, 0, 1.5, 3, 0, 1.5, 1), (10, 10, 5, 10, 10, 5, 1)]
            for j, r in enumerate(rows):
                for k in range(6):
                    c[k][j] = r[k]
                hb.code[j] = capi.make_code(np.array([r[6]]))[0]
        R.solve_batch(hb.alloc_outputs(), count_iters=True)
        np.savez_compressed(
            os.path.join(GOLD, f"numeric_k{kind}.npz"), kind=kind, cols=np.stack(hb.cols), code=hb.code,
            cand=hb.cand, iters=hb.iters, converged=hb.converged, root=hb.root_index, out=np.stack(hb.out))
        print("numeric", kind, "iters", hb.iters.min(), hb.iters.max(), "root2", int((hb.root_index == 2).sum()))


# ---------------------------------------------------------------------------------------------
def rigid(rng, mirror=False):
    th = rng.uniform(0, 2 * math.pi)
    c, s = math.cos(th), math.sin(th)
    t = rng.uniform(0, 1000, size=2)
    m = -1.0 if mirror else 1.0

    def f(p):
        x, y = p[0], p[1] * m
        return [c * x - s * y + t[0], s * x + c * y + t[1]]
    return f


def pdist(p, a, b):
    ex, ey = b[0] - a[0], b[1] - a[1]
    return abs(ex * (p[1] - a[1]) - ey * (p[0] - a[0])) / math.hypot(ex, ey)


def rpoint(rng, lo=-300, hi=300):
    return [float(rng.uniform(lo, hi)), float(rng.uniform(lo, hi))]


def rline(rng):
    a = rpoint(rng)
    th = rng.uniform(0, 2 * math.pi)
    ln = rng.uniform(40, 400)
    return a + [a[0] + ln * math.cos(th), a[1] + ln * math.sin(th)]


def canvas_of(rng, true_elems):
    """canvas = the true configuration moved rigidly (sometimes mirrored) plus a little noise"""
    f = rigid(rng, mirror=rng.uniform() < 0.3)
    out = []
    for e in true_elems:
        pts = [e[0:2], e[2:4]] if len(e) == 4 else [e]
        cv = []
        for p in pts:
            q = f(p)
            cv += [q[0] + rng.normal(0, 2.0), q[1] + rng.normal(0, 2.0)]
        out.append(cv)
    return out


def make_component(rng, shape):
    """Returns (elements, edges) for one of the eight solver shapes (true geometry -> constraints)."""
    P, L = 0, 1
    DIST, ANG, VIRT = 0, 1, 2
    if shape in (1, 4):
        t = [rpoint(rng) for _ in range(3)]
        types = [P, P, P]
    elif shape in (2, 5, 6):
        t = [rpoint(rng), rpoint(rng), rline(rng)]
        types = [P, P, L]
    else:
        t = [rline(rng), rline(rng), rpoint(rng)]
        types = [L, L, P]
    cv = canvas_of(rng, t)
    order = list(rng.permutation(3))
    inv = {int(o): i for i, o in enumerate(order)}  # true index -> position in the element list
    solved = set()
    if shape == 4:
        solved = {0, 1}
    elif shape == 5:
        solved = {0, 1}
    elif shape == 6:
        solved = {0, 2}
    elif shape == 7:
        solved = {0, 1}
    elif shape == 8:
        solved = {0, 2}
    elements = []
    for o in order:
        o = int(o)
        e = {"type": types[o], "canvas": [float(v) for v in cv[o]]}
        if o in solved:
            # solver space = another rigid placement of the true geometry would also do; use the
            # true coordinates themselves
            e["is_set"] = True
            e["pos"] = [float(v) for v in t[o]]
        elements.append(e)

    def edge(a, b, typ, value=0.0, flip=False):
        return {"a": inv[a], "b": inv[b], "type": typ, "value": float(value), "flip": bool(flip)}

    def between_solved(a, b, value):
        return edge(a, b, VIRT) if rng.uniform() < 0.6 else edge(a, b, DIST, value)

    edges = []
    if shape == 1:
        edges = [edge(0, 1, DIST, math.dist(t[0], t[1])), edge(0, 2, DIST, math.dist(t[0], t[2])),
                 edge(1, 2, DIST, math.dist(t[1], t[2]))]
    elif shape == 2:
        ln = t[2]
        edges = [edge(0, 1, DIST, math.dist(t[0], t[1])), edge(0, 2, DIST, pdist(t[0], ln[:2], ln[2:])),
                 edge(1, 2, DIST, pdist(t[1], ln[:2], ln[2:]))]
    elif shape in (3, 8):
        l1, l2, p = t
        d1 = [l1[2] - l1[0], l1[3] - l1[1]]
        d2 = [l2[2] - l2[0], l2[3] - l2[1]]
        ang = math.acos(max(-1.0, min(1.0, (d1[0] * d2[0] + d1[1] * d2[1]) / (math.hypot(*d1) * math.hypot(*d2)))))
        flip = rng.uniform() < 0.5
        if shape == 3:
            edges = [edge(0, 1, ANG, ang, flip), edge(2, 0, DIST, pdist(p, l1[:2], l1[2:])),
                     edge(2, 1, DIST, pdist(p, l2[:2], l2[2:]))]
        else:  # line 0 and the point are solved, line 1 is free
            edges = [edge(0, 1, ANG, ang, flip), edge(2, 1, DIST, pdist(p, l2[:2], l2[2:])), edge(2, 0, VIRT)]
    elif shape == 4:
        edges = [edge(0, 2, DIST, math.dist(t[0], t[2])), edge(1, 2, DIST, math.dist(t[1], t[2])),
                 between_solved(0, 1, math.dist(t[0], t[1]))]
    elif shape == 5:
        ln = t[2]
        edges = [edge(0, 2, DIST, pdist(t[0], ln[:2], ln[2:])), edge(1, 2, DIST, pdist(t[1], ln[:2], ln[2:])),
                 between_solved(0, 1, math.dist(t[0], t[1]))]
    elif shape == 6:  # point 0 and the line are solved, point 1 is free
        ln = t[2]
        edges = [edge(0, 1, DIST, math.dist(t[0], t[1])), edge(2, 1, DIST, pdist(t[1], ln[:2], ln[2:])),
                 between_solved(0, 2, pdist(t[0], ln[:2], ln[2:]))]
    elif shape == 7:
        l1, l2, p = t
        edges = [edge(0, 2, DIST, pdist(p, l1[:2], l1[2:])), edge(1, 2, DIST, pdist(p, l2[:2], l2[2:])), edge(0, 1, VIRT)]
    rng.shuffle(edges)
    return elements, edges


def components():
    rng = np.random.default_rng(20261018)
    items = []
    # the BASELINE configs[0] sketch: points (100,100), (200,100), (150,200); distances 3,4,5
    for dists in ((3.0, 4.0, 5.0), (100.0, 100.0, 100.0)):
        el = [{"type": 0, "canvas": [100.0, 100.0]}, {"type": 0, "canvas": [200.0, 100.0]}, {"type": 0, "canvas": [150.0, 200.0]}]
        ed = [{"a": 0, "b": 1, "type": 0, "value": dists[0], "flip": False}, {"a": 0, "b": 2, "type": 0, "value": dists[1], "flip": False},
              {"a": 1, "b": 2, "type": 0, "value": dists[2], "flip": False}]
        items.append({"shape": 1, "elements": el, "edges": ed})
    for shape in range(1, 9):
        for _ in range(96):
            el, ed = make_component(rng, shape)
            items.append({"shape": shape, "elements": el, "edges": ed})
    # unsupported shapes: three lines; two solved lines + point with an angle edge between the lines
    el = [{"type": 1, "canvas": rline(rng)} for _ in range(3)]
    ed = [{"a": 0, "b": 1, "type": 1, "value": 0.5, "flip": False}, {"a": 1, "b": 2, "type": 1, "value": 0.7, "flip": False},
          {"a": 0, "b": 2, "type": 1, "value": 0.9, "flip": False}]
    items.append({"shape": 0, "elements": el, "edges": ed})
    el, ed = make_component(rng, 7)
    for e in ed:
        if e["type"] == 2:
            e["type"], e["value"] = 1, 0.4
    items.append({"shape": 0, "elements": el, "edges": ed})
    for it in items:
        status, out = R.component_solve(it["elements"], it["edges"])
        it["status"] = status
        it["expected"] = out
    with open(os.path.join(GOLD, "components.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "items": items}, f)
    by = {}
    for it in items:
        by.setdefault((it["shape"], it["status"]), 0)
        by[(it["shape"], it["status"])] += 1
    print("components:", sorted(by.items()))


def sketches():
    """Multi-leaf sketches (shared elements, dependency chains) through the reference's
    sequential leaf loop: pins the batched scheduler's final element state."""
    import sketch_gen as S
    items = []
    for seed, n, first, loc in ((11, 300, 1, None), (12, 300, 2, None), (13, 300, 3, None), (14, 120, 1, 6)):
        el, lv = S.make_sketch(n, seed=seed, first_shape=first, locality=loc)
        rc, status, out = R.leaves_solve(el, lv)
        items.append({"seed": seed, "first_shape": first, "locality": loc, "elements": el, "leaves": lv, "rc": rc,
                      "status": status, "expected": out})
        print("sketch", seed, "leaves", n, "rc", rc, "statuses", sorted(set(status)),
              "set", sum(e["is_set"] for e in out), "of", len(out))
    with open(os.path.join(GOLD, "sketch_leaves.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "items": items}, f)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if "--sketches-only" not in sys.argv:
        numeric()
        components()
    sketches()
