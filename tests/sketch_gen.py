"""Synthetic multi-leaf sketches for the leaf scheduler tests: a Henneberg-I style construction
(every new element hangs on two already placed ones), mixed points and lines, emitted as the
leaf list DeficitStreeBasedTopDownStrategy::solveGcs receives (right shape for each of the
eight sub-problem solvers).  Element / edge dicts follow host_lib / ref_lib."""
import math

import numpy as np

P, L = 0, 1
DIST, ANG, VIRT = 0, 1, 2


def _pdist(p, ln):
    ex, ey = ln[2] - ln[0], ln[3] - ln[1]
    return abs(ex * (p[1] - ln[1]) - ey * (p[0] - ln[0])) / math.hypot(ex, ey)


def _angle(l1, l2):
    d1 = (l1[2] - l1[0], l1[3] - l1[1])
    d2 = (l2[2] - l2[0], l2[3] - l2[1])
    c = (d1[0] * d2[0] + d1[1] * d2[1]) / (math.hypot(*d1) * math.hypot(*d2))
    return math.acos(max(-1.0, min(1.0, c)))


def _centre(e):
    return (e[0], e[1]) if len(e) == 2 else ((e[0] + e[2]) / 2, (e[1] + e[3]) / 2)


def make_sketch(n_leaves, seed=1, first_shape=1, p_line=0.35, locality=None):
    """Returns (elements, leaves).  locality=None picks parents anywhere (dependency depth
    ~ log n); an integer k picks them among the last k elements (deeper chains)."""
    rng = np.random.default_rng(seed)
    true = []

    def rpoint(c=(0.0, 0.0), r=300.0):
        return [float(c[0] + rng.uniform(-r, r)), float(c[1] + rng.uniform(-r, r))]

    def rline(c=(0.0, 0.0), r=300.0):
        a = rpoint(c, r)
        th = rng.uniform(0, 2 * math.pi)
        ln = rng.uniform(40, 400)
        return a + [a[0] + ln * math.cos(th), a[1] + ln * math.sin(th)]

    leaves = []

    def edge(a, b, typ, value=0.0, flip=False):
        return {"a": int(a), "b": int(b), "type": typ, "value": float(value), "flip": bool(flip)}

    def emit(ids, edges):
        ids = [int(i) for i in rng.permutation(ids)]
        edges = [edges[i] for i in rng.permutation(len(edges))]
        leaves.append({"elems": ids, "edges": edges})

    if first_shape == 1:
        true += [rpoint(), rpoint(), rpoint()]
        emit([0, 1, 2], [edge(0, 1, DIST, math.dist(true[0], true[1])), edge(0, 2, DIST, math.dist(true[0], true[2])),
                         edge(1, 2, DIST, math.dist(true[1], true[2]))])
    elif first_shape == 2:
        true += [rpoint(), rpoint(), rline()]
        emit([0, 1, 2], [edge(0, 1, DIST, math.dist(true[0], true[1])), edge(0, 2, DIST, _pdist(true[0], true[2])),
                         edge(1, 2, DIST, _pdist(true[1], true[2]))])
    else:
        true += [rline(), rline(), rpoint()]
        emit([0, 1, 2], [edge(0, 1, ANG, _angle(true[0], true[1]), rng.uniform() < 0.5),
                         edge(2, 0, DIST, _pdist(true[2], true[0])), edge(2, 1, DIST, _pdist(true[2], true[1]))])

    while len(leaves) < n_leaves:
        n = len(true)
        lo = 0 if locality is None else max(0, n - locality)
        a, b = (int(v) for v in rng.choice(np.arange(lo, n), size=2, replace=False))
        ta, tb = len(true[a]) == 4, len(true[b]) == 4
        ca, cb = _centre(true[a]), _centre(true[b])
        mid = ((ca[0] + cb[0]) / 2, (ca[1] + cb[1]) / 2)
        want_line = rng.uniform() < p_line
        new = n
        between = rng.uniform()
        if not ta and not tb:
            if want_line:  # shape 5
                ln = rline(mid, 150)
                true.append(ln)
                es = [edge(a, new, DIST, _pdist(true[a], ln)), edge(b, new, DIST, _pdist(true[b], ln))]
            else:          # shape 4
                pt = rpoint(mid, 150)
                true.append(pt)
                es = [edge(a, new, DIST, math.dist(true[a], pt)), edge(b, new, DIST, math.dist(true[b], pt))]
            if between < 0.5:
                es.append(edge(a, b, VIRT))
            elif between < 0.75:
                es.append(edge(a, b, DIST, math.dist(true[a], true[b])))
        elif ta and tb:      # shape 7: two fixed lines, free point
            if abs(math.sin(_angle(true[a], true[b]))) < 0.2:
                continue
            pt = rpoint(mid, 150)
            true.append(pt)
            es = [edge(a, new, DIST, _pdist(pt, true[a])), edge(b, new, DIST, _pdist(pt, true[b]))]
            if between < 0.5:
                es.append(edge(a, b, VIRT))
        else:
            pnt, lin = (b, a) if ta else (a, b)
            if want_line:  # shape 8: fixed line + fixed point, free line (angle + distance)
                ln = rline(mid, 150)
                true.append(ln)
                es = [edge(lin, new, ANG, _angle(true[lin], ln), rng.uniform() < 0.5), edge(pnt, new, DIST, _pdist(true[pnt], ln))]
                if between < 0.5:
                    es.append(edge(pnt, lin, VIRT))
            else:          # shape 6: fixed point + fixed line, free point
                pt = rpoint(mid, 150)
                if _pdist(pt, true[lin]) < 1.0:
                    continue
                true.append(pt)
                es = [edge(pnt, new, DIST, math.dist(true[pnt], pt)), edge(lin, new, DIST, _pdist(pt, true[lin]))]
                if between < 0.5:
                    es.append(edge(pnt, lin, VIRT))
                elif between < 0.75:
                    es.append(edge(pnt, lin, DIST, _pdist(true[pnt], true[lin])))
        emit([a, b, new], es)

    # the canvas: the true layout moved rigidly (sometimes mirrored) with a little noise
    th = rng.uniform(0, 2 * math.pi)
    c, s = math.cos(th), math.sin(th)
    t = rng.uniform(0, 1000, size=2)
    m = -1.0 if rng.uniform() < 0.3 else 1.0
    elements = []
    for e in true:
        cv = []
        for k in range(0, len(e), 2):
            x, y = e[k], e[k + 1] * m
            cv += [c * x - s * y + t[0] + rng.normal(0, 1.0), s * x + c * y + t[1] + rng.normal(0, 1.0)]
        elements.append({"type": L if len(e) == 4 else P, "canvas": [float(v) for v in cv]})
    return elements, leaves


def sketch_graph(elements, leaves):
    """The whole-sketch constraint graph of a generated leaf list: the three constraints of the
    first leaf and the two constraints that hang each later element on its parents (the redundant
    parent-parent edges some leaves repeat are dropped), no virtual edges: 2n - 3 constraints, what
    a user would draw; decomposing it is the solver's job."""
    seen, edges = set(), []
    for k, lf in enumerate(leaves):
        new = max(lf["elems"])
        for e in lf["edges"]:
            if e["type"] == VIRT or (k > 0 and new not in (e["a"], e["b"])):
                continue
            key = (min(e["a"], e["b"]), max(e["a"], e["b"]))
            if key in seen:
                continue
            seen.add(key)
            edges.append(dict(e))
    return edges


def make_linkage(n_points, seed=1):
    """BASELINE config 4: a rigidly well-constrained linkage of points and distances (2n - 3
    constraints), every new point hung on two earlier ones.  Returns (elements, edges)."""
    el, lv = make_sketch(n_points - 2, seed=seed, first_shape=1, p_line=0.0)
    return el, sketch_graph(el, lv)
