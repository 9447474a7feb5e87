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


def make_linkage_unchecked(n_points, seed=1):
    """A linkage of points and distances (2n - 3 constraints), every new point hung on two earlier
    ones, drawn without looking at what the reference's solver does with it: about one leaf in
    seventy has both Newton seeds on the same side of its base line, the reference's heuristic then
    returns the mirrored root unchecked (heuristics.hpp:56), and everything hanging below has circles
    that no longer meet (runs to the iteration cap).  Kept as the stress case of the scheduler."""
    el, lv = make_sketch(n_points - 2, seed=seed, first_shape=1, p_line=0.0)
    return el, sketch_graph(el, lv)


def _newton_pp(ax, ay, ra, bx, by, rb, gx, gy, iters=80):
    """Plain Newton on two circles from one seed, vectorised (numpy; not the checker and not the
    product: the generator only needs to know WHICH root a seed reaches, and keeps away from the
    cases where that could depend on rounding)."""
    x = np.full_like(ax, gx)
    y = np.full_like(ax, gy)
    live = np.ones(ax.shape, dtype=bool)
    count = np.zeros(ax.shape, dtype=np.int64)
    wmin = np.full_like(ax, np.inf)  # closest approach (in |det J| / 4) to the line where J is singular
    with np.errstate(all="ignore"):
        for _ in range(iters):
            dxa, dya, dxb, dyb = x - ax, y - ay, x - bx, y - by
            f = dxa * dxa + dya * dya - ra * ra
            g = dxb * dxb + dyb * dyb - rb * rb
            det = dxa * dyb - dya * dxb            # det(J) / 4
            sx = -(f * dyb - g * dya) / (2.0 * det)
            sy = -(dxa * g - dxb * f) / (2.0 * det)
            wmin = np.where(live, np.minimum(wmin, np.abs(det)), wmin)
            x = np.where(live, x + sx, x)
            y = np.where(live, y + sy, y)
            count += live
            live &= ~((np.abs(sx) < 1e-9) & (np.abs(sy) < 1e-9))
            if not live.any():
                break
    return x, y, count, ~live, wmin


def make_linkage(n_points, seed=1, max_rounds=60):
    """BASELINE config 4: a rigidly well-constrained linkage of n_points points and 2n - 3 distances,
    every new point hung on two earlier ones (dependency depth ~ log n), drawn on a noisy canvas -
    and SOLVABLE BY THE REFERENCE: a point is redrawn until the reference's own rule
    (pickByTriangleOrientation, heuristics.hpp:46-57: candidate 0 if its orientation matches the
    canvas, else candidate 1 unchecked) returns the root that agrees with the sketch, with margins
    (no near-collinear triple, both seeds converge, neither passes near the singular line).  So the
    sketch stays rigid: every constraint holds in the solved result.
    Needs the leaf list the product's decomposition emits (host_lib.decompose; structure only).
    Returns (elements, edges)."""
    import host_lib as H
    rng = np.random.default_rng(seed)
    n = int(n_points)
    assert n >= 3
    # ---- structure: parents of every point, levels ----
    v = np.arange(n)
    pa = np.zeros(n, dtype=np.int64)
    pb = np.zeros(n, dtype=np.int64)
    u = rng.random((2, n))
    pa[3:] = np.floor(u[0, 3:] * v[3:]).astype(np.int64)
    pb[3:] = np.floor(u[1, 3:] * (v[3:] - 1)).astype(np.int64)
    pb[3:] += pb[3:] >= pa[3:]
    level = np.zeros(n, dtype=np.int64)
    for i in range(3, n):
        level[i] = 1 + max(level[pa[i]], level[pb[i]])
    ea = np.concatenate([[0, 0, 1], pa[3:], pb[3:]])
    eb = np.concatenate([[1, 2, 2], v[3:], v[3:]])
    # ---- exact layout, level by level ----
    E = np.zeros((n, 2))
    E[:3] = rng.uniform(-300, 300, size=(3, 2))

    def draw(idx):
        mid = 0.5 * (E[pa[idx]] + E[pb[idx]])
        E[idx] = mid + rng.uniform(-150, 150, size=(len(idx), 2))
    for lv in range(1, int(level.max()) + 1):
        draw(np.nonzero(level == lv)[0])
    th = rng.uniform(0, 2 * math.pi)
    c, s_ = math.cos(th), math.sin(th)
    t = rng.uniform(0, 1000, size=2)
    noise = rng.normal(0, 1.0, size=(n, 2))

    def canvas():
        return np.stack([c * E[:, 0] - s_ * E[:, 1] + t[0], s_ * E[:, 0] + c * E[:, 1] + t[1]], axis=1) + noise

    def edge_dicts():
        val = np.hypot(*(E[ea] - E[eb]).T)
        return [{"a": int(a), "b": int(b), "type": DIST, "value": float(d), "flip": False} for a, b, d in zip(ea, eb, val)]

    # ---- the leaves the product's decomposition emits (depends on the structure only) ----
    C0 = canvas()
    nl, triples, _, _ = H.decompose([{"type": P, "canvas": [float(x), float(y)]} for x, y in C0], edge_dicts())
    assert nl == n - 2, (nl, H.last_error())
    tri = np.array(triples, dtype=np.int64)
    placed = np.zeros(n, dtype=bool)
    placed[tri[0]] = True
    free = np.zeros(nl, dtype=np.int64)
    fix = np.zeros((nl, 2), dtype=np.int64)
    for k in range(1, nl):
        m = ~placed[tri[k]]
        assert m.sum() == 1, "a leaf of a Henneberg-I linkage has exactly one new point"
        free[k] = tri[k][m][0]
        fix[k] = np.sort(tri[k][~m])  # fixed1, fixed2 = solved points in ascending node id (point_point_solvers.cpp:110-123)
        placed[free[k]] = True
    p1, p2, p3 = np.sort(tri[0])  # ZeroFixedPoints: P1 -> (0, 0), P2 -> (d12, 0) (point_point_solvers.cpp:48-50)
    G = 20000.0

    def ori(a, b, q):
        return (b[:, 0] - a[:, 0]) * (q[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (q[:, 0] - a[:, 0])

    for rnd in range(max_rounds):
        C = canvas()
        d12 = E[p2] - E[p1]
        ux, uy = d12 / np.hypot(*d12)
        rel = E - E[p1]
        SP = np.stack([rel[:, 0] * ux + rel[:, 1] * uy, -rel[:, 0] * uy + rel[:, 1] * ux], axis=1)  # the solver's frame
        # anchor leaf first: an anchored triangle always has its seeds on opposite sides of y = 0
        tri0 = np.array([[p1, p2, p3]])
        bad = np.zeros(n, dtype=bool)
        k = np.arange(1, nl)
        A, B, Q = SP[fix[k, 0]], SP[fix[k, 1]], SP[free[k]]
        ra, rb = np.hypot(*(E[free[k]] - E[fix[k, 0]]).T), np.hypot(*(E[free[k]] - E[fix[k, 1]]).T)
        csign = np.sign(ori(C[fix[k, 0]], C[fix[k, 1]], C[free[k]]))
        co = np.abs(ori(C[fix[k, 0]], C[fix[k, 1]], C[free[k]]))
        flat = co < 0.05 * np.hypot(*(C[fix[k, 1]] - C[fix[k, 0]]).T) * np.hypot(*(C[free[k]] - C[fix[k, 0]]).T)
        flat |= np.sign(ori(A, B, Q)) != csign
        x0, y0, n0, ok0, w0 = _newton_pp(A[:, 0], A[:, 1], ra, B[:, 0], B[:, 1], rb, G, G)
        x1, y1, n1, ok1, w1 = _newton_pp(A[:, 0], A[:, 1], ra, B[:, 0], B[:, 1], rb, -G, -G)
        s0 = np.sign(ori(A, B, np.stack([x0, y0], axis=1)))
        pickx = np.where(s0 == csign, x0, x1)
        picky = np.where(s0 == csign, y0, y1)
        right = np.hypot(pickx - Q[:, 0], picky - Q[:, 1]) < 1e-6 * (1.0 + np.hypot(Q[:, 0], Q[:, 1]))
        hroot = np.abs(ori(A, B, Q))  # |det J| / 4 at the root
        safe = ok0 & ok1 & (n0 <= 40) & (n1 <= 40) & (np.minimum(w0, w1) > 1e-3 * hroot)
        bad[free[k][~(right & safe) | flat]] = True
        a0 = np.abs(ori(C[tri0[:, 0]], C[tri0[:, 1]], C[tri0[:, 2]]))[0]
        if a0 < 0.05 * np.hypot(*(C[p2] - C[p1])) * np.hypot(*(C[p3] - C[p1])) or                 np.sign(ori(C[tri0[:, 0]], C[tri0[:, 1]], C[tri0[:, 2]]))[0] != np.sign(ori(E[tri0[:, 0]], E[tri0[:, 1]], E[tri0[:, 2]]))[0]:
            bad[p3] = True
        idx = np.nonzero(bad)[0]
        if len(idx) == 0:
            break
        base = idx[idx < 3]
        E[base] = rng.uniform(-300, 300, size=(len(base), 2))
        draw(idx[idx >= 3])
        noise[idx] = rng.normal(0, 1.0, size=(len(idx), 2))
    else:
        raise RuntimeError(f"make_linkage: {len(idx)} leaves still fail the reference's own heuristic after {max_rounds} rounds")
    C = canvas()
    return [{"type": P, "canvas": [float(x), float(y)]} for x, y in C], edge_dicts()
