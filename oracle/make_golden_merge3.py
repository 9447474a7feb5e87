#!/usr/bin/env python
"""Golden vectors for the bottom-up Merge3 numeric helpers, generated
from the reference's own merge3_solver_common.cpp (oracle/_ref/libgcs_ref.so, built by
oracle/build_ref.sh from /root/reference).  Run in the build container:

    python oracle/make_golden_merge3.py        ->  tests/golden/merge3.npz

Per case: seeded random configurations built from a TRUE solution (so the distances are
consistent), a canvas layout that is a rigid motion (sometimes mirrored) of the true layout plus
noise, and hand-made degenerate rows (zero-length canvas lines, collinear canvas triples, parallel
lines, degenerate line A).  Stored: the argument rows and what the reference returns.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_lib as R  # noqa: E402

N = 192


def motion(rng, mirror):
    th = rng.uniform(0, 2 * math.pi)
    c, s = math.cos(th), math.sin(th)
    t = rng.uniform(0, 1000, size=2)
    m = -1.0 if mirror else 1.0

    def f(p):
        x, y = p[0], m * p[1]
        return np.array([c * x - s * y + t[0], s * x + c * y + t[1]])
    return f


def sd(p, a, b):
    d = b - a
    return (d[0] * (p[1] - a[1]) - d[1] * (p[0] - a[0])) / math.hypot(d[0], d[1])


def rows_pp(rng):
    out = []
    for i in range(N):
        a, b, p = rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2)
        f = motion(rng, mirror=(i % 3 == 0))
        ca, cb, cp = f(a), f(b), f(p) + rng.normal(0, 3, 2)
        if i % 37 == 5:
            cp = ca + 0.25 * (cb - ca)  # collinear canvas triple: orientation sign 0
        out.append(np.concatenate([a, b, [np.linalg.norm(p - a), np.linalg.norm(p - b)], ca, cb, cp]))
    return np.array(out)


def rows_line(rng):
    out = []
    for i in range(N):
        a, b = rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2)
        th = rng.uniform(0, 2 * math.pi)
        q = rng.uniform(-300, 300, 2)          # a point of the true free line
        d = np.array([math.cos(th), math.sin(th)])
        l1, l2 = q - 80 * d, q + 120 * d
        f = motion(rng, mirror=(i % 4 == 1))
        ca, cb = f(a), f(b)
        c1, c2 = f(l1) + rng.normal(0, 2, 2), f(l2) + rng.normal(0, 2, 2)
        if i % 41 == 7:
            c2 = c1.copy()                      # zero-length canvas line: direction (1,0), length 50
        out.append(np.concatenate([a, b, [abs(sd(a, l1, l2)), abs(sd(b, l1, l2))], ca, cb, c1, c2]))
    return np.array(out)


def rows_pl(rng):
    out = []
    for i in range(N):
        fp = rng.uniform(-500, 500, 2)
        l1, l2 = rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2)
        p = rng.uniform(-500, 500, 2)
        f = motion(rng, mirror=(i % 3 == 2))
        cfp, c1, c2, cp = f(fp), f(l1), f(l2), f(p) + rng.normal(0, 3, 2)
        if i % 29 == 3:
            # canvas free point on the line through the canvas fixed point and its foot: collinear -> nearest
            d = c2 - c1
            t = np.dot(d, cfp - c1) / np.dot(d, d)
            foot = c1 + t * d
            cp = cfp + 0.4 * (foot - cfp)
        out.append(np.concatenate([fp, l1, l2, [np.linalg.norm(p - fp), abs(sd(p, l1, l2))], cfp, c1, c2, cp]))
    return np.array(out)


def rows_ll(rng):
    out = []
    for i in range(N):
        a1, a2 = rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2)
        b1, b2 = rng.uniform(-500, 500, 2), rng.uniform(-500, 500, 2)
        if i % 23 == 4:
            b2 = b1 + 1.5 * (a2 - a1)           # parallel solver lines: nearest-to-canvas
        if i % 47 == 11:
            a2 = a1 + np.array([4e-10, 3e-10])  # solver line A shorter than EPSILON, intersections exist: nullopt
        p = rng.uniform(-500, 500, 2)
        f = motion(rng, mirror=(i % 5 == 0))
        ca1, ca2, cb1, cb2, cp = f(a1), f(a2), f(b1), f(b2), f(p) + rng.normal(0, 3, 2)
        if i % 31 == 6:
            cb1, cb2 = ca1 + np.array([7.0, 3.0]), ca2 + np.array([7.0, 3.0])  # parallel canvas lines
        if i % 43 == 9:
            ca2 = ca1 + np.array([4e-10, 3e-10])  # canvas line A shorter than EPSILON, intersections exist: nullopt
        out.append(np.concatenate([a1, a2, b1, b2, [abs(sd(p, a1, a2)), abs(sd(p, b1, b2))], ca1, ca2, cb1, cb2, cp]))
    return np.array(out)


def rigid_cases(rng):
    src, dst, npts = [], [], []
    for i in range(96):
        n = [1, 2, 2, 3, 4, 6][i % 6]
        s = rng.uniform(-400, 400, (n, 2))
        f = motion(rng, mirror=(i % 7 == 3))    # mirrored targets exercise the determinant fix
        d = np.array([f(p) for p in s]) + rng.normal(0, 0.5, (n, 2)) * (i % 2)
        if i % 19 == 8 and n >= 2:
            s[1] = s[0]                         # coincident sources: rank-1 covariance
        pad = np.zeros((6, 2))
        ps, pd = pad.copy(), pad.copy()
        ps[:n], pd[:n] = s, d
        src.append(ps), dst.append(pd), npts.append(n)
    return np.array(src), np.array(dst), np.array(npts, dtype=np.int32)


def score_cases(rng):
    types, canvas, pose, inp = [], [], [], []
    for i in range(48):
        n = 8
        t = (rng.uniform(size=n) < 0.4).astype(np.int32)
        c = rng.uniform(0, 1000, (n, 4))
        p = c + rng.normal(0, 20, (n, 4))
        m = (rng.uniform(size=n) < 0.8).astype(np.uint8)
        if i == 5:
            m[:] = 0                            # empty pose: +inf
        if i % 11 == 2:
            p[t == 1, 2:] = p[t == 1, :2]       # zero-length solved lines: no direction term
        types.append(t), canvas.append(c), pose.append(p), inp.append(m)
    return np.array(types), np.array(canvas), np.array(pose), np.array(inp)


def main():
    rng = np.random.default_rng(0x3E26E3)
    data = {}
    for kase, gen in ((1, rows_pp), (2, rows_line), (3, rows_pl), (4, rows_ll)):
        rows = gen(rng)
        out, ok = R.m3_solve(kase, rows)
        data[f"rows{kase}"], data[f"out{kase}"], data[f"ok{kase}"] = rows, out, ok
        print("case", kase, rows.shape, "nullopt", int((ok == 0).sum()), "nan", int(np.isnan(out[ok == 1]).sum()))
    src, dst, npts = rigid_cases(rng)
    tr = np.zeros((len(npts), 6))
    rc = np.zeros(len(npts), dtype=np.int32)
    for i in range(len(npts)):
        rc[i], tr[i] = R.m3_rigid_transform(src[i, :npts[i]], dst[i, :npts[i]])
    data.update(rigid_src=src, rigid_dst=dst, rigid_n=npts, rigid_rc=rc, rigid_out=tr)
    t, c, p, m = score_cases(rng)
    data.update(score_types=t, score_canvas=c, score_pose=p, score_in=m,
                score=np.array([R.m3_score(t[i], c[i], p[i], m[i]) for i in range(len(t))]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "merge3.npz"), **data)
    print("rigid ok", int(rc.sum()), "of", len(rc), "; scores", data["score"][:4])


if __name__ == "__main__":
    main()
