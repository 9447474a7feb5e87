"""Synthetic sub-system batches for the BASELINE.json configs.

Counter-based RNG: splitmix64 keyed by (seed, index, field); every instance is a pure function
of its index, so shards and the on-device generator (csrc/synth.cu, K1) reproduce the same
stream.  Only + - * / sqrt are used (no sin/cos), so host (numpy) and device agree bit for bit.

What is generated is what the host packer (host/, Appendix B of SURVEY.md) would emit for a
matched leaf: solver-space constants, the orientation code derived from a canvas layout through
the reference's own canvas-side formulas, canvas normals as guesses.
"""
from __future__ import annotations

import numpy as np

from . import capi

BASE_SEED = 0x5EED0001
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
FIELDS = 32


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, idx: np.ndarray, field: int) -> np.ndarray:
    """U[0,1) double for (seed, index, field): splitmix64 of seed + (index*32+field+1)*GOLD."""
    with np.errstate(over="ignore"):
        ctr = idx.astype(np.uint64) * np.uint64(FIELDS) + np.uint64(field + 1)
        z = _mix(np.uint64(seed) + ctr * _GOLD)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _rot(t):
    """Rational rotation from t = tan(theta/2): (cos, sin), no transcendental functions."""
    den = 1.0 + t * t
    return (1.0 - t * t) / den, (2.0 * t) / den


def _sgn3(x):
    return (x > 0).astype(np.int64) - (x < 0).astype(np.int64)


def _sign_of(x):
    """two-valued signOf of the reference (zero -> -1): point_line_solvers.cpp:195."""
    return np.where(x > 0.0, 1.0, -1.0)


def _tri_ori(ax, ay, bx, by, cx, cy):
    return ((bx - ax) * (cy - ay)) - ((by - ay) * (cx - ax))


def _signed_dist(px, py, l1x, l1y, l2x, l2y):
    """signedDistanceToLine, heuristics.hpp:113-125."""
    dx, dy = l2x - l1x, l2y - l1y
    ln = np.sqrt(dx * dx + dy * dy)
    cross = (dx * (py - l1y)) - (dy * (px - l1x))
    return cross / ln


def _frame(seed, idx, f0):
    """A rigid motion per instance: rotation (c, s) with |theta| < pi and a translation."""
    t = 2.0 * uniform(seed, idx, f0) - 1.0
    # widen: use t/(1-|t|*0.999) so the angle covers nearly (-pi, pi)
    t = t / (1.0 - np.abs(t) * 0.999)
    c, s = _rot(t)
    tx = 1000.0 * uniform(seed, idx, f0 + 1)
    ty = 1000.0 * uniform(seed, idx, f0 + 2)
    return c, s, tx, ty


def _apply(fr, x, y):
    c, s, tx, ty = fr
    return (c * x - s * y) + tx, (s * x + c * y) + ty


def make_pp(n, seed=BASE_SEED, first=0, n_seeds=2, perturb_of=0, scale=1.0, flat=None) -> capi.HostBatch:
    """K1 batch (config 2/3/5).  perturb_of > 0: instance i is base cluster (i % perturb_of)
    with ra, rb, d multiplied by 1 + 0.05*U[-1,1] (config 5, redrawn until the circles meet).
    `flat` (0..1) shrinks |Py| to make near-degenerate (flat) triangles for the slow QR paths."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    base = idx % np.uint64(perturb_of) if perturb_of else idx
    d = (10.0 + 490.0 * uniform(seed, base, 0)) * scale
    px = (-300.0 + 1100.0 * uniform(seed, base, 1)) * scale
    pym = (5.0 + 495.0 * uniform(seed, base, 2)) * scale
    if flat is not None:
        pym = pym * flat
    side = np.where(uniform(seed, base, 3) < 0.5, 1.0, -1.0)
    py = side * pym
    ra = np.sqrt(px * px + py * py)
    rb = np.sqrt((px - d) * (px - d) + py * py)
    if perturb_of:
        ok = np.zeros(n, dtype=bool)
        ra2, rb2, d2 = ra.copy(), rb.copy(), d.copy()
        for attempt in range(8):
            f = 4 + 3 * attempt
            pa = ra * (1.0 + 0.05 * (2.0 * uniform(seed ^ 0xABCDEF, idx, f) - 1.0))
            pb = rb * (1.0 + 0.05 * (2.0 * uniform(seed ^ 0xABCDEF, idx, f + 1) - 1.0))
            pd = d * (1.0 + 0.05 * (2.0 * uniform(seed ^ 0xABCDEF, idx, f + 2) - 1.0))
            good = (pa + pb > pd * 1.001) & (np.abs(pa - pb) < pd * 0.999)
            take = good & ~ok
            ra2[take], rb2[take], d2[take] = pa[take], pb[take], pd[take]
            ok |= good
        ra, rb, d = ra2, rb2, d2  # instances never accepted keep the unperturbed base values
    # canvas layout: the true triple moved rigidly
    cf = _frame(seed, base, 28)
    cax, cay = _apply(cf, np.zeros(n), np.zeros(n))
    cbx, cby = _apply(cf, d, np.zeros(n))
    cpx, cpy = _apply(cf, px, py)
    sign = _sgn3(_tri_ori(cax, cay, cbx, cby, cpx, cpy))
    # odd instances: general fixed positions (TwoFixedPointsDistance shape)
    sf = _frame(seed, base, 24)
    odd = (base % np.uint64(2)) == 1
    sax, say = _apply(sf, np.zeros(n), np.zeros(n))
    sbx, sby = _apply(sf, d, np.zeros(n))
    ax = np.where(odd, sax, 0.0)
    ay = np.where(odd, say, 0.0)
    bx = np.where(odd, sbx, d)
    by = np.where(odd, sby, 0.0)
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in (ax, ay, ra, bx, by, rb)]
    return capi.HostBatch(capi.KIND_PP, n_seeds, cols, capi.make_code(sign))


def make_ang(n, seed=BASE_SEED + 5, first=0) -> capi.HostBatch:
    """K5 batch: fixed line direction, angle constraint, canvas normal guess with noise, point
    distance (both LLP-anchor and FixedLineAndPoint shapes: the kernel sees the same columns)."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    lf = 50.0 + 450.0 * uniform(seed, idx, 0)
    c, s, _, _ = _frame(seed, idx, 1)
    even = (idx % np.uint64(2)) == 0
    # even: anchored shape (line 1 on the x axis, fd = (len, 0)); odd: general direction
    fdx = np.where(even, lf, lf * c)
    fdy = np.where(even, 0.0, lf * s)
    cos_a = -0.9961946980917455 + 2.0 * 0.9961946980917455 * uniform(seed, idx, 4)  # cos(5..175 deg)
    sin_a = np.sqrt(1.0 - cos_a * cos_a)
    ux, uy = fdx / lf, fdy / lf
    branch = np.where(uniform(seed, idx, 5) < 0.5, 1.0, -1.0)
    # true free direction = fixed direction rotated by +-angle; normal n with dir = (-ny, nx)
    dirx = ux * cos_a - branch * uy * sin_a
    diry = uy * cos_a + branch * ux * sin_a
    nx, ny = diry, -dirx
    # canvas frame + noise on the canvas free-line normal
    cc, cs, _, _ = _frame(seed, idx, 6)
    t = 0.15 * (2.0 * uniform(seed, idx, 9) - 1.0)
    nc, ns_ = _rot(t)
    gx0 = nx * nc - ny * ns_
    gy0 = nx * ns_ + ny * nc
    # rotate into the canvas frame only for the odd (general) shape; the guess is what the
    # packer reads from the canvas: unit normal of the canvas free line
    gnx = np.where(even, gx0, cc * gx0 - cs * gy0)
    gny = np.where(even, gy0, cs * gx0 + cc * gy0)
    nrm = np.sqrt(gnx * gnx + gny * gny)
    gnx, gny = gnx / nrm, gny / nrm
    cfdx = np.where(even, fdx, cc * fdx - cs * fdy)
    cfdy = np.where(even, fdy, cs * fdx + cc * fdy)
    flip = uniform(seed, idx, 10) < 0.5
    cfreex, cfreey = -gny, gnx  # canvas free direction from its normal
    cfreex = np.where(flip, -cfreex, cfreex)
    cfreey = np.where(flip, -cfreey, cfreey)
    sign = _sgn3(cfdx * cfreey - cfdy * cfreex)
    px = np.where(even, 0.0, -500.0 + 1000.0 * uniform(seed, idx, 11))
    py = (5.0 + 295.0 * uniform(seed, idx, 12)) * np.where(uniform(seed, idx, 13) < 0.5, 1.0, -1.0)
    sdist = (5.0 + 295.0 * uniform(seed, idx, 14)) * np.where(uniform(seed, idx, 15) < 0.5, 1.0, -1.0)
    r2x = np.where(even, 0.0, -500.0 + 1000.0 * uniform(seed, idx, 16))
    r2y = np.where(even, 0.0, -500.0 + 1000.0 * uniform(seed, idx, 17))
    clen = 50.0 + 450.0 * uniform(seed, idx, 18)
    cols = [fdx, fdy, cos_a, gnx, gny, cfdx, cfdy, px, py, sdist, r2x, r2y, clen]
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    return capi.HostBatch(capi.KIND_ANG, 2, cols, capi.make_code(sign))


def make_sdd(n, seed=BASE_SEED + 2, first=0) -> capi.HostBatch:
    """K2 batch: two fixed points and a free line at given distances."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    even = (idx % np.uint64(2)) == 0
    d12 = 20.0 + 480.0 * uniform(seed, idx, 0)
    fr = _frame(seed, idx, 1)
    p1x, p1y = _apply(fr, np.zeros(n), np.zeros(n))
    p2x, p2y = _apply(fr, d12, np.zeros(n))
    p1x, p1y = np.where(even, 0.0, p1x), np.where(even, 0.0, p1y)
    p2x, p2y = np.where(even, d12, p2x), np.where(even, 0.0, p2y)
    # true line: unit normal n (rational rotation), offset chosen so distances are feasible
    t = 2.0 * uniform(seed, idx, 4) - 1.0
    t = t / (1.0 - np.abs(t) * 0.99)
    nx, ny = _rot(t)
    off = -400.0 + 800.0 * uniform(seed, idx, 5)
    p = (nx * p1x + ny * p1y) - off
    sd1 = (nx * p1x + ny * p1y) - p
    sd2 = (nx * p2x + ny * p2y) - p
    # canvas line = true line +- small rotation; canvas sides give the signs
    tt = 0.1 * (2.0 * uniform(seed, idx, 6) - 1.0)
    # every 5th instance: a badly drawn canvas line (up to ~+-125 degrees off), so that guess 0
    # can fall into the other root's basin and candidate 1 gets selected
    tt = np.where((idx % np.uint64(5)) == 0, 19.0 * tt, tt)
    nc, ns_ = _rot(tt)
    gnx = nx * nc - ny * ns_
    gny = nx * ns_ + ny * nc
    sign1 = _sgn3(sd1)
    sign2 = _sgn3(sd2)
    s1 = _sign_of(sd1) * np.abs(sd1)
    s2 = _sign_of(sd2) * np.abs(sd2)
    clen = 50.0 + 450.0 * uniform(seed, idx, 7)
    cols = [p1x, p1y, p2x, p2y, s1, s2, gnx, gny, clen]
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    return capi.HostBatch(capi.KIND_SDD, 2, cols, capi.make_code(sign1, sign2))


def make_ppl(n, seed=BASE_SEED + 3, first=0, n_seeds=2, collinear_every=17) -> capi.HostBatch:
    """K3 batch: fixed point + fixed line, free point."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    fr = _frame(seed, idx, 0)
    # local frame: line along x through origin; fixed point at (fx, fy); free point at (qx, qy)
    llen = 50.0 + 450.0 * uniform(seed, idx, 3)
    fx = -200.0 + 400.0 * uniform(seed, idx, 4)
    fy = -200.0 + 400.0 * uniform(seed, idx, 5)
    qx = -300.0 + 600.0 * uniform(seed, idx, 6)
    qy = (5.0 + 295.0 * uniform(seed, idx, 7)) * np.where(uniform(seed, idx, 8) < 0.5, 1.0, -1.0)
    xa, ya = _apply(fr, -llen / 2.0, np.zeros(n))
    xb, yb = _apply(fr, llen / 2.0, np.zeros(n))
    px, py = _apply(fr, fx, fy)
    r = np.sqrt((qx - fx) * (qx - fx) + (qy - fy) * (qy - fy))
    # canvas layout = another rigid motion of the same local configuration
    cf = _frame(seed, idx, 9)
    cxa, cya = _apply(cf, -llen / 2.0, np.zeros(n))
    cxb, cyb = _apply(cf, llen / 2.0, np.zeros(n))
    cpx, cpy = _apply(cf, fx, fy)
    cqx, cqy = _apply(cf, qx, qy)
    s = _sign_of(_signed_dist(cqx, cqy, cxa, cya, cxb, cyb)) * np.abs(qy)
    # canvas perpendicular foot of the fixed point on the canvas line (heuristics.hpp:144-150)
    dx, dy = cxb - cxa, cyb - cya
    tt = (dx * (cpx - cxa) + dy * (cpy - cya)) / (dx * dx + dy * dy)
    cfx_, cfy_ = cxa + tt * dx, cya + tt * dy
    ori = _tri_ori(cpx, cpy, cfx_, cfy_, cqx, cqy)
    coll = np.abs(ori) < 1e-8
    if collinear_every:
        coll = coll | ((idx % np.uint64(collinear_every)) == 0)
    flags = np.where(coll, capi.CODE_COLLINEAR, 0)
    cols = [px, py, r, xa, ya, xb, yb, s, cqx, cqy]
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    return capi.HostBatch(capi.KIND_PPL, n_seeds, cols, capi.make_code(_sgn3(ori), 0, flags))


def make_pll(n, seed=BASE_SEED + 4, first=0, n_seeds=2, parallel_every=13) -> capi.HostBatch:
    """K4 batch: two fixed lines, free point (linear system; some parallel pairs)."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    fr = _frame(seed, idx, 0)
    l1 = 50.0 + 450.0 * uniform(seed, idx, 3)
    l2 = 50.0 + 450.0 * uniform(seed, idx, 4)
    t = 0.05 + 0.9 * uniform(seed, idx, 5)
    par = ((idx % np.uint64(parallel_every)) == 0) if parallel_every else np.zeros(n, dtype=bool)
    c2, s2 = _rot(t)
    c2 = np.where(par, 1.0, c2)
    s2 = np.where(par, 0.0, s2)
    ox = -100.0 + 200.0 * uniform(seed, idx, 6)
    oy = np.where(par, 40.0, -100.0 + 200.0 * uniform(seed, idx, 7))
    qx = -300.0 + 600.0 * uniform(seed, idx, 8)
    qy = -300.0 + 600.0 * uniform(seed, idx, 9)
    # local: line 1 along x through the origin; line 2 through (ox, oy) with direction (c2, s2)
    a1 = (-l1 / 2.0, np.zeros(n))
    b1 = (l1 / 2.0, np.zeros(n))
    a2 = (ox - (l2 / 2.0) * c2, oy - (l2 / 2.0) * s2)
    b2 = (ox + (l2 / 2.0) * c2, oy + (l2 / 2.0) * s2)
    d1 = np.abs(_signed_dist(qx, qy, a1[0], a1[1], b1[0], b1[1]))
    d2 = np.abs(_signed_dist(qx, qy, a2[0], a2[1], b2[0], b2[1]))
    xa1, ya1 = _apply(fr, *a1)
    xb1, yb1 = _apply(fr, *b1)
    xa2, ya2 = _apply(fr, *a2)
    xb2, yb2 = _apply(fr, *b2)
    cf = _frame(seed, idx, 10)
    ca1, cb1, ca2, cb2 = _apply(cf, *a1), _apply(cf, *b1), _apply(cf, *a2), _apply(cf, *b2)
    cqx, cqy = _apply(cf, qx, qy)
    s1 = _sign_of(_signed_dist(cqx, cqy, ca1[0], ca1[1], cb1[0], cb1[1])) * d1
    s2_ = _sign_of(_signed_dist(cqx, cqy, ca2[0], ca2[1], cb2[0], cb2[1])) * d2
    # canvas intersection (heuristics.hpp:165-181) and the reference triangle
    e1x, e1y = cb1[0] - ca1[0], cb1[1] - ca1[1]
    e2x, e2y = cb2[0] - ca2[0], cb2[1] - ca2[1]
    cross = e1x * e2y - e1y * e2x
    cpar = np.abs(cross) < 1e-10
    safe = np.where(cpar, 1.0, cross)
    dlx, dly = ca2[0] - ca1[0], ca2[1] - ca1[1]
    tt = (dlx * e2y - dly * e2x) / safe
    ix, iy = ca1[0] + tt * e1x, ca1[1] + tt * e1y
    z = e1x * e1x + e1y * e1y
    nz = np.sqrt(z)
    ux, uy = e1x / nz, e1y / nz
    ori = _tri_ori(ix, iy, ix + ux, iy + uy, cqx, cqy)
    coll = (np.abs(ori) < 1e-8) & ~cpar
    flags = np.where(cpar, capi.CODE_CANVAS_PARALLEL, 0) | np.where(coll, capi.CODE_COLLINEAR, 0)
    sign = np.where(cpar, 0, _sgn3(ori))
    cols = [xa1, ya1, xb1, yb1, s1, xa2, ya2, xb2, yb2, s2_, cqx, cqy]
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    return capi.HostBatch(capi.KIND_PLL, n_seeds, cols, capi.make_code(sign, 0, flags))


MAKERS = {capi.KIND_PP: make_pp, capi.KIND_SDD: make_sdd, capi.KIND_PPL: make_ppl,
          capi.KIND_PLL: make_pll, capi.KIND_ANG: make_ang}


def make(kind, n, **kw) -> capi.HostBatch:
    return MAKERS[kind](n, **kw)


# algorithmic work model (SURVEY.md section 8d / BASELINE.md section 5): flops per evaluation
F_EVAL = {1: 64, 2: 56, 3: 58, 4: 52, 5: 56}
F_SELECT = {1: 12, 2: 50, 3: 14, 4: 14, 5: 50}


def algorithmic_flops(kind, iters: np.ndarray) -> float:
    """W = sum_seeds (iters_s + 1) * F_kind + F_select per solve, from MEASURED iteration counts."""
    it = iters.astype(np.int64)
    n = it.shape[1]
    return float((it + 1).sum()) * F_EVAL[kind] + float(n) * F_SELECT[kind]
