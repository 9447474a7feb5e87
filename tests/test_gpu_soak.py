"""Bounded soak of the contracted kernels (GCS_VARIANT_CONTRACTED) where the driver sees it.

The discrete half of their contract (iteration counts, convergence flags, root indices equal to the
reference arithmetic) rests on guards whose sufficiency is an error-analysis argument, not a proof
(csrc/newton_relaxed.cuh), so the evidence has to be kept fresh: every build draws NEW generator
streams here (the stream offset is derived from the source hash compiled into the library), over
every kind, scales from 1e-7 to 1e6, flat triangles, rescaled K2..K5 systems, explicit seeds next to
the roots / inside the iteration-0 box / next to the line where the Jacobian is singular, and the
8-seed multi-start.  About 9e6 sub-systems, 15-30 s.  Bar: counts / flags / roots equal,
coordinates within 1e-9 relative, NaN / inf patterns equal (tests/util.py).
"""
import re

import numpy as np
import pytest

import oracle_lib as O
from util import assert_batches_within_contract

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def offset(gpu):
    """Generator-stream offset of this build: the source hash in gcs_b200_version()."""
    m = re.search(rb"src ([0-9a-f]+)", gpu.load().gcs_b200_version())
    return int(m.group(1)[:7], 16) if m else 0


def _check(gpu, make, what, variant=None):
    a = make()
    a.variant = gpu.VARIANT_CONTRACTED if variant is None else variant
    gpu.solve_host(a.alloc_outputs(), 0)
    b = O.solve(make().alloc_outputs(), threads=0)
    worst = assert_batches_within_contract(a, b, what)
    assert worst <= 1e-9
    return a, b


@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("stream", [0, 1])
def test_fresh_generator_streams(gpu, gcs, offset, kind, stream):
    seed = 0x1234 + 7919 * stream + offset
    _check(gpu, lambda: gcs.synth.make(kind, 1 << 18, seed=seed), f"K{kind} seed {seed:#x}")


@pytest.mark.parametrize("scale,flat", [(1e-7, None), (1e-6, None), (3e-5, None), (1e-3, 1e-2), (1.0, 1e-7), (30.0, 1e-3),
                                         (1e3, 1e-4), (1e6, None)])
def test_k1_scales_and_flat_triangles(gpu, gcs, offset, scale, flat):
    """The absolute 1e-5 threshold above, inside and below the size of the system; triangles flattened
    until the conditioning guard, careful mode and the literal re-run all take part."""
    _check(gpu, lambda: gcs.synth.make_pp(1 << 18, seed=4242 + offset, scale=scale, flat=flat), f"K1 scale {scale} flat {flat}")


LENGTH_COLS = {2: [0, 1, 2, 3, 4, 5, 8], 3: list(range(10)), 4: list(range(12)), 5: [0, 1, 7, 8, 9, 10, 11, 12]}


@pytest.mark.parametrize("kind", [2, 3, 4, 5])
@pytest.mark.parametrize("scale", [1e-7, 3e-5, 1e-3, 1e6])
def test_rescaled_systems_of_the_other_kinds(gpu, gcs, offset, kind, scale):
    def make():
        hb = gcs.synth.make(kind, 1 << 17, seed=0x5CA1E + kind + offset)
        for c in LENGTH_COLS[kind]:
            hb.cols[c] = np.ascontiguousarray(hb.cols[c] * scale)
        return hb
    _check(gpu, make, f"K{kind} lengths x {scale:g}")


def _guesses(rng, kind, mode, base, n):
    unit = kind in (2, 5)  # unknowns are unit normals
    size = 1.0 if unit else 1000.0
    if mode == "box":
        return rng.uniform(-3 * size, 3 * size, size=(2, 2, n))
    if mode == "wide":
        return rng.uniform(-1.0, 1.0, size=(2, 2, n)) * np.exp(rng.uniform(np.log(1e-4), np.log(1e6), size=(2, 1, n))) * size
    if mode == "near_root":
        g = np.array(base.cand, copy=True)
        return g * (1.0 + rng.normal(0, 1.0, size=g.shape) * np.exp(rng.uniform(np.log(1e-12), np.log(0.3), size=(2, 1, n))))
    if mode == "tiny":  # straddles the iteration-0 box |guess| < 1e-5
        return rng.uniform(-3e-5, 3e-5, size=(2, 2, n))
    assert mode == "singular"
    if kind == 1:  # a hair off the line through the two fixed points, where the Jacobian is singular
        ax, ay, _, bx, by, _ = base.cols
        t = rng.uniform(-0.5, 1.5, size=(2, n))
        off = np.exp(rng.uniform(np.log(1e-12), np.log(1e-1), size=(2, n))) * rng.choice([-1.0, 1.0], size=(2, n))
        g = np.empty((2, 2, n))
        g[:, 0, :] = ax + t * (bx - ax) - off * (by - ay)
        g[:, 1, :] = ay + t * (by - ay) + off * (bx - ax)
        return g
    mid = 0.5 * (base.cand[0] + base.cand[1])  # the mid point of the two roots lies on the singular line
    return mid[None] * (1.0 + rng.normal(0, 1, size=(2, 2, n)) * np.exp(rng.uniform(np.log(1e-12), np.log(1e-2), size=(2, 1, n))))


@pytest.mark.parametrize("kind", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("mode", ["box", "wide", "near_root", "tiny", "singular"])
def test_explicit_seeds(gpu, gcs, offset, kind, mode):
    n = 1 << 16
    rng = np.random.default_rng(2026 + offset + 31 * kind)
    base = gcs.synth.make(kind, n, seed=0xC0DE + kind + offset)
    base.want_cand = True
    O.solve(base.alloc_outputs(), threads=0)
    g = np.ascontiguousarray(_guesses(rng, kind, mode, base, n))

    def make():
        hb = gcs.synth.make(kind, n, seed=0xC0DE + kind + offset)
        hb.guesses = g
        return hb
    _check(gpu, make, f"K{kind} explicit seeds: {mode}")


@pytest.mark.parametrize("kind", [1, 3])
def test_eight_seeds(gpu, gcs, offset, kind):
    _check(gpu, lambda: gcs.synth.make(kind, 1 << 16, seed=31337 + kind + offset, n_seeds=8), f"K{kind} x 8 seeds")


@pytest.mark.parametrize("kind", [1, 5])
def test_sorted_mapping_of_the_same_arithmetic(gpu, gcs, offset, kind):
    _check(gpu, lambda: gcs.synth.make(kind, 1 << 18, seed=0xABBA + offset), f"K{kind} contracted-sorted",
           variant=gpu.VARIANT_CONTRACTED_SORTED)


@pytest.mark.parametrize("lo,hi", [(-6, -2), (-9, -5), (-3, 0)])
def test_needle_triangles(gpu, gcs, offset, lo, hi):
    """K1 with the free point 10^lo .. 10^hi from one of the centres: well conditioned, so the runs stay on
    the closed-form path, and the line form's half chord has to come from cancellation-free factors
    (tests/test_gpu_relaxed.py::test_contract_roots_next_to_a_centre is the fixed-seed regression test)."""
    n = 1 << 18
    rng = np.random.default_rng(7001 + offset + hi)
    rho = 10.0 ** rng.uniform(lo, hi, size=n)
    phi = rng.uniform(0.02, np.pi - 0.02, size=n) * np.where(rng.random(n) < 0.5, 1.0, -1.0)
    near_b = rng.random(n) < 0.5

    def make():
        hb = gcs.synth.make(1, n, seed=0xBEE5 + offset)
        ax, ay, ra, bx, by, rb = hb.cols
        cx, cy = np.where(near_b, bx, ax), np.where(near_b, by, ay)
        px, py = cx + rho * np.cos(phi), cy + rho * np.sin(phi)
        ra[:] = np.hypot(px - ax, py - ay)
        rb[:] = np.hypot(px - bx, py - by)
        return hb
    _check(gpu, make, f"K1 needles 1e{lo} .. 1e{hi}")


@pytest.mark.parametrize("n_seeds", [2, 8])
def test_linear_pairs(gpu, gcs, offset, n_seeds):
    """K4 through newton_linear_kernel on fresh streams, lines of mismatched lengths (1 .. 1e-3)."""
    n = 1 << 18
    rng = np.random.default_rng(515 + offset + n_seeds)
    ratio = 10.0 ** rng.uniform(-3, 0, size=n)

    def make():
        hb = gcs.synth.make(4, n, seed=0xF00D + offset, n_seeds=n_seeds)
        c = hb.cols
        mx, my = 0.5 * (c[5] + c[7]), 0.5 * (c[6] + c[8])
        c[5][:], c[7][:] = mx + (c[5] - mx) * ratio, mx + (c[7] - mx) * ratio
        c[6][:], c[8][:] = my + (c[6] - my) * ratio, my + (c[8] - my) * ratio
        return hb
    _check(gpu, make, f"K4 x {n_seeds} seeds, mismatched lengths")
