"""GPU parity of the tolerance-class variants (GCS_VARIANT_CONTRACTED*, csrc/newton_relaxed.cuh)
against the CPU oracle, through the C ABI.

Bar (BASELINE.json north_star, spelled out in util.assert_batches_within_contract): iteration
counts, convergence flags and the chosen root IDENTICAL; coordinates within 1e-9 relative.  The
cases are those of test_gpu_parity.py, including the ones built to sit on decision boundaries
(flat triangles, extreme scales, NaN / inf / degenerate inputs, never-converging rows, explicit
guesses inside the iteration-0 box), because those are where a different arithmetic could change a
discrete output and the guards must hand the run to the literal code."""
import numpy as np
import pytest

import oracle_lib as O
from util import assert_batches_identical, assert_batches_within_contract, bits

pytestmark = pytest.mark.gpu

KINDS = [1, 2, 3, 4, 5]
VARIANTS = [6, 7, 8]  # contracted static, contracted sorted, contracted sequential (one lane per sub-system)


def _solve_pair(gpu, synth, kind, n, variant, **kw):
    hb = synth.make(kind, n, **kw)
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = O.solve(synth.make(kind, n, **kw).alloc_outputs())
    return hb, ref


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [1, 2, 31, 64, 65, 127, 4099])
def test_contract_small_and_ragged(gpu, gcs, kind, n, variant):
    hb, ref = _solve_pair(gpu, gcs.synth, kind, n, variant)
    assert_batches_within_contract(hb, ref, f"kind {kind} n {n} variant {variant}")


@pytest.mark.parametrize("variant", VARIANTS + [5])
@pytest.mark.parametrize("kind", KINDS)
def test_contract_512k(gpu, gcs, kind, variant):
    hb, ref = _solve_pair(gpu, gcs.synth, kind, 1 << 19, variant)
    worst = assert_batches_within_contract(hb, ref, f"kind {kind} variant {variant}")
    assert worst <= 1e-9


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kind", [1, 3, 4])
def test_contract_multistart_8_seeds(gpu, gcs, kind, variant):
    hb, ref = _solve_pair(gpu, gcs.synth, kind, 20001, variant, n_seeds=8)
    assert_batches_within_contract(hb, ref, f"8 seeds kind {kind}")


@pytest.mark.parametrize("variant", VARIANTS)
def test_contract_explicit_guesses_and_early_exit(gpu, gcs, variant):
    synth = gcs.synth
    n = 5000
    rng = np.random.default_rng(7)
    g = rng.uniform(-3000, 3000, size=(2, 2, n))
    g[:, :, ::7] = rng.uniform(-9e-6, 9e-6, size=g[:, :, ::7].shape)  # |guess| < 1e-5: exits at i = 0
    g = np.ascontiguousarray(g)
    hb = synth.make_pp(n)
    hb.guesses, hb.variant = g, variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = synth.make_pp(n)
    ref.guesses = g
    O.solve(ref.alloc_outputs())
    assert_batches_within_contract(hb, ref, "explicit guesses")
    assert (hb.iters[:, ::7] == 0).all() and (hb.converged[:, ::7] == 1).all()
    assert np.array_equal(bits(hb.cand[:, :, ::7]), bits(g[:, :, ::7]))


@pytest.mark.parametrize("variant", VARIANTS)
def test_contract_guesses_next_to_the_singular_line(gpu, gcs, variant):
    """Seeds a hair off the line through the two centres, where the distance-distance Jacobian is
    singular: the first update throws the iterate far out (the one event that amplifies a
    deviation between two arithmetics, guard G2) - iteration counts must still be the oracle's."""
    synth = gcs.synth
    n = 20000
    rng = np.random.default_rng(11)
    hb = synth.make_pp(n)
    ax, ay, _, bx, by, _ = hb.cols
    t = rng.uniform(-0.5, 1.5, size=(2, n))
    off = np.exp(rng.uniform(np.log(1e-9), np.log(1e-1), size=(2, n))) * rng.choice([-1.0, 1.0], size=(2, n))
    ux, uy = bx - ax, by - ay
    g = np.empty((2, 2, n))
    g[:, 0, :] = ax + t * ux - off * uy
    g[:, 1, :] = ay + t * uy + off * ux
    g = np.ascontiguousarray(g)
    hb.guesses, hb.variant = g, variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = synth.make_pp(n)
    ref.guesses = g
    O.solve(ref.alloc_outputs())
    assert_batches_within_contract(hb, ref, "guesses next to the singular line")
    assert ref.converged.mean() > 0.9  # the case is about runs that do converge, after a detour


@pytest.mark.parametrize("variant", VARIANTS)
def test_contract_nan_inf_and_degenerate_inputs(gpu, gcs, variant):
    """Inputs on which no arithmetic can be trusted: every such run must come out of the literal
    code, i.e. bit-identical to the oracle."""
    synth = gcs.synth
    n = 256
    hb = synth.make_pp(n)
    ref = synth.make_pp(n)
    for b in (hb, ref):
        c = b.cols
        c[2][0::16] = np.nan
        c[0][1::16] = np.inf
        c[3][2::16] = c[0][2::16]
        c[4][2::16] = c[1][2::16]          # B == A: singular Jacobian for ever
        c[2][3::16] = 0.0
        c[5][3::16] = 0.0                  # zero radii
        c[2][4::16] = 1e-3
        c[5][4::16] = 1e-3                 # circles that do not meet
        c[0][5::16] = 1e300                # overflow in the residual
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    O.solve(ref.alloc_outputs())
    assert_batches_within_contract(hb, ref, "degenerate")
    bad = np.zeros(n, bool)
    for o in range(6):
        bad[o::16] = True
    a, b = hb.cand[..., bad], ref.cand[..., bad]
    assert ((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))).all()
    assert (hb.iters[:, 0::16] == 1000).all() and (hb.converged[:, 0::16] == 0).all()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("scale,flat", [(1e-6, None), (1e6, None), (1.0, 1e-7), (1.0, 1e-12), (1e3, 1e-4), (1e-150, None), (1e140, None)])
def test_contract_flat_triangles_and_extreme_scales(gpu, gcs, variant, scale, flat):
    """Flat triangles (ill-conditioned at the root: guard G1), tiny and huge scales (the band of
    guard G3 against an absolute threshold of 1e-5)."""
    n = 2000 if scale > 1e100 else 20000
    hb, ref = _solve_pair(gpu, gcs.synth, 1, n, variant, scale=scale, flat=flat)
    assert_batches_within_contract(hb, ref, f"scale {scale} flat {flat}")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("scale", [1e-6, 1e-7, 3e-5])
def test_contract_systems_smaller_than_the_tolerance_at_4m(gpu, gcs, variant, scale):
    """Systems whose roots lie closer together than the absolute 1e-5 threshold meet `< tol` while
    the iteration is still halving its way in from the far seed, carrying the relative deviation of
    the ill-conditioned first update (cond ~ 20000 / d, up to 2e9 here).  A soak of 1.1e8
    sub-systems found one run in 4.2e6 deciding differently before that deviation entered the
    guard band (RelaxGuard::add_carry); 2^22 sub-systems per case make such a run likely."""
    n = 1 << 22
    hb, ref = _solve_pair(gpu, gcs.synth, 1, n, variant, seed=4242, scale=scale)
    assert_batches_within_contract(hb, ref, f"scale {scale} variant {variant}")


def test_contract_device_resident_batch_and_default_is_still_bit_identical(gpu, gcs):
    import torch
    synth, capi = gcs.synth, gcs.capi
    hb = synth.make_ang(70001)
    db = capi.DeviceBatch(hb, "cuda:0", want_cand=True)
    ref = O.solve(synth.make_ang(70001).alloc_outputs())
    for variant in (5, 6, 7, 8):
        db.set_variant(variant)
        db.solve()
        torch.cuda.synchronize()
        assert_batches_within_contract(db.to_host(synth.make_ang(70001)), ref, f"device batch variant {variant}")
    db.set_variant(0)
    db.solve()
    torch.cuda.synchronize()
    assert_batches_identical(db.to_host(synth.make_ang(70001)), ref, "default variant")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("kind", [1, 3, 4])
def test_contract_seeds_so_far_away_that_the_landing_error_straddles_the_threshold(gpu, gcs, kind, variant):
    """A linear pair (K4) lands on its solution with its first update, from any seed; from a seed
    1e8 .. 1e11 away it lands with an error of eps * cond * |seed| ~ 1e-8 .. 1e-3, and the SECOND update
    is that error - rounding noise of the first update, different in any other arithmetic, sitting
    around the threshold 1e-5.  The margin of that decision must carry the first update's length
    (newton_relaxed.cuh: w1); found by tests/test_gpu_soak.py on a fresh stream (one run in 65536 with
    a different iteration count before w1 existed).  K1 / K3 from the same seeds halve their way in and
    never decide anything there: they ride along as controls."""
    n = 1 << 17
    rng = np.random.default_rng(4040 + kind)
    g = rng.uniform(-1.0, 1.0, size=(2, 2, n)) * 10.0 ** rng.uniform(7, 11, size=(2, 1, n))
    hb = gcs.synth.make(kind, n, seed=0xFA5 + kind)
    hb.guesses, hb.variant = np.ascontiguousarray(g), variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = gcs.synth.make(kind, n, seed=0xFA5 + kind)
    ref.guesses = hb.guesses
    O.solve(ref.alloc_outputs())
    assert_batches_within_contract(hb, ref, f"kind {kind} variant {variant} far seeds")
    if kind == 4:
        assert len(np.unique(ref.iters)) >= 2  # the landing error really decides: some runs need a third update


@pytest.mark.parametrize("variant", VARIANTS + [5])
def test_contract_roots_next_to_a_centre(gpu, gcs, variant):
    """Needle triangles: the free point lies 1e-6 .. 1e-2 from one of the two fixed points (one radius
    is orders of magnitude under the other and under the base).  The system is WELL conditioned there
    (the circles cross at the angle phi), so the runs stay on the closed-form path - and a line form
    that takes the half chord from `ra^2 - t0^2` cancels to eps ra^2 / (2 h): coordinates 1e-9 .. 1e-6
    off and late updates moved by more than the guards' band (scratch/soak_relaxed.py found it on
    the flat-triangle case: 1.1e-9 on the 0.1 % of rows with the apex next to A or B).  RLine<K1>
    builds h^2 from Heron's factors instead; this is the regression test."""
    n = 1 << 18
    rng = np.random.default_rng(9091)
    hb = gcs.synth.make(1, n, seed=0xC0FFEE)
    ax, ay, ra, bx, by, rb = hb.cols
    rho = 10.0 ** rng.uniform(-6, -2, size=n)
    phi = rng.uniform(0.02, np.pi - 0.02, size=n) * np.where(rng.random(n) < 0.5, 1.0, -1.0)
    near_b = rng.random(n) < 0.5
    cx, cy = np.where(near_b, bx, ax), np.where(near_b, by, ay)
    px, py = cx + rho * np.cos(phi), cy + rho * np.sin(phi)
    ra[:] = np.hypot(px - ax, py - ay)
    rb[:] = np.hypot(px - bx, py - by)
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = gcs.capi.HostBatch(1, 2, [c.copy() for c in hb.cols], hb.code.copy())
    O.solve(ref.alloc_outputs())
    worst = assert_batches_within_contract(hb, ref, f"needle triangles, variant {variant}")
    assert worst <= 1e-10  # the closed form is far inside the contract here, as the Cramer form was
    assert (ref.converged == 1).mean() > 0.99


# ---- K4 in the contracted class: the linear kernel (GCS_VARIANT_CONTRACTED_LINEAR, newton_linear_kernel) ----
LINEAR = 10


def _k4_pair(gpu, gcs, n, mutate=None, guesses=None, **kw):
    def make():
        hb = gcs.synth.make(4, n, **kw)
        if mutate is not None:
            mutate(hb.cols, hb)
        hb.guesses = guesses
        return hb
    hb = make()
    hb.variant = LINEAR
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = O.solve(make().alloc_outputs())
    return hb, ref


@pytest.mark.parametrize("n_seeds", [2, 8])
@pytest.mark.parametrize("n", [1, 2, 31, 64, 65, 127, 4099, 1 << 19])
def test_linear_kernel_sizes_and_seed_counts(gpu, gcs, n, n_seeds):
    """Includes the generator's parallel pairs (one row in 13: never converge, nearest-to-canvas) and
    collinear codes: not certified, so they come out of the literal code bit for bit."""
    import ctypes as C
    lib, st = gcs.capi.load(), (C.c_uint64 * 8)()
    lib.gcs_b200_contracted_stats_ex(0, st, 1)
    hb, ref = _k4_pair(gpu, gcs, n, n_seeds=n_seeds)
    lib.gcs_b200_contracted_stats_ex(0, st, 1)
    worst = assert_batches_within_contract(hb, ref, f"K4 linear n {n} seeds {n_seeds}")
    assert worst <= 1e-10
    if n >= 4099:  # parallel pairs (one row in 13) and collinear codes cannot be certified: nearly all of the rest must be
        hard = ((ref.iters != 2).any(axis=0) | ((hb.code & 0x30) != 0)).sum()
        assert sum(st) <= (hard + 0.02 * n) * n_seeds, f"literal runs {list(st)} of {n * n_seeds}, {hard} hard rows"


def test_linear_kernel_resolution(gpu, gcs):
    R = gcs.capi.load().gcs_b200_resolve_variant
    assert R(gcs.capi.VARIANT_CONTRACTED, 4, 17, 2) == LINEAR and R(LINEAR, 4, 1 << 20, 8) == LINEAR
    hb = gcs.synth.make(1, 3000)  # any other kind: as VARIANT_CONTRACTED
    hb.variant = LINEAR
    gpu.solve_host(hb.alloc_outputs(), 0)
    assert_batches_within_contract(hb, O.solve(gcs.synth.make(1, 3000).alloc_outputs()), "K1 through the linear variant id")


@pytest.mark.parametrize("mode", ["box", "far", "near_root", "tiny", "mixed"])
def test_linear_kernel_explicit_seeds(gpu, gcs, mode):
    """Seeds the two certified decisions depend on: inside / astride the iteration-0 box, within a few
    tolerances of the solution (the first update is then NOT longer than the threshold for certain),
    and 1e7 .. 1e11 away (the second update - the landing's rounding error - straddles the threshold:
    the case that found the w1 hole of the generic guards)."""
    n = 1 << 16
    rng = np.random.default_rng(808 + len(mode))
    base = O.solve(gcs.synth.make(4, n, seed=0xB0B).alloc_outputs())
    sol = np.stack([base.out[0], base.out[1]])  # [2, n]
    if mode == "box":
        g = rng.uniform(-3000, 3000, size=(2, 2, n))
    elif mode == "far":
        g = rng.uniform(-1.0, 1.0, size=(2, 2, n)) * 10.0 ** rng.uniform(7, 11, size=(2, 1, n))
    elif mode == "near_root":
        g = sol[None] + rng.normal(0, 1, size=(2, 2, n)) * 10.0 ** rng.uniform(-9, -3, size=(2, 1, n))
    elif mode == "tiny":
        g = rng.uniform(-3e-5, 3e-5, size=(2, 2, n))
    else:
        g = rng.uniform(-3000, 3000, size=(2, 2, n))
        g[0, :, ::3] = rng.uniform(-9e-6, 9e-6, size=g[0, :, ::3].shape)     # seed 0 exits at i = 0, seed 1 does not
        g[1, :, 1::3] = (sol + rng.normal(0, 3e-6, size=sol.shape))[:, 1::3]  # seed 1 starts within the tolerance of the solution
    g = np.ascontiguousarray(np.where(np.isfinite(g), g, 1.0))
    hb, ref = _k4_pair(gpu, gcs, n, guesses=g, seed=0xB0B)
    assert_batches_within_contract(hb, ref, f"K4 linear, explicit seeds: {mode}")
    if mode == "far":
        assert len(np.unique(ref.iters)) >= 2  # some runs do need a third update: those must not have been certified


def test_linear_kernel_degenerate_inputs(gpu, gcs):
    """What no closed form may vouch for comes out of the literal code: bit-identical candidates."""
    n = 4096

    def mutate(c, hb):
        c[0][0::16] = np.nan
        c[5][1::16] = np.inf
        c[2][2::16], c[3][2::16] = c[0][2::16], c[1][2::16]              # line 1 is a point
        c[7][3::16] = c[5][3::16] + (c[2][3::16] - c[0][3::16])
        c[8][3::16] = c[6][3::16] + (c[3][3::16] - c[1][3::16])          # parallel lines
        c[5][4::16], c[6][4::16], c[7][4::16], c[8][4::16] = c[0][4::16], c[1][4::16], c[2][4::16], c[3][4::16]  # the same line twice
        c[4][5::16] = 0.0                                                # s1 = 0: the orientation the selection tests is rounding noise
        c[4][6::16] = 1e-13                                              # ... or next to it
        c[0][7::16] = 1e300
        c[7][8::16] = c[5][8::16] + (c[2][8::16] - c[0][8::16]) * (1 + 1e-9)
        c[8][8::16] = c[6][8::16] + (c[3][8::16] - c[1][8::16]) * (1 - 1e-9)   # a hair off parallel: ill conditioned
        for col in (0, 1, 2, 3, 5, 6, 7, 8):
            c[col][9::16] *= 1e-7                                        # lines shorter than the tolerance, |cross| under the absolute epsilon
        hb.code[10::16] |= gcs.capi.CODE_COLLINEAR
        hb.code[11::16] |= gcs.capi.CODE_CANVAS_PARALLEL
    hb, ref = _k4_pair(gpu, gcs, n, mutate=mutate, parallel_every=0)
    assert_batches_within_contract(hb, ref, "K4 linear, degenerate inputs")
    for o in (0, 1, 2, 3, 4, 5, 7, 10, 11):
        a, b = hb.cand[..., o::16], ref.cand[..., o::16]
        assert ((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))).all(), f"row class {o} was certified"


@pytest.mark.parametrize("scale", [1e-7, 1e-5, 1e-3, 1e3, 1e6])
@pytest.mark.parametrize("ratio", [1.0, 1e-4])
def test_linear_kernel_scales_and_mismatched_lengths(gpu, gcs, scale, ratio):
    """Every length scaled (the thresholds are absolute), and line 2 made `ratio` times as long as
    line 1 (|J|_F^2 / |det| grows with the mismatch: the bound on the second update must see it)."""
    n = 1 << 17

    def mutate(c, hb):
        for col in range(12):
            c[col] *= scale
        mx, my = 0.5 * (c[5] + c[7]), 0.5 * (c[6] + c[8])
        c[5][:], c[7][:] = mx + (c[5] - mx) * ratio, mx + (c[7] - mx) * ratio
        c[6][:], c[8][:] = my + (c[6] - my) * ratio, my + (c[8] - my) * ratio
    hb, ref = _k4_pair(gpu, gcs, n, mutate=mutate, parallel_every=0, seed=0xD1CE)
    assert_batches_within_contract(hb, ref, f"K4 linear scale {scale} ratio {ratio}")


# ---- degenerate inputs of the kinds whose runs take the line form through their own constants (K2, K3, K5) ----
def _degenerate_rows(kind, c):
    """Mutates the columns `c` of a generator batch, row class o = index % 16; returns the classes whose
    candidates must come out of the literal code bit for bit (nothing a closed form may vouch for)."""
    if kind == 3:  # circle (px, py, r) and the line at signed distance s from (xa, ya) -> (xb, yb)
        ex, ey = c[5] - c[3], c[6] - c[4]
        ln = np.hypot(ex, ey)
        off = ((c[0] - c[3]) * ey - (c[1] - c[4]) * ex) / ln + c[7]   # the centre's offset from the line
        c[2][0::16] = np.nan
        c[3][1::16] = np.inf
        c[5][2::16], c[6][2::16] = c[3][2::16], c[4][2::16]          # the line is a point
        c[2][3::16] = 0.0                                            # zero radius
        c[2][4::16] = np.abs(off)[4::16]                             # tangent line: a double root, linear convergence into det = 0
        c[2][5::16] = 0.5 * np.abs(off)[5::16]                       # the line misses the circle: no real root
        c[0][6::16] = 1e300
        c[2][7::16] = np.abs(off)[7::16] * (1 + 1e-12)               # a hair inside tangency
        c[7][8::16] -= off[8::16]                                    # the line through the centre: c = 0, h = r (well conditioned)
        return (0, 1, 2, 6)
    if kind == 2:  # dX x + dY y + (s1 - s2) = 0 and the unit circle
        dX, dY = c[2] - c[0], c[3] - c[1]
        ln = np.hypot(dX, dY)
        c[4][0::16] = np.nan
        c[0][1::16] = np.inf
        c[2][2::16], c[3][2::16] = c[0][2::16], c[1][2::16]          # p2 = p1: no line
        c[4][3::16], c[5][3::16] = ln[3::16], 0.0                    # |s1 - s2| = |p2 - p1|: tangent
        c[4][4::16], c[5][4::16] = 2.0 * ln[4::16], 0.0              # |s1 - s2| > |p2 - p1|: no root
        c[4][5::16] = c[5][5::16]                                    # s1 = s2: the line through the origin
        c[2][6::16] = 1e300
        c[4][7::16], c[5][7::16] = ln[7::16] * (1 - 1e-12), 0.0      # a hair inside tangency
        return (0, 1, 2, 6)
    assert kind == 5  # fdy x - fdx y = cosA |fd| and the unit circle
    c[2][0::16] = np.nan
    c[0][1::16] = np.inf
    c[0][2::16], c[1][2::16] = 0.0, 0.0                              # no direction
    c[2][3::16] = 1.0                                                # cosA = 1: tangent
    c[2][4::16] = 1.5                                                # no root
    c[2][5::16] = 0.0                                                # right angle: the line through the origin
    c[0][6::16] = 1e300
    c[2][7::16] = -1.0 + 1e-12                                       # a hair inside tangency
    return (0, 1, 2, 6)


@pytest.mark.parametrize("variant", [5, 6, 8])
@pytest.mark.parametrize("kind", [2, 3, 5])
def test_contract_degenerate_inputs_of_the_line_form_kinds(gpu, gcs, kind, variant):
    n = 4096

    def make():
        hb = gcs.synth.make(kind, n, seed=0xDE6 + kind)
        make.literal = _degenerate_rows(kind, hb.cols)
        return hb
    hb = make()
    hb.variant = variant
    gpu.solve_host(hb.alloc_outputs(), 0)
    ref = O.solve(make().alloc_outputs())
    assert_batches_within_contract(hb, ref, f"K{kind} variant {variant}, degenerate inputs")
    for o in make.literal:
        a, b = hb.cand[..., o::16], ref.cand[..., o::16]
        assert ((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))).all(), f"K{kind} row class {o} did not come out of the literal code"
