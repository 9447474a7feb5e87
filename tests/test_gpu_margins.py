"""Property test of the contracted kernels' guards, independent of the kernels' own arithmetic.

Claim under test (csrc/newton_relaxed.cuh, G3): a Newton run the guards ACCEPT on the closed-form path
had every convergence decision further from the threshold 1e-5 than the stated margin -
first level 2^-36 S + 2^-16 tol, second level / careful mode 2^-44 S dr/|det J| + 2^-40 tol.  The
kernel reports, through the test hook gcs_b200_debug_path_buffer, how each run was decided; the CPU
checker re-runs every run with the LITERAL arithmetic and records how close its own update lengths
came to the threshold (oracle/gcs_oracle.c, newton2d_decision_slack).  If a guard-accepted run's
literal trajectory ever sat within HALF the stated margin of the threshold, the guard logic (integer
thresholds, carry term, careful mode) does not implement what it states: the test fails.  Half,
because the two arithmetics' own update lengths differ by a small fraction of the margin.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from util import assert_batches_within_contract

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _scales(kind, cols):
    """(S, dr, step_scale) of every sub-system as Rsys<KIND>::load derives them."""
    k = [np.zeros_like(cols[2]) if c is None else c for c in cols]
    a = np.abs
    if kind == 1:
        return a(k[0]) + a(k[1]) + a(k[3]) + a(k[4]) + a(k[2]) + a(k[5]), a(k[2] * k[5]), 0.5
    if kind == 2:
        l1 = a(k[2] - k[0]) + a(k[3] - k[1])
        return 1.0 + (a(k[4]) + a(k[5])) / l1, 2.0 * l1 * 0.70710678118654746, 1.0
    if kind == 3:
        ln = np.sqrt((k[5] - k[3]) ** 2 + (k[6] - k[4]) ** 2)
        return a(k[0]) + a(k[1]) + a(k[2]) + a(k[3]) + a(k[4]) + a(k[7]), 2.0 * a(k[2]) * ln, 1.0
    if kind == 4:
        l1 = np.sqrt((k[2] - k[0]) ** 2 + (k[3] - k[1]) ** 2)
        l2 = np.sqrt((k[7] - k[5]) ** 2 + (k[8] - k[6]) ** 2)
        return a(k[0]) + a(k[1]) + a(k[5]) + a(k[6]) + a(k[4]) + a(k[9]), l1 * l2, 1.0
    ln = np.sqrt(k[0] ** 2 + k[1] ** 2)
    return np.full_like(ln, 2.0), 2.0 * ln, 1.0


CASES = [
    ("K1 bench-like", 1, dict(seed=0x5EED0001)),
    ("K1 flat 1e-3", 1, dict(seed=77, scale=30.0, flat=1e-3)),
    ("K1 flat 1e-4", 1, dict(seed=78, scale=1e3, flat=1e-4)),
    ("K1 scale 1e-6 (carry regime)", 1, dict(seed=79, scale=1e-6)),
    ("K1 scale 3e-5", 1, dict(seed=80, scale=3e-5)),
    ("K2", 2, dict(seed=81)),
    ("K3", 3, dict(seed=82)),
    ("K5", 5, dict(seed=83)),
]


@pytest.mark.parametrize("what,kind,kw", CASES, ids=[c[0] for c in CASES])
def test_guard_accepted_runs_kept_their_distance_from_the_threshold(gpu, gcs, what, kind, kw):
    import torch
    capi, synth = gcs.capi, gcs.synth
    n = 1 << 18
    hb = synth.make(kind, n, **kw)
    db = capi.DeviceBatch(hb, "cuda:0", want_cand=True, variant=capi.VARIANT_CONTRACTED_STATIC)
    path = torch.full((2, n), 255, dtype=torch.uint8, device="cuda:0")
    lib = capi.load()
    capi.check(lib.gcs_b200_debug_path_buffer(0, C.c_void_p(path.data_ptr()), 2 * n))
    try:
        db.solve()
        torch.cuda.synchronize()
    finally:
        capi.check(lib.gcs_b200_debug_path_buffer(0, None, 0))
    got = db.to_host(synth.make(kind, n, **kw))
    path = path.cpu().numpy()
    assert path.max() <= 4, "a run did not report its path"

    S, dr, step_scale = _scales(kind, hb.cols)
    # the kernels take |det| of the matrix they solve: K1 works on J/2, whose determinant is det(J)/4
    sdr = 2.0 ** -44 * S * dr / (step_scale * step_scale)
    band = 2.0 ** -36 * S + 2.0 ** -16 * TOL
    ref = synth.make(kind, n, **kw)
    slack = O.decision_slack(ref.alloc_outputs(), sdr, band)
    assert_batches_within_contract(got, ref, what)

    closed = path <= 2
    first = (path <= 1) & closed
    careful = path == 2
    # every run accepted on the closed-form path kept its distance (plane 0: min(first level, second level))
    bad = first & ~(slack[0] >= 0.0)
    assert not bad.any(), f"{what}: {int(bad.sum())} guard-accepted runs came within half the stated margin of the threshold, " \
        f"worst slack {np.nanmin(np.where(bad, slack[0], np.inf)):.3e}"
    # careful mode claims the conditioning-scaled margin at every late decision (plane 1)
    badc = careful & ~(slack[1] >= 0.0)
    assert not badc.any(), f"{what}: {int(badc.sum())} careful-mode runs inside half their margin"
    # runs handed to the literal code come out of it bit for bit
    lit = path >= 3
    if lit.any():
        same = got.cand.view(np.uint64) == ref.cand.view(np.uint64)
        same |= np.isnan(got.cand) & np.isnan(ref.cand)
        assert same[:, 0, :][lit].all() and same[:, 1, :][lit].all(), f"{what}: a literal re-run is not bit-identical to the checker"
    stats = {int(v): int((path == v).sum()) for v in range(5)}
    print(f"{what}: decided by first level / second level / careful / literal (run) / literal (selection): {stats}")
    if "flat" in what:
        assert stats[2] > 0, "the flat-triangle case is meant to exercise careful mode"


def test_path_buffer_is_a_test_hook_and_off_by_default(gpu, gcs):
    import torch
    capi, synth = gcs.capi, gcs.synth
    n = 4096
    db = capi.DeviceBatch(synth.make_pp(n), "cuda:0", variant=capi.VARIANT_CONTRACTED)
    path = torch.full((2, n), 255, dtype=torch.uint8, device="cuda:0")
    db.solve()  # hook not set: nothing is written
    torch.cuda.synchronize()
    assert int(path.min()) == 255
    lib = capi.load()
    capi.check(lib.gcs_b200_debug_path_buffer(0, C.c_void_p(path.data_ptr()), n))  # too small for 2 * n runs: ignored
    db.solve()
    torch.cuda.synchronize()
    capi.check(lib.gcs_b200_debug_path_buffer(0, None, 0))
    assert int(path.min()) == 255
