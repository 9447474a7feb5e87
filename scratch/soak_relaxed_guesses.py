"""Soak of the contracted variant under ARBITRARY seeds (explicit guesses): uniform boxes of several
sizes, seeds next to the roots, seeds next to the line where the Jacobian is singular, tiny and huge
seeds - every kind, against the CPU oracle (counts / flags / roots equal, coordinates 1e-9).
Usage: python scratch/soak_relaxed_guesses.py [n_per_case]"""
import importlib, os, sys, time
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
import oracle_lib
from util import assert_batches_within_contract
capi, synth = gcs.capi, gcs.synth
capi.init([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
rng = np.random.default_rng(2026)
total = bad = 0
t0 = time.time()


def guesses(kind, mode, base):
    """[2][2][n] seeds for the batch `base` (already solved by the oracle from its default seeds)."""
    unit = kind in (2, 5)  # unknowns are unit normals
    size = 1.0 if unit else 1000.0
    if mode == "box":
        return rng.uniform(-3 * size, 3 * size, size=(2, 2, n))
    if mode == "wide":
        return rng.uniform(-1.0, 1.0, size=(2, 2, n)) * np.exp(rng.uniform(np.log(1e-4), np.log(1e6), size=(2, 1, n))) * size
    if mode == "near_root":
        g = np.array(base.cand, copy=True)
        return g * (1.0 + rng.normal(0, 1.0, size=g.shape) * np.exp(rng.uniform(np.log(1e-12), np.log(0.3), size=(2, 1, n))))
    if mode == "tiny":
        return rng.uniform(-3e-5, 3e-5, size=(2, 2, n))
    if mode == "singular":
        if kind == 1:
            ax, ay, _, bx, by, _ = base.cols
            t = rng.uniform(-0.5, 1.5, size=(2, n))
            off = np.exp(rng.uniform(np.log(1e-12), np.log(1e-1), size=(2, n))) * rng.choice([-1.0, 1.0], size=(2, n))
            g = np.empty((2, 2, n))
            g[:, 0, :] = ax + t * (bx - ax) - off * (by - ay)
            g[:, 1, :] = ay + t * (by - ay) + off * (bx - ax)
            return g
        # the other kinds: a seed with the determinant of the Jacobian nearly zero is a seed nearly
        # perpendicular (unit kinds) to / nearly on the foot line of the linear equation: perturb the
        # mid point of the two roots
        mid = 0.5 * (base.cand[0] + base.cand[1])
        return mid[None] * (1.0 + rng.normal(0, 1, size=(2, 2, n)) * np.exp(rng.uniform(np.log(1e-12), np.log(1e-2), size=(2, 1, n))))
    raise ValueError(mode)


for kind in (1, 2, 3, 4, 5):
    base = synth.make(kind, n, seed=0xC0DE + kind)
    base.want_cand = True
    oracle_lib.solve(base.alloc_outputs(), threads=0)
    for mode in ("box", "wide", "near_root", "tiny", "singular"):
        g = np.ascontiguousarray(guesses(kind, mode, base))
        a = synth.make(kind, n, seed=0xC0DE + kind)
        a.guesses, a.variant, a.want_cand = g, int(os.environ.get("SOAK_VARIANT", capi.VARIANT_CONTRACTED)), True
        b = synth.make(kind, n, seed=0xC0DE + kind)
        b.guesses, b.want_cand = g, True
        capi.solve_host(a.alloc_outputs(), 0)
        oracle_lib.solve(b.alloc_outputs(), threads=0)
        try:
            worst = assert_batches_within_contract(a, b, f"K{kind} {mode}")
            verdict = f"within contract (max rel err {worst:.2e})"
        except AssertionError as e:
            bad += 1
            verdict = "VIOLATION: " + str(e)[:300]
        total += n
        print(f"K{kind} seeds={mode}: {verdict}; iters {int(b.iters.min())}..{int(b.iters.max())}, converged {float(np.mean(b.converged)):.4f}, "
              f"candidate words not bit-identical {float(np.mean(a.cand.view(np.uint64) != b.cand.view(np.uint64))):.3f}", flush=True)
print(f"relaxed soak with explicit seeds: 25 cases, {total} sub-systems, {bad} cases violate the contract, {time.time()-t0:.0f} s")
