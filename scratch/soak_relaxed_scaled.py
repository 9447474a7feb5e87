"""Soak of the contracted variant on RESCALED systems of the kinds whose generators have no scale
option: every length column of K2..K5 multiplied by 1e-7 .. 1e6 (unit-less columns - normals,
cos(angle) - untouched), so that the absolute 1e-5 threshold sits above, inside and below the size
of the system.  Against the CPU oracle: counts / flags / roots equal, coordinates 1e-9.
Usage: python scratch/soak_relaxed_scaled.py [n_per_case]"""
import importlib, os, sys, time
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
import oracle_lib
from util import assert_batches_within_contract
capi, synth = gcs.capi, gcs.synth
capi.init([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
LENGTH_COLS = {2: [0, 1, 2, 3, 4, 5, 8], 3: list(range(10)), 4: list(range(12)), 5: [0, 1, 7, 8, 9, 10, 11, 12]}
total = bad = 0
t0 = time.time()
for kind in (2, 3, 4, 5):
    for scale in (1e-7, 1e-6, 3e-5, 1e-3, 1e3, 1e6):
        def build():
            hb = synth.make(kind, n, seed=0x5CA1E + kind)
            for c in LENGTH_COLS[kind]:
                hb.cols[c] = np.ascontiguousarray(hb.cols[c] * scale)
            return hb
        a = build(); a.variant = capi.VARIANT_CONTRACTED; a.alloc_outputs()
        b = build().alloc_outputs()
        capi.solve_host(a, 0)
        oracle_lib.solve(b, threads=0)
        try:
            worst = assert_batches_within_contract(a, b, f"K{kind} x{scale}")
            verdict = f"within contract (max rel err {worst:.2e})"
        except AssertionError as e:
            bad += 1
            verdict = "VIOLATION: " + str(e)[:300]
        total += n
        print(f"K{kind} lengths x {scale:g} n={n}: {verdict}; iters {int(b.iters.min())}..{int(b.iters.max())}, converged {float(np.mean(b.converged)):.4f}, "
              f"candidate words not bit-identical {float(np.mean(a.cand.view(np.uint64) != b.cand.view(np.uint64))):.3f}", flush=True)
print(f"relaxed soak on rescaled systems: 24 cases, {total} sub-systems, {bad} cases violate the contract, {time.time()-t0:.0f} s")
