"""Parity soak against the REFERENCE BUILD (oracle/_ref: the reference's own solve2D / primitives /
heuristics compiled from /root/reference): CUDA path vs reference on fresh seeds.
Usage: python scratch/soak_ref.py [n_per_case] [variant]   (variant 5 = contracted: compared to the
north star's contract - counts, flags, roots equal, coordinates 1e-9 - instead of bit for bit)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
import ref_lib as R
from test_golden import check_against_golden, check_against_golden_contract
capi, synth = gcs.capi, gcs.synth
capi.init([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(2)
total, t0 = 0, time.time()
for seed in (0xA11CE, 0xB0B):
    for kind in (1, 2, 3, 4, 5):
        a = synth.make(kind, n, seed=seed + kind)
        a.want_cand = True
        a.variant = variant
        a.alloc_outputs()
        capi.solve_host(a, 0)
        b = synth.make(kind, n, seed=seed + kind).alloc_outputs()
        os.dup2(devnull, 2)
        R.solve_batch(b, count_iters=True)
        os.dup2(saved, 2)
        z = {"cand": b.cand, "iters": b.iters, "converged": b.converged, "root": b.root_index, "out": np.stack(b.out)}
        total += n
        if variant >= 5:
            worst = check_against_golden_contract(a, z, kind, f"K{kind} seed {seed:#x}")
            print(f"K{kind} seed {seed:#x} n={n}: iteration counts, flags, roots equal to the reference build; candidates and chosen results within {worst:.2e} relative", flush=True)
        else:
            check_against_golden(a, z, kind, f"K{kind} seed {seed:#x}")
            print(f"K{kind} seed {seed:#x} n={n}: candidates, iteration counts, flags, roots, chosen results identical to the reference build", flush=True)
print(f"soak vs reference build (variant {variant}): {total} sub-systems, all {'within the contract' if variant >= 5 else 'identical'}, {time.time()-t0:.0f} s")
