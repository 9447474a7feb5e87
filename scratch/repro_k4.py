"""The soak case that found the K4 far-seed hole (tests/test_gpu_soak.py::test_explicit_seeds[wide-4] on
the build with source-hash offset 170584801): one run in 65536 with a different iteration count."""
import importlib, os, sys
import numpy as np
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
import oracle_lib as O
import test_gpu_soak as T
capi, synth = gcs.capi, gcs.synth
capi.init([0])
offset, kind, n = 170584801, 4, 1 << 16
rng = np.random.default_rng(2026 + offset + 31 * kind)
base = synth.make(kind, n, seed=0xC0DE + kind + offset); base.want_cand = True
O.solve(base.alloc_outputs(), threads=0)
g = np.ascontiguousarray(T._guesses(rng, kind, "wide", base, n))
ref = synth.make(kind, n, seed=0xC0DE + kind + offset); ref.guesses = g
O.solve(ref.alloc_outputs())
for variant in (5, 6, 7, 8):
    a = synth.make(kind, n, seed=0xC0DE + kind + offset); a.guesses, a.variant = g, variant
    capi.solve_host(a.alloc_outputs(), 0)
    bad = np.argwhere(a.iters != ref.iters)
    print(f"variant {variant}: {len(bad)} runs with a different iteration count", bad[:4].tolist(),
          [(int(a.iters[s, i]), int(ref.iters[s, i]), g[s, :, i].tolist()) for s, i in bad[:2]])
