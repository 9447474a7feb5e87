"""One small contracted K1 launch (2^13 sub-systems = 128 CTAs: fewer than one per SM), L2 flushed: the
life of a CTA that has an SM to itself (profiles/r2_launch_vs_size.md).  For ncu --set full."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 13
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 5
db = capi.DeviceBatch(synth.make_pp(n, first=9 * (1 << 13)), "cuda:0", want_cand=False, variant=variant)  # window 9: no literal re-run in it
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.fill_(1)
    db.solve()
torch.cuda.synchronize()
