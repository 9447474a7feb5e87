"""K4 (two linear equations: two updates per seed) without the never-converging parallel rows of the
parity generator: the HBM-bound kind.  Usage: python scratch/k4_hbm.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
for n in (1 << 19, 1 << 22):
    hb = synth.make_pll(n, parallel_every=0)
    for variant in (5, 8, 9):
        db = capi.DeviceBatch(hb, "cuda:0", want_cand=False, variant=variant)
        for _ in range(3): db.solve()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); db.solve(); e1.record(st); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts)) * 1e-3
        it = db.iters.cpu().numpy()
        print(f"K4 n={n} variant {variant}: {t*1e6:.1f} us, iters {it.min()}..{it.max()}, algorithmic {120*n/t/1e9:.0f} GB/s ({120*n/t/6.54e12*100:.0f} % of 6540 GB/s)")
