"""Launch time against launch size (the intercept = what does not scale: launch, one CTA's life, re-run tail).
Usage: [GCS_B200_LIB=alt.so] python scratch/ksize.py [variant] [kind]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
xs, ys = [], []
for lg in range(13, 22):
    n = 1 << lg
    db = capi.DeviceBatch(synth.make(kind, n), "cuda:0", want_cand=False, variant=variant)
    for _ in range(3):
        db.solve()
    ts = []
    for _ in range(15):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); db.solve(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    xs.append(n); ys.append(float(np.median(ts)))
    print(f"K{kind} variant {variant} n 2^{lg}: median {np.median(ts):.1f} us  min {np.min(ts):.1f} us")
a, b = np.polyfit(np.array(xs[3:]) / (1 << 19), np.array(ys[3:]), 1)
print(f"fit over 2^16..2^21: {b:.1f} us + {a:.1f} us per 2^19")
