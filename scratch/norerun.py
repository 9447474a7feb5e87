"""What one literal re-run costs a launch: windows of 2^13 K1 sub-systems (128 CTAs: fewer than one per SM)
of the bench stream, contracted static kernel; launch time (events, L2 flushed) next to the number of
runs the guards handed to the literal code in that window."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
lib = capi.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
n = 1 << 13
rows = []
for w in range(24):
    db = capi.DeviceBatch(synth.make_pp(n, first=w * n), "cuda:0", want_cand=False, variant=5)
    for _ in range(3):
        db.solve()
    ts = []
    for _ in range(9):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); db.solve(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    s8 = (C.c_uint64 * 8)()
    lib.gcs_b200_contracted_stats_ex(0, s8, 1)
    db.solve()
    lib.gcs_b200_contracted_stats_ex(0, s8, 1)
    rows.append((int(sum(s8)), float(np.median(ts))))
    print(f"window {w:2d}: literal re-runs {int(sum(s8))} (cond {s8[0]} band {s8[3]}), launch median {np.median(ts):.1f} us min {np.min(ts):.1f} us", flush=True)
a = np.array(rows)
for k in sorted(set(a[:, 0])):
    print(f"windows with {int(k)} re-runs: {int((a[:, 0] == k).sum())}, median launch {np.median(a[a[:, 0] == k, 1]):.1f} us")
