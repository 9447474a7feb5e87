"""Kernel-only timing of every kind (device-resident, L2 flushed), + bit-check against a baseline
library.  Usage: [GCS_B200_LIB=alt.so] python scratch/kbench.py [variant] [kinds] [n] [n_seeds]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kinds = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 5]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 19
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
for kind in kinds:
    hb = synth.make(kind, n, n_seeds=ns) if ns != 2 else synth.make(kind, n)
    db = capi.DeviceBatch(hb, "cuda:0", want_cand=False, variant=variant)
    for _ in range(3):
        db.solve()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); db.solve(); e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    it = db.iters.cpu().numpy()
    extra = ""
    if variant >= 5:
        import ctypes as C
        st2 = (C.c_uint64 * 8)()
        capi.load().gcs_b200_contracted_stats_ex(0, st2, 1)
        db.solve()
        capi.load().gcs_b200_contracted_stats_ex(0, st2, 1)
        extra = (f"  literal re-runs per launch of {n * ns} runs: cond {st2[0]} selection {st2[1]} bounce {st2[2]} band {st2[3]} "
                 f"cap {st2[4]} huge {st2[5]}")
    chk = int(db.out[0].view(torch.int64).sum().item()) ^ int(db.out[1].view(torch.int64).sum().item()) ^ int(it.astype(np.int64).sum()) ^ (int(db.root_index.sum().item()) << 20)
    print(f"K{kind} n {n} seeds {ns} variant {variant}: median {np.median(ts)*1e3:.1f} us  min {np.min(ts)*1e3:.1f} us  checksum {chk & 0xffffffffffff:012x}{extra}")
