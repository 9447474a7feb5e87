"""Cost of one small host-buffer call (a wave of the scheduler).  Usage: python scratch/small_call.py [n]"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for label, pinned in (("pageable columns", False), ("pinned slab", True)):
    h = synth.make_pp(n)
    if pinned:
        t = torch.empty((6, n), dtype=torch.float64, pin_memory=True); slab = t.numpy(); slab[...] = np.stack(h.cols)
        tc = torch.empty(n, dtype=torch.uint8, pin_memory=True); code = tc.numpy(); code[...] = h.code
        hb = capi.HostBatch(1, 2, [slab[c] for c in range(6)], code, None, 0, want_cand=False)
        to = torch.empty((2, n), dtype=torch.float64, pin_memory=True); o = to.numpy()
        hb.out = [o[0], o[1]]
        ti = torch.empty((2, n), dtype=torch.int16, pin_memory=True); hb.iters = ti.numpy()
        tv = torch.empty((2, n), dtype=torch.uint8, pin_memory=True); hb.converged = tv.numpy()
        tr = torch.empty(n, dtype=torch.uint8, pin_memory=True); hb.root_index = tr.numpy()
        hb.cand = None
    else:
        hb = h
        hb.want_cand = False
        hb.alloc_outputs()
    import ctypes as C
    cb = hb.cbatch(); lib = capi.load()
    for _ in range(20): lib.gcs_b200_solve_host(C.byref(cb), 0)
    t0 = time.perf_counter()
    for _ in range(200): lib.gcs_b200_solve_host(C.byref(cb), 0)
    print(f"{label}: {(time.perf_counter()-t0)/200*1e6:.1f} us per gcs_b200_solve_host call, n={n}")
